#!/bin/bash
# compute-sanitizer over small runs of both networks (forward + backward, train and frozen BatchNorm, the window conv
# variant forced on with QEB_WIN=2) and of the integer / fused kernels. Output: gpurun_out/sanitizer_*.log
# (copy the summaries to profiles/). Each tool run is bounded by `timeout`.
OUT=${1:-gpurun_out}
mkdir -p $OUT
export QEB_WIN=2
for tool in memcheck racecheck; do
  for what in crnn unet aux; do
    script=scripts/dev_san.py; arg=$what
    if [ $what = aux ]; then script=scripts/dev_san_aux.py; arg=; fi
    echo "=== compute-sanitizer --tool $tool $script $arg" | tee -a $OUT/sanitizer_$tool.log
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python $script $arg 2>&1 | grep -v "^\s*$" | tail -25 >> $OUT/sanitizer_$tool.log
    echo "exit: ${PIPESTATUS[0]}" >> $OUT/sanitizer_$tool.log
  done
done
tail -4 $OUT/sanitizer_memcheck.log $OUT/sanitizer_racecheck.log
