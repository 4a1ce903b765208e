#!/bin/bash
# compute-sanitizer, ONE tool per gpurun call (B200_PROFILING.md): scripts/sanitize.sh memcheck|racecheck|synccheck|initcheck
# over one small process that runs both networks (forward + backward, train and frozen BatchNorm, the window conv variant
# forced on), the fused head / jittered forward and the integer kernels. Output: gpurun_out/sanitizer_<tool>.log (copy to
# profiles/). Bounded by `timeout`.
TOOL=${1:-memcheck}
OUT=gpurun_out
mkdir -p $OUT
export QEB_WIN=2
python scripts/dev_san_all.py > $OUT/sanitizer_plain.log 2>&1 || { echo "plain run failed"; tail -5 $OUT/sanitizer_plain.log; exit 1; }
echo "=== compute-sanitizer --tool $TOOL python scripts/dev_san_all.py (QEB_WIN=2)" > $OUT/sanitizer_$TOOL.log
timeout 1200 compute-sanitizer --tool $TOOL --print-limit 20 python scripts/dev_san_all.py 2>&1 | grep -v "^\s*$" | grep -v "Warning\|warn" | tail -40 >> $OUT/sanitizer_$TOOL.log
echo "exit: ${PIPESTATUS[0]}" >> $OUT/sanitizer_$TOOL.log
tail -6 $OUT/sanitizer_$TOOL.log
