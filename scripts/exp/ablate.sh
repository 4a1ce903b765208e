#!/bin/bash
# True critical-path share of each kernel family inside the CUDA-graph replay of the step: the family's launches are skipped
# (QEB_DBG_SKIP, results are garbage) and the step is timed again. usage: scripts/exp/ablate.sh [workload]
WL=${1:-prep_step}
run() { QEB_DBG_SKIP=$1 python bench.py --workload $WL --no-extras --skip-cpu-baseline --skip-eager --skip-profile --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-46s %.3f ms' % ('$1', d['ms_per_step']))"; }
run none
for t in tc_conv_fprop tc_conv_wgrad tc_convT tc_head bn_apply bn_bwd_apply bn_bwd_reduce bn_stats bn_finalize bn_ maxpool_fwd maxpool_bwd colsum pack lstm_fwd lstm_bwd c1_conv o1_conv memset vec_add ctc log_softmax mse adam jitter "bn_,maxpool,colsum,pack,c1_conv,o1_conv,memset,vec_add" "tc_conv_fprop,tc_conv_wgrad,tc_convT,tc_head"; do run $t; done
