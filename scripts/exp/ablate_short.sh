run() { QEB_DBG_SKIP=$1 python bench.py --no-extras --skip-cpu-baseline --skip-eager --skip-profile --steps 40 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%-46s %.3f ms' % ('$1', d['ms_per_step']))"; }
run none
for t in tc_conv_fprop.f16 tc_conv_fprop tc_conv_wgrad bn_ maxpool lstm pack colsum "bn_,maxpool,colsum,pack,c1_conv,o1_conv,memset,vec_add" "tc_conv_fprop,tc_conv_wgrad,tc_convT,tc_head"; do run $t; done
