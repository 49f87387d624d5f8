#!/bin/bash
# two-pass softmax (TFA_TWO_PASS): quarters 1..3 re-read S from TMEM so ptxas cannot hoist their FMA work: parity + A/B
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
export TFA_NO_BUILD=1
for v in ${VARIANTS:-twopass}; do
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_$v.so timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_lazy_rescale.py tests/test_general_attn.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b21_tests_$v.log 2>&1; echo "tests($v) rc=$?"; tail -3 gpurun_out/b21_tests_$v.log | cut -c1-200
done
ARGS=("persist|TFA_KERNEL=persist|")
for v in ${VARIANTS:-twopass}; do ARGS+=("$v|TFA_KERNEL=persist|libtfa_b200_$v.so"); done
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false],[4,32,4096,64,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "${ARGS[@]}" > gpurun_out/b21_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b21_ab.log
