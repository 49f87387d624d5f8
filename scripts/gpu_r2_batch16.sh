#!/bin/bash
mkdir -p gpurun_out
export TFA_NO_BUILD=1
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_defer.so timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_lazy_rescale.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b16_tests_defer.log 2>&1; echo "tests(defer) rc=$?"; tail -3 gpurun_out/b16_tests_defer.log | cut -c1-200
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "persist|TFA_KERNEL=persist|" "persist-defer|TFA_KERNEL=persist|libtfa_b200_defer.so" > gpurun_out/b16_ab.log 2>&1; echo "ab rc=$?"; head -24 gpurun_out/b16_ab.log
