#!/bin/bash
# epilogue with one x64 TMEM load per pass, the next pass's load in flight behind the staging/store (TFA_EPI_PIPELINED)
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
export TFA_NO_BUILD=1
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_epi.so timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b25_tests_epi.log 2>&1; echo "tests(epi) rc=$?"; tail -3 gpurun_out/b25_tests_epi.log | cut -c1-200
CFG='[[4,32,4096,128,true],[8,32,4096,128,true],[16,32,1024,128,true],[32,32,512,128,true],[4,32,4096,128,false],[4,16,2048,64,false],[16,16,1024,64,false]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "persist|TFA_KERNEL=persist|" "epi|TFA_KERNEL=persist|libtfa_b200_epi.so" > gpurun_out/b25_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b25_ab.log
