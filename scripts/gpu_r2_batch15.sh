#!/bin/bash
# Round-2 GPU batch 15: two UMMA issuer warps (one per Q tile) in the persistent kernel
mkdir -p gpurun_out
export TFA_NO_BUILD=1
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_two.so timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py tests/test_fwd_properties.py tests/test_lazy_rescale.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b15_tests_two.log 2>&1; echo "tests(two) rc=$?"; tail -3 gpurun_out/b15_tests_two.log | cut -c1-200
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false],[4,32,4096,64,true],[64,32,4096,128,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "classic|TFA_KERNEL=classic|" "persist|TFA_KERNEL=persist|" "persist-two|TFA_KERNEL=persist|libtfa_b200_two.so" > gpurun_out/b15_ab.log 2>&1; echo "ab rc=$?"; head -34 gpurun_out/b15_ab.log
