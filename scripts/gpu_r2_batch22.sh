#!/bin/bash
# column-split softmax (TFA_KERNEL=split): parity + A/B against persist and classic
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
export TFA_NO_BUILD=1
TFA_KERNEL=split timeout 900 python -m pytest tests/test_fwd_parity.py tests/test_lazy_rescale.py tests/test_general_attn.py tests/test_fused_exchange.py -m gpu -x -q --no-header -p no:cacheprovider > gpurun_out/b22_tests_split.log 2>&1; echo "tests(split) rc=$?"; tail -5 gpurun_out/b22_tests_split.log | cut -c1-300
CFG=${CFG:-'[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false],[4,32,4096,64,true]]'} \
  timeout 900 bash scripts/gpu_ab_env.sh "persist|TFA_KERNEL=persist|" "split|TFA_KERNEL=split|" "split112|TFA_KERNEL=split|libtfa_b200_r112.so" > gpurun_out/b22_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b22_ab.log
