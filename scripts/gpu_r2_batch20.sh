#!/bin/bash
# speculative first quarter (TFA_SPEC_Q0) and split S load (TFA_LD_SPLIT) in the persistent kernel: parity + A/B
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
export TFA_NO_BUILD=1
for v in spec specld; do
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_$v.so timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_lazy_rescale.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b20_tests_$v.log 2>&1; echo "tests($v) rc=$?"; tail -3 gpurun_out/b20_tests_$v.log | cut -c1-200
done
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false],[4,32,4096,64,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "persist|TFA_KERNEL=persist|" "spec|TFA_KERNEL=persist|libtfa_b200_spec.so" "specld|TFA_KERNEL=persist|libtfa_b200_specld.so" > gpurun_out/b20_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b20_ab.log
