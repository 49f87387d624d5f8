#!/bin/bash
# Round-2 GPU batch 7: persist with the warp-uniform item index: tests, A/B, timeline
mkdir -p gpurun_out
export TFA_NO_BUILD=1
for v in persist persist64; do
TFA_KERNEL=$v timeout 400 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py tests/test_lazy_rescale.py tests/test_fwd_properties.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b7_tests_$v.log 2>&1; echo "tests($v) rc=$?"; tail -3 gpurun_out/b7_tests_$v.log | cut -c1-200
done
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false],[4,32,4096,64,true],[16,16,1024,64,false]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "default||" "persist|TFA_KERNEL=persist|" "alt|TFA_KERNEL=persist|libtfa_b200_alt.so" "noqpf|TFA_KERNEL=persist|libtfa_b200_noqpf.so" "persist64|TFA_KERNEL=persist64|" > gpurun_out/b7_ab.log 2>&1; echo "ab rc=$?"; head -52 gpurun_out/b7_ab.log
for shape in '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":5,"limit":700}'; do
  TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py "$shape" > gpurun_out/b7_trace_persist_S4096.txt 2>&1; echo "trace rc=$?"
  tail -3 gpurun_out/b7_trace_persist_S4096.txt
done
