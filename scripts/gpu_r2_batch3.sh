#!/bin/bash
# Round-2 GPU batch 3: lean issuer A/B + timeline.
mkdir -p gpurun_out
export TFA_NO_BUILD=1
TFA_KERNEL=persist timeout 300 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py tests/test_lazy_rescale.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b3_tests_persist.log 2>&1; echo "tests(persist) rc=$?"; tail -3 gpurun_out/b3_tests_persist.log | cut -c1-200
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "default||" "persist|TFA_KERNEL=persist|" "alt|TFA_KERNEL=persist|libtfa_b200_alt.so" "noqpf|TFA_KERNEL=persist|libtfa_b200_noqpf.so" > gpurun_out/b3_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b3_ab.log
for shape in '{"B":1,"H":32,"S":16384,"D":128,"causal":true,"block":5,"limit":10}' '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":5,"limit":600}'; do
  tag=$(echo $shape | python -c "import sys,json; d=json.load(sys.stdin); print('S%d'%d['S'])")
  TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py "$shape" > gpurun_out/b3_trace_persist_$tag.txt 2>&1; echo "trace persist $tag rc=$?"
  tail -3 gpurun_out/b3_trace_persist_$tag.txt
done
