#!/bin/bash
# Round-2 GPU batch 2: full suites (default / persist / persist64), persist tuning variants, timelines.
mkdir -p gpurun_out
export TFA_NO_BUILD=1
timeout 300 python -m pytest tests/test_lazy_rescale.py tests/test_fused_exchange.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b2_gpu_tests_default_new.log 2>&1; echo "new tests(default) rc=$?"; tail -3 gpurun_out/b2_gpu_tests_default_new.log | cut -c1-200
for v in persist persist64; do
  [ "$v" = "default" ] && unset TFA_KERNEL || export TFA_KERNEL=$v
  timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b2_gpu_tests_$v.log 2>&1; echo "gpu_tests($v) rc=$?"; tail -6 gpurun_out/b2_gpu_tests_$v.log | cut -c1-200
done
unset TFA_KERNEL
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "default||" "persist|TFA_KERNEL=persist|" "noqpf|TFA_KERNEL=persist|libtfa_b200_noqpf.so" "nohoist|TFA_KERNEL=persist|libtfa_b200_nohoist.so" "r224|TFA_KERNEL=persist|libtfa_b200_r224.so" "alt|TFA_KERNEL=persist|libtfa_b200_alt.so" "persist64|TFA_KERNEL=persist64|" > gpurun_out/b2_ab.log 2>&1; echo "ab rc=$?"
cp gpurun_out/ab.log gpurun_out/b2_ab_full.log 2>/dev/null
for shape in '{"B":1,"H":32,"S":16384,"D":128,"causal":true,"block":5,"limit":260}' '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":5,"limit":420}'; do
  tag=$(echo $shape | python -c "import sys,json; d=json.load(sys.stdin); print('S%d'%d['S'])")
  TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py "$shape" > gpurun_out/b2_trace_default_$tag.txt 2>&1; echo "trace default $tag rc=$?"
  TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py "$shape" > gpurun_out/b2_trace_persist_$tag.txt 2>&1; echo "trace persist $tag rc=$?"
  tail -3 gpurun_out/b2_trace_default_$tag.txt; tail -3 gpurun_out/b2_trace_persist_$tag.txt
done
