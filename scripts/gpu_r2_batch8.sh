#!/bin/bash
# Round-2 GPU batch 8: persist with early first loads / first item without atomics
mkdir -p gpurun_out
export TFA_NO_BUILD=1
TFA_KERNEL=persist timeout 400 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b8_tests_persist.log 2>&1; echo "tests(persist) rc=$?"; tail -3 gpurun_out/b8_tests_persist.log | cut -c1-200
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false],[4,32,4096,64,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "default||" "persist|TFA_KERNEL=persist|" "persist64|TFA_KERNEL=persist64|" > gpurun_out/b8_ab.log 2>&1; echo "ab rc=$?"; head -30 gpurun_out/b8_ab.log
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":5,"limit":700}' > gpurun_out/b8_trace_persist_S4096.txt 2>&1; echo "trace rc=$?"
tail -3 gpurun_out/b8_trace_persist_S4096.txt; grep -n "O_ready\|epi_done\|HOISTED\|S(0)_issued" gpurun_out/b8_trace_persist_S4096.txt | head -16
