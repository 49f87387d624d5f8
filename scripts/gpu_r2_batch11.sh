#!/bin/bash
# Round-2 GPU batch 9: chunked masking (both kernels) + late draw of the next item (persist)
mkdir -p gpurun_out
export TFA_NO_BUILD=1
timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py tests/test_fwd_properties.py tests/test_lazy_rescale.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b11_tests_default.log 2>&1; echo "tests(default) rc=$?"; tail -3 gpurun_out/b11_tests_default.log | cut -c1-200
TFA_KERNEL=persist timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py tests/test_fwd_properties.py tests/test_lazy_rescale.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b11_tests_persist.log 2>&1; echo "tests(persist) rc=$?"; tail -3 gpurun_out/b11_tests_persist.log | cut -c1-200
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true],[4,16,2048,64,false],[64,32,4096,128,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "classic-old||libtfa_b200_noqpf.so" "classic||" "persist|TFA_KERNEL=persist|" > gpurun_out/b11_ab.log 2>&1; echo "ab rc=$?"; head -40 gpurun_out/b11_ab.log
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":5,"limit":10}' > gpurun_out/b11_trace_persist_S4096.txt 2>&1; echo "trace rc=$?"
tail -3 gpurun_out/b11_trace_persist_S4096.txt; grep -n "O_ready\|epi_done\|HOISTED\|S(0)_issued" gpurun_out/b11_trace_persist_S4096.txt | head -16
