#!/bin/bash
mkdir -p gpurun_out
export TFA_NO_BUILD=1
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[8,32,4096,128,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "classic||" "persist|TFA_KERNEL=persist|" "persist-unrollt|TFA_KERNEL=persist|libtfa_b200_unrollt.so" > gpurun_out/b12_ab.log 2>&1; echo "ab rc=$?"; head -30 gpurun_out/b12_ab.log
TFA_KERNEL=persist TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py '{"B":1,"H":32,"S":16384,"D":128,"causal":true,"block":5,"limit":10}' > gpurun_out/b12_trace_persist_S16384.txt 2>&1; echo "trace rc=$?"
tail -3 gpurun_out/b12_trace_persist_S16384.txt
TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py '{"B":1,"H":32,"S":16384,"D":128,"causal":true,"block":5,"limit":10}' > gpurun_out/b12_trace_classic_S16384.txt 2>&1; echo "trace rc=$?"
tail -3 gpurun_out/b12_trace_classic_S16384.txt
