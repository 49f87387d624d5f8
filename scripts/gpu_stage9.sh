#!/bin/bash
mkdir -p gpurun_out
for v in v2trace trace; do
  TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_$v.so timeout 300 python scripts/trace_run.py '{"B":1,"H":32,"S":16384,"D":128,"causal":true,"block":0,"limit":511}' > gpurun_out/trace_cfg4_$v.log 2>&1
done
CFG='[[1,32,16384,128,true],[4,32,4096,128,true]]' bash scripts/gpu_ab.sh libtfa_b200_v2plain.so libtfa_b200.so libtfa_b200_r208.so 2>&1 | tail -30
