#!/bin/bash
# in-kernel timelines of the persistent kernel with and without the column-split softmax
mkdir -p gpurun_out
export TFA_NO_BUILD=1
for k in persist split; do
TFA_KERNEL=$k TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 120 python scripts/trace_run.py '{"B":1,"H":32,"S":16384,"D":128,"causal":true,"block":5,"limit":330}' > gpurun_out/b23_trace_${k}_S16384.txt 2>&1; echo "trace $k rc=$?"
tail -3 gpurun_out/b23_trace_${k}_S16384.txt
done
