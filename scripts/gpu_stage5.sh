#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/segv_repro.py > gpurun_out/segv.log 2>&1
echo "segv_repro rc=$?" | tee -a gpurun_out/summary.txt
if ! grep -q "^done" gpurun_out/segv.log; then
  timeout 600 cuda-gdb -batch -ex run -ex bt --args python scripts/segv_repro.py > gpurun_out/segv_gdb.log 2>&1
  grep -A30 "SIGSEGV" gpurun_out/segv_gdb.log | head -50
fi
tail -5 gpurun_out/segv.log
export TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so
timeout 300 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":false,"block":300,"limit":260}' > gpurun_out/trace_noncausal.log 2>&1
timeout 300 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":304,"limit":400}' > gpurun_out/trace_causal.log 2>&1
timeout 300 python scripts/trace_run.py '{"B":4,"H":16,"S":2048,"D":64,"causal":false,"block":100,"limit":300}' > gpurun_out/trace_d64.log 2>&1
unset TFA_LIB
sed -n 1,140p gpurun_out/trace_noncausal.log
