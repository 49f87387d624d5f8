#!/bin/bash
# Forward-kernel bring-up: isolated single cases (each its own process), one under compute-sanitizer, then parity + timing.
mkdir -p gpurun_out
run1() { timeout 300 python scripts/one_case.py "$1" 2>&1 | grep "ONE_CASE" | tee -a gpurun_out/cases.log; }
run1 '{"B":1,"H":2,"S":128,"D":64,"causal":false,"kind":"bf16","out_fp32":true}'
run1 '{"B":1,"H":1,"S":256,"D":64,"causal":false,"kind":"bf16","out_fp32":true}'
run1 '{"B":1,"H":1,"S":256,"D":128,"causal":false,"kind":"bf16","out_fp32":true}'
run1 '{"B":1,"H":2,"S":512,"D":128,"causal":true,"kind":"bf16","out_fp32":false}'
run1 '{"B":2,"H":4,"S":1024,"D":64,"causal":true,"kind":"fp16","out_fp32":false}'
run1 '{"B":1,"H":2,"S":200,"D":64,"causal":true,"kind":"bf16","out_fp32":true}'
if grep -q '"ok": false' gpurun_out/cases.log; then
  timeout 600 /usr/local/cuda/bin/compute-sanitizer --tool memcheck --print-limit 20 python scripts/one_case.py '{"B":1,"H":1,"S":256,"D":64,"causal":false,"kind":"bf16","out_fp32":true}' > gpurun_out/sanitizer.log 2>&1
  tail -40 gpurun_out/sanitizer.log
fi
timeout 1200 python -m pytest tests/test_fwd_parity.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/summary.txt
tail -12 gpurun_out/parity.log
timeout 600 python scripts/quick_time.py > gpurun_out/quick_time.log 2>&1
echo "quick_time rc=$?" | tee -a gpurun_out/summary.txt
tail -10 gpurun_out/quick_time.log
