#!/bin/bash
# First bring-up trip: primitives -> (sweeps if needed) -> parity -> quick timing. Logs under gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests/test_umma_primitives.py -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/prim.log 2>&1
echo "prim rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_umma_primitives.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/prim_all.log 2>&1
echo "prim_all rc=$?" | tee -a gpurun_out/summary.txt
if ! grep -q " passed" gpurun_out/prim_all.log || grep -q "failed" gpurun_out/prim_all.log; then
  for m in 0 1 2; do
    timeout 600 python tests/prim_runner.py sweep "{\"N\":128,\"K\":128,\"mode\":$m}" >> gpurun_out/sweep.log 2>&1
  done
fi
timeout 1200 python -m pytest tests/test_fwd_parity.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/parity.log 2>&1
echo "parity rc=$?" | tee -a gpurun_out/summary.txt
timeout 600 python scripts/quick_time.py > gpurun_out/quick_time.log 2>&1
echo "quick_time rc=$?" | tee -a gpurun_out/summary.txt
tail -5 gpurun_out/prim_all.log; tail -15 gpurun_out/parity.log; cat gpurun_out/quick_time.log | tail -10
