#!/bin/bash
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_multi_gpu.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/multi_test.log 2>&1
echo "multi-gpu test rc=$?"; tail -15 gpurun_out/multi_test.log
for ex in fused nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e --exchange $ex > gpurun_out/bench_n${N}_$ex.json 2> gpurun_out/bench_n${N}_$ex.err
echo "bench N=$N $ex rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/bench_n${N}_$ex.json"))
    print("  value %.1f TFLOP/s  ms_per_step %.3f  exchange=%s check=%s  kernel_ms_mean %.3f  compute_only %.1f" % (d["value"], d["ms_per_step"], d["config"]["exchange"], d["config"].get("exchange_check"), d["roofline"]["kernel_ms_mean"], d["compute_only"]["value"]))
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/bench_n${N}_$ex.err").read()[-1500:])
PY
done
