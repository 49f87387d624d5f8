#!/bin/bash
# usage: gpu_multi.sh N   (run under gpurun --gpus N)
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L | head -8
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 10 --warmup 3 --no-e2e > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N rc=$?"; cat gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 5 --warmup 3 --chunks 1 --no-e2e > gpurun_out/bench_n${N}_chunks1.json 2>> gpurun_out/bench_n$N.err
echo "bench N=$N chunks=1 rc=$?"; cut -c1-300 gpurun_out/bench_n${N}_chunks1.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $N --steps 3 --warmup 3 > gpurun_out/bench_n${N}_e2e.json 2>> gpurun_out/bench_n$N.err
echo "bench N=$N with e2e rc=$?"; python -c "
import json; d=json.load(open('gpurun_out/bench_n${N}_e2e.json')); print('value',d['value'],'e2e',d['e2e'])"
