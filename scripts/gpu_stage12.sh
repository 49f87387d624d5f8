#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"; cut -c1-400 gpurun_out/bench_ref.json
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; cat gpurun_out/bench_n1.json; tail -3 gpurun_out/bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -f -o gpurun_out/prof_cfg5shard python scripts/quick_time.py '[[8,32,4096,128,true]]' > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu_launches rc=$?"
