#!/bin/bash
# Round-2 final verification (after the two-hand-off change and the kernel-choice refinement): final script + fresh ncu captures
bash scripts/gpu_r2_final.sh
export TFA_NO_BUILD=1
TFA_KERNEL=persist timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist -s 3 -c 1 -f -o gpurun_out/f2_prof_persist_cfg5shard python scripts/quick_time.py '[[8,32,4096,128,true]]' > gpurun_out/f2_ncu_full.log 2>&1; echo "ncu_full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/f2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/f2_bench_under_ncu.log 2>&1; echo "ncu_launches rc=$?"
ls -la gpurun_out/f2_*
