#!/bin/bash
# Round-2 GPU batch 14: ncu captures of the persistent kernel (cfg5 shard), launch list of bench.py, sanitizer on the final persist, smoke
mkdir -p gpurun_out
export TFA_NO_BUILD=1
timeout 300 python __graft_entry__.py smoke > gpurun_out/b14_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/b14_smoke.log
TFA_KERNEL=persist timeout 600 ncu --set full --clock-control none --import-source on -k regex:persist -s 3 -c 1 -f -o gpurun_out/b14_prof_persist_cfg5shard python scripts/quick_time.py '[[8,32,4096,128,true]]' > gpurun_out/b14_ncu_full.log 2>&1; echo "ncu_full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/b14_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/b14_bench_under_ncu.log 2>&1; echo "ncu_launches rc=$?"
for tool in memcheck racecheck synccheck; do
  TFA_KERNEL=persist timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/b14_sanitizer_persist_$tool.log 2>&1; echo "sanitizer(persist) $tool rc=$?"; tail -3 gpurun_out/b14_sanitizer_persist_$tool.log
done
