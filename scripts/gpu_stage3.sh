#!/bin/bash
# Full GPU test-suite, bench line, ncu launch list + one full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1
echo "gpu_tests rc=$?" | tee -a gpurun_out/summary.txt
tail -15 gpurun_out/gpu_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err
echo "bench rc=$?" | tee -a gpurun_out/summary.txt
cat gpurun_out/bench_n1.json; tail -5 gpurun_out/bench_n1.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -f -o gpurun_out/prof_cfg3 python scripts/quick_time.py '[[4,32,4096,128,true]]' > gpurun_out/ncu_full.log 2>&1
echo "ncu_full rc=$?" | tee -a gpurun_out/summary.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1
echo "ncu_launches rc=$?" | tee -a gpurun_out/summary.txt
ls -la gpurun_out
