#!/bin/bash
# Round-end style verification on one B200: GPU tests, smoke, both bench arms, ncu captures -> gpurun_out/
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -2 gpurun_out/gpu_tests.log
TFA_KERNEL=persistent timeout 900 python -m pytest tests/test_fwd_parity.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/gpu_tests_persistent.log 2>&1; echo "persistent parity rc=$?"
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "bench ref rc=$?"
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print("value %.1f  roofline.frac %.3f  e2e %.1f  cpu %.4f  clocks %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"]))
for k,v in d["configs"].items(): print("  ", k, "%.3f ms %.0f TFLOP/s frac %.3f (std %.0f / %.3f)" % (v["ms"], v["tflops"], v["roofline_frac"], v["tflops_std"], v["roofline_frac_std"]))
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -f -o gpurun_out/prof_cfg5shard python scripts/quick_time.py '[[8,32,4096,128,true]]' > gpurun_out/ncu_full.log 2>&1; echo "ncu_full rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras > gpurun_out/bench_under_ncu.log 2>&1; echo "ncu_launches rc=$?"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'splitkv_combine|fa_fwd_sm100' -s 6 -c 2 -f -o gpurun_out/prof_splitkv python scripts/splitkv_case.py 0 > gpurun_out/ncu_splitkv.log 2>&1; echo "ncu_splitkv rc=$?"
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print(json.dumps(d.get("next_rows"), indent=1))
PY
