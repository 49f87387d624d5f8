#!/bin/bash
# Round-2 GPU batch 13: full GPU suite with the AUTO kernel choice, smoke, both bench arms, D=64 A/B with repeats
mkdir -p gpurun_out
export TFA_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/b13_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -4 gpurun_out/b13_gpu_tests.log | cut -c1-200
timeout 300 python __graft_entry__.py smoke > gpurun_out/b13_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/b13_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/b13_bench_ref.json 2> gpurun_out/b13_bench_ref.err; echo "bench ref rc=$?"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/b13_bench_n1.json 2> gpurun_out/b13_bench_n1.err; echo "bench rc=$?"; tail -5 gpurun_out/b13_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/b13_bench_n1.json"))
print("value %.1f  roofline.frac %.3f (%s)  e2e %.1f  cpu %.4f  clocks %s" % (d["value"], d["roofline"]["frac"], d["roofline"]["kernel"][:40], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"]))
print("parity", d["parity"]["pass_frac_rtol1e-3_atol1e-3"], d["parity"]["max_abs_err"])
for k,v in d["configs"].items(): print("  ", k, v.get("kernel"), "%.4f ms %.0f TFLOP/s frac %.3f (std %.0f / %.3f)" % (v["ms"], v["tflops"], v["roofline_frac"], v["tflops_std"], v["roofline_frac_std"]))
print(json.dumps(d.get("comparators"))[:3000])
print(json.dumps(d["cpu_baseline"])[:1500])
PY
for rep in 1 2 3; do
CFG='[[4,16,2048,64,false],[4,32,4096,64,true],[16,16,1024,64,false],[2,16,8192,64,true]]' timeout 300 bash scripts/gpu_ab_env.sh "classic|TFA_KERNEL=classic|" "persist|TFA_KERNEL=persist|" "persist64|TFA_KERNEL=persist64|" > gpurun_out/b13_ab_d64_$rep.log 2>&1
done; grep -h "D64" -B0 gpurun_out/b13_ab_d64_*.log | sort | uniq -c | sort -k3 | head -60
