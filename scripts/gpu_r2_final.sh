#!/bin/bash
# Round-2 final verification on one B200: full GPU suite (AUTO kernel choice), each kernel forced on the parity files, smoke, both bench arms
mkdir -p gpurun_out
export TFA_NO_BUILD=1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/f_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -3 gpurun_out/f_gpu_tests.log | cut -c1-200
for v in classic persist; do
TFA_KERNEL=$v timeout 600 python -m pytest tests/test_fwd_parity.py tests/test_general_attn.py tests/test_fused_exchange.py tests/test_fwd_properties.py tests/test_lazy_rescale.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/f_tests_$v.log 2>&1; echo "tests($v) rc=$?"; tail -2 gpurun_out/f_tests_$v.log | cut -c1-200
done
timeout 300 python __graft_entry__.py smoke > gpurun_out/f_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/f_smoke.log
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/f_bench_ref.json 2> gpurun_out/f_bench_ref.err; echo "bench ref rc=$?"
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/f_bench_n1.json 2> gpurun_out/f_bench_n1.err; echo "bench rc=$?"; tail -3 gpurun_out/f_bench_n1.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/f_bench_n1.json"))
print("value %.1f  roofline.frac %.3f (%s)  traffic %s  e2e %.1f  cpu %.4f  clocks %s" % (d["value"], d["roofline"]["frac"], d["roofline"]["kernel"][:40], d["roofline"]["traffic"], d["e2e"]["value"], d["cpu_baseline"]["value"], d["clocks"]))
print("parity", d["parity"]["pass_frac_rtol1e-3_atol1e-3"], d["parity"]["max_abs_err"], "kernel ms mean/min", d["roofline"]["kernel_ms_mean"], d["roofline"]["kernel_ms_min"])
for k,v in d["configs"].items(): print("  ", k, v.get("kernel"), "%.4f ms %.0f TFLOP/s frac %.3f (std %.0f / %.3f)" % (v["ms"], v["tflops"], v["roofline_frac"], v["tflops_std"], v["roofline_frac_std"]))
for cfg,r in (d.get("comparators") or {}).items():
    print("  cmp", cfg, {k:(round(v["ms"],4) if "ms" in v else v) for k,v in r.items()})
PY
