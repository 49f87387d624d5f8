#!/bin/bash
# trace of one CTA + full bench line + ncu of the split-KV pair of kernels
mkdir -p gpurun_out
TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 200 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":150,"limit":260}' > gpurun_out/trace_cfg3.txt 2>&1; echo "trace rc=$?"; tail -3 gpurun_out/trace_cfg3.txt
TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 200 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":false,"block":150,"limit":200}' > gpurun_out/trace_noncausal.txt 2>&1; tail -3 gpurun_out/trace_noncausal.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench_n1.json 2> gpurun_out/bench_n1.err; echo "bench rc=$?"; python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_n1.json"))
print("value %.1f  roofline.frac %.3f  e2e %.1f  clocks %s" % (d["value"], d["roofline"]["frac"], d["e2e"]["value"], d["clocks"]))
for k,v in d["configs"].items(): print("  ", k, "%.3f ms %.0f TFLOP/s frac %.3f (std %.0f / %.3f)" % (v["ms"], v["tflops"], v["roofline_frac"], v["tflops_std"], v["roofline_frac_std"]))
print(json.dumps(d["next_rows"], indent=1))
PY
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'splitkv_combine|fa_fwd_sm100' -s 6 -c 2 -f -o gpurun_out/prof_splitkv python scripts/splitkv_case.py 0 > gpurun_out/ncu_splitkv.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_splitkv.log
