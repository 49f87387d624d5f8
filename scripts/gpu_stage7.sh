#!/bin/bash
mkdir -p gpurun_out
run1() { timeout 300 python scripts/one_case.py "$1" 2>&1 | grep "ONE_CASE" | tee -a gpurun_out/cases.log; }
run1 '{"B":1,"H":2,"S":128,"D":64,"causal":false,"kind":"bf16","out_fp32":true}'
run1 '{"B":1,"H":2,"S":512,"D":128,"causal":true,"kind":"bf16","out_fp32":false}'
run1 '{"B":2,"H":40,"S":1024,"D":64,"causal":true,"kind":"fp16","out_fp32":false}'
run1 '{"B":3,"H":50,"S":1280,"D":128,"causal":true,"kind":"bf16","out_fp32":true}'
run1 '{"B":1,"H":3,"S":200,"D":64,"causal":true,"kind":"bf16","out_fp32":true}'
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu_tests rc=$?" | tee -a gpurun_out/summary.txt
tail -6 gpurun_out/gpu_tests.log
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false],[4,32,4096,64,true],[8,32,4096,128,true],[64,32,4096,128,true]]'
timeout 300 python scripts/quick_time.py "$CFG" 2>&1 | grep QT | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l[3:]); print('   B%d H%d S%d D%d %s: %.3f ms  %.0f TFLOPs(std)  %.1f%%' % (r['B'],r['H'],r['S'],r['D'],'causal' if r['causal'] else 'full  ',r['ms_med'],r['tflops_std'],100*r['frac_std_of_peak']))" | tee -a gpurun_out/variants.log
export TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so
timeout 300 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":10,"limit":511}' > gpurun_out/trace_causal.log 2>&1
unset TFA_LIB
grep -n "epi_done\|O_ready\|start" gpurun_out/trace_causal.log | head -30
