#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu_tests rc=$?" | tee -a gpurun_out/summary.txt
tail -6 gpurun_out/gpu_tests.log
timeout 600 python scripts/microbench.py > gpurun_out/microbench.log 2>&1
cat gpurun_out/microbench.log | grep -v "^MICRO"
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false],[4,32,4096,64,true],[8,32,4096,128,true]]'
for v in "" _emu0 _emu2 _emu3; do
  echo "== variant ${v:-default}" | tee -a gpurun_out/variants.log
  TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200$v.so timeout 300 python scripts/quick_time.py "$CFG" 2>&1 | grep QT | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l[3:]); print('   B%d H%d S%d D%d %s: %.3f ms  %.0f TFLOPs(std)  %.1f%%' % (r['B'],r['H'],r['S'],r['D'],'causal' if r['causal'] else 'full  ',r['ms_med'],r['tflops_std'],100*r['frac_std_of_peak']))" | tee -a gpurun_out/variants.log
done
export TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so
timeout 300 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":false,"block":300,"limit":330}' > gpurun_out/trace_noncausal.log 2>&1
timeout 300 python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":304,"limit":500}' > gpurun_out/trace_causal.log 2>&1
timeout 300 python scripts/trace_run.py '{"B":4,"H":16,"S":2048,"D":64,"causal":false,"block":100,"limit":300}' > gpurun_out/trace_d64.log 2>&1
unset TFA_LIB
sed -n 150,250p gpurun_out/trace_noncausal.log
