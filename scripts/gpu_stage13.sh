#!/bin/bash
mkdir -p gpurun_out
TFA_KERNEL=split timeout 900 python -m pytest tests/test_fwd_parity.py tests/test_fwd_properties.py tests/test_drop_in_driver.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/gpu_tests_split.log 2>&1
echo "gpu_tests(split) rc=$?"; tail -4 gpurun_out/gpu_tests_split.log
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false],[4,32,4096,64,true],[8,32,4096,128,true]]'
fmt() { grep QT | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l[3:]); print('   B%d H%d S%d D%d %s: %.3f ms  %.0f TFLOPs(std)  %.1f%%' % (r['B'],r['H'],r['S'],r['D'],'causal' if r['causal'] else 'full  ',r['ms_med'],r['tflops_std'],100*r['frac_std_of_peak']))"; }
for rep in 1 2; do
  echo "== default (rep $rep)"; timeout 300 python scripts/quick_time.py "$CFG" 2>&1 | fmt
  echo "== split (rep $rep)"; TFA_KERNEL=split timeout 300 python scripts/quick_time.py "$CFG" 2>&1 | fmt
done
