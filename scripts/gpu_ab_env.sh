#!/bin/bash
# A/B of (library, env) combinations in ONE session.  Each argument: "label|ENV=VAL ENV2=VAL|lib.so" (lib optional)
mkdir -p gpurun_out
CFG=${CFG:-'[[4,32,4096,128,true],[8,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false],[16,32,1024,128,true]]'}
for rep in 1 2; do
for spec in "$@"; do
  label=${spec%%|*}; rest=${spec#*|}; envs=${rest%%|*}; lib=${rest#*|}
  [ "$lib" = "$rest" ] && lib=""
  echo "== $label (rep $rep)" | tee -a gpurun_out/ab.log
  ( [ -n "$lib" ] && export TFA_LIB=$PWD/tiny-flash-attention_b200/$lib; for e in $envs; do export $e; done
    timeout 300 python scripts/quick_time.py "$CFG" 2>&1 | grep QT | python -c "
import sys,json
for l in sys.stdin:
    r=json.loads(l[3:]); print('   B%d H%d S%d D%d %s: %.4f ms  %.0f TFLOPs(std)  %.1f%%' % (r['B'],r['H'],r['S'],r['D'],'causal' if r['causal'] else 'full  ',r['ms_med'],r['tflops_std'],100*r['frac_std_of_peak']))" ) | tee -a gpurun_out/ab.log
done
done
