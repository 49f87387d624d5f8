#!/bin/bash
# Round-2 GPU batch 4: why is the persistent issuer slow?  one-item-per-CTA / wait-hint experiments + ncu source-level captures.
mkdir -p gpurun_out
export TFA_NO_BUILD=1
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false]]' \
  timeout 600 bash scripts/gpu_ab_env.sh "default||" "persist|TFA_KERNEL=persist|" "oneitem|TFA_KERNEL=persist|libtfa_b200_oneitem.so" "waithint|TFA_KERNEL=persist|libtfa_b200_waithint.so" > gpurun_out/b4_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b4_ab.log
TFA_KERNEL=persist timeout 400 ncu --set full --clock-control none --import-source on -k regex:persist -s 3 -c 1 -f -o gpurun_out/b4_prof_persist python scripts/quick_time.py '[[4,32,4096,128,true]]' > gpurun_out/b4_ncu_persist.log 2>&1; echo "ncu persist rc=$?"
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -f -o gpurun_out/b4_prof_default python scripts/quick_time.py '[[4,32,4096,128,true]]' > gpurun_out/b4_ncu_default.log 2>&1; echo "ncu default rc=$?"
ls -la gpurun_out/*.ncu-rep
