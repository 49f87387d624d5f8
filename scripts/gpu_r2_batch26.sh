#!/bin/bash
# where the first of two P hand-offs sits (persistent kernel) and two-stage hand-offs in the one-CTA-per-item kernel
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
export TFA_NO_BUILD=1
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "persist|TFA_KERNEL=persist|" "persist-first96|TFA_KERNEL=persist|libtfa_b200_pq2.so" "persist-first32|TFA_KERNEL=persist|libtfa_b200_pq0.so" \
  "classic|TFA_KERNEL=classic|" "classic-2stage-64|TFA_KERNEL=classic|libtfa_b200_cl22.so" "classic-2stage-96|TFA_KERNEL=classic|libtfa_b200_cl23.so" > gpurun_out/b26_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b26_ab.log
