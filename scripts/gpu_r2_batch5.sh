#!/bin/bash
# Round-2 GPU batch 5: bisecting the persist kernel's slow steady state (one-item mode variants) + smem-size control
mkdir -p gpurun_out
export TFA_NO_BUILD=1
CFG='[[4,32,4096,128,true],[1,32,16384,128,true]]' \
  timeout 600 bash scripts/gpu_ab_env.sh "default||" "default+32KB smem||libtfa_b200_pad.so" "oneitem|TFA_KERNEL=persist|libtfa_b200_oneitem.so" "o2 nostg|TFA_KERNEL=persist|libtfa_b200_o2.so" "o3 nostg nosched|TFA_KERNEL=persist|libtfa_b200_o3.so" > gpurun_out/b5_ab.log 2>&1; echo "ab rc=$?"; head -24 gpurun_out/b5_ab.log
