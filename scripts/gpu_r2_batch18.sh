#!/bin/bash
mkdir -p gpurun_out
export TFA_NO_BUILD=1
CFG='[[2,32,8192,128,false],[1,32,16384,128,false],[2,32,8192,128,true],[8,32,2048,128,true],[16,32,1024,128,true],[32,32,512,128,true],[2,32,8192,64,false],[1,32,16384,64,false],[8,32,2048,64,true],[32,32,512,64,true]]' \
  timeout 900 bash scripts/gpu_ab_env.sh "classic|TFA_KERNEL=classic|" "persist|TFA_KERNEL=persist|" > gpurun_out/b18_ab.log 2>&1; echo "ab rc=$?"; head -46 gpurun_out/b18_ab.log
