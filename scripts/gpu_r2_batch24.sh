#!/bin/bash
# issuer/softmax barrier polling (test_wait) and two hand-offs: A/B on the persistent kernel with and without column split
mkdir -p gpurun_out; rm -f gpurun_out/ab.log
export TFA_NO_BUILD=1
CFG='[[4,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false]]' \
  timeout 1200 bash scripts/gpu_ab_env.sh "persist|TFA_KERNEL=persist|" "persist-spini|TFA_KERNEL=persist|libtfa_b200_spini.so" "persist-spinb|TFA_KERNEL=persist|libtfa_b200_spinb.so" "persist-two|TFA_KERNEL=persist|libtfa_b200_two.so" "persist-spin2|TFA_KERNEL=persist|libtfa_b200_spin2.so" \
  "split|TFA_KERNEL=split|" "split-spini|TFA_KERNEL=split|libtfa_b200_spini.so" "split-spinb|TFA_KERNEL=split|libtfa_b200_spinb.so" "split-two|TFA_KERNEL=split|libtfa_b200_two.so" "split-spin2|TFA_KERNEL=split|libtfa_b200_spin2.so" > gpurun_out/b24_ab.log 2>&1; echo "ab rc=$?"; cat gpurun_out/b24_ab.log
