#!/bin/bash
# compute-sanitizer on the FINAL persistent kernel (two hand-offs): memcheck, racecheck, synccheck over scripts/sanitize_cases.py
mkdir -p gpurun_out
export TFA_NO_BUILD=1
for tool in memcheck racecheck synccheck; do
  TFA_KERNEL=persist timeout 150 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/b27_sanitizer_persist_$tool.log 2>&1; echo "sanitizer(persist) $tool rc=$?"; tail -3 gpurun_out/b27_sanitizer_persist_$tool.log
done
