#!/bin/bash
# Round-2 GPU batch 1: full GPU suite, reference drivers, sanitizer, variant A/B, comparators, D=64 ncu capture.
mkdir -p gpurun_out
export TFA_NO_BUILD=1   # use the libraries that travelled with the snapshot
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/b1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/b1_gpu_tests.log 2>&1; echo "gpu_tests rc=$?"; tail -3 gpurun_out/b1_gpu_tests.log
TFA_KERNEL=persist timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/b1_gpu_tests_persist.log 2>&1; echo "gpu_tests(persist) rc=$?"; tail -3 gpurun_out/b1_gpu_tests_persist.log
TFA_KERNEL=persist64 timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/b1_gpu_tests_persist64.log 2>&1; echo "gpu_tests(persist64) rc=$?"; tail -3 gpurun_out/b1_gpu_tests_persist64.log
# variants A/B (existing round-1 experimental kernels): run them or delete them
CFG='[[4,32,4096,128,true],[8,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false],[4,32,4096,64,true],[16,16,1024,64,false]]' \
  timeout 600 bash scripts/gpu_ab_env.sh "default||" "persist|TFA_KERNEL=persist|" "persist64|TFA_KERNEL=persist64|" "persistent|TFA_KERNEL=persistent|" "persistent2|TFA_KERNEL=persistent2|" "colsplit|TFA_KERNEL=colsplit|" > gpurun_out/b1_ab.log 2>&1; echo "ab rc=$?"
for v in persistent2 colsplit; do
  TFA_KERNEL=$v timeout 300 python -m pytest tests/test_fwd_parity.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/b1_parity_$v.log 2>&1; echo "parity $v rc=$?"; tail -2 gpurun_out/b1_parity_$v.log
done
timeout 420 python scripts/comparators.py > gpurun_out/b1_comparators.log 2>&1; echo "comparators rc=$?"; grep CMP gpurun_out/b1_comparators.log | head -c 3000
for tool in memcheck racecheck synccheck; do
  timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/b1_sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -4 gpurun_out/b1_sanitizer_$tool.log
  TFA_KERNEL=persist timeout 500 compute-sanitizer --tool $tool --print-limit 20 python scripts/sanitize_cases.py > gpurun_out/b1_sanitizer_persist_$tool.log 2>&1; echo "sanitizer(persist) $tool rc=$?"; tail -4 gpurun_out/b1_sanitizer_persist_$tool.log
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:fa_fwd_sm100 -s 3 -c 1 -f -o gpurun_out/b1_prof_cfg2_d64 python scripts/quick_time.py '[[4,16,2048,64,false]]' > gpurun_out/b1_ncu_d64.log 2>&1; echo "ncu d64 rc=$?"
ls -la gpurun_out | head -50
