#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fwd_parity.py tests/test_drop_in_driver.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu_tests rc=$?"; tail -2 gpurun_out/gpu_tests.log
bash scripts/gpu_ab.sh libtfa_b200_v2plain.so libtfa_b200.so 2>&1 | tail -30
TFA_LIB=$PWD/tiny-flash-attention_b200/libtfa_b200_trace.so timeout 300 python scripts/trace_run.py '{"B":1,"H":32,"S":16384,"D":128,"causal":true,"block":0,"limit":511}' > gpurun_out/trace_cfg4_trace.log 2>&1
