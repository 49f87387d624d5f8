#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_fwd_parity.py tests/test_fwd_properties.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/gpu_tests.log 2>&1
echo "gpu_tests rc=$?"; tail -2 gpurun_out/gpu_tests.log
bash scripts/gpu_ab.sh libtfa_b200_v2plain.so libtfa_b200.so 2>&1 | tail -28
