#!/bin/bash
# One GPU session: new-feature tests first (fail fast), then an A/B of library variants on the same box.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_general_attn.py tests/test_fwd_parity.py tests/test_fwd_properties.py -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/gpu_tests_new.log 2>&1; echo "new tests rc=$?"; tail -15 gpurun_out/gpu_tests_new.log
rm -f gpurun_out/ab.log
CFG='[[4,32,4096,128,true],[8,32,4096,128,true],[1,32,16384,128,true],[4,32,4096,128,false],[4,16,2048,64,false]]' bash scripts/gpu_ab.sh "$@"
