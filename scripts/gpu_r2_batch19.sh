#!/bin/bash
# final library after the kernel-choice refinement: GPU suite (AUTO), smoke, sweep, bench
mkdir -p gpurun_out
export TFA_NO_BUILD=1
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/b19_gpu_tests.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/b19_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/b19_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/b19_smoke.log
timeout 600 python scripts/sweep.py > gpurun_out/r02_sweep.md 2> gpurun_out/b19_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r02_sweep.md
timeout 900 python bench.py > gpurun_out/b19_bench.json 2> gpurun_out/b19_bench.err; echo "bench rc=$?"; cut -c1-600 gpurun_out/b19_bench.json
