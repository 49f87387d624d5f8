#!/bin/bash
mkdir -p gpurun_out
export TFA_NO_BUILD=1
timeout 900 python scripts/sweep.py > gpurun_out/r02_sweep.md 2> gpurun_out/r02_sweep.err; echo "sweep rc=$?"; cat gpurun_out/r02_sweep.md
