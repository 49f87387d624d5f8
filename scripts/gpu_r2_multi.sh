#!/bin/bash
# Round-2 multi-GPU batch (gpurun --gpus N): fused-exchange tests against the oracle + bench with the fixed parity check.
# usage: gpu_r2_multi.sh N "variant1 variant2 ..."   (variant = value of TFA_KERNEL, "default" = unset)
N=${1:-2}; VARIANTS=${2:-"default persist"}
mkdir -p gpurun_out
export TFA_NO_BUILD=1
nvidia-smi topo -m > gpurun_out/m${N}_topo.txt 2>&1
for v in $VARIANTS; do
  [ "$v" = "default" ] && unset TFA_KERNEL || export TFA_KERNEL=$v
  timeout 600 python -m pytest tests/test_multi_gpu.py tests/test_multi_device.py -m gpu -q --no-header -p no:cacheprovider > gpurun_out/m${N}_test_$v.log 2>&1
  echo "multi-gpu tests [$v] rc=$?"; tail -4 gpurun_out/m${N}_test_$v.log
  for ex in fused ${NCCL_ARM:-nccl}; do
    [ "$ex" = "none" ] && continue; [ "$ex" = "nccl" ] && [ "$v" != "default" ] && continue
    timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus $N --steps 10 --warmup 3 --exchange $ex > gpurun_out/m${N}_bench_${v}_$ex.json 2> gpurun_out/m${N}_bench_${v}_$ex.err
    echo "bench N=$N [$v] $ex rc=$?"; python - <<PY
import json
try:
    d=json.load(open("gpurun_out/m${N}_bench_${v}_$ex.json"))
    print("  value %.1f TFLOP/s  ms_per_step %.3f  kernel_ms_mean %.3f  compute_only %.1f  e2e %s" % (d["value"], d["ms_per_step"], d["roofline"]["kernel_ms_mean"], d["compute_only"]["value"], (d.get("e2e") or {}).get("value")))
    print("  parity", d["parity"]["pass_frac_rtol1e-3_atol1e-3"], d["parity"]["max_abs_err"], "| e2e matches:", (d.get("e2e") or {}).get("matches_device_path"), "| exchange_check:", d["config"].get("exchange_check"))
    print("  fused_roofline", d.get("fused_roofline"))
    print("  numa", (d.get("e2e") or {}).get("numa"))
except Exception as e:
    print("  parse failed", e); print(open("gpurun_out/m${N}_bench_${v}_$ex.err").read()[-1500:])
PY
  done
done
