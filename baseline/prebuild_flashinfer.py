"""Pre-build flashinfer's CUTLASS sm100a FMHA JIT module (a LIBRARY comparator, scripts/comparators.py) in the CPU
container so the GPU box does not spend GPU-minutes compiling it.  The JIT cache lives under
baseline/_ref/flashinfer_ws (git-ignored, travels with the snapshot; /root/repo is the same absolute path on the box).
Not product code."""
import os
import sys

WS = "/root/repo/baseline/_ref/flashinfer_ws"
os.makedirs(WS, exist_ok=True)
os.environ["FLASHINFER_WORKSPACE_BASE"] = WS
os.environ.setdefault("TORCH_CUDA_ARCH_LIST", "10.0a")
os.environ.setdefault("FLASHINFER_CUDA_ARCH_LIST", "10.0a")
os.environ.setdefault("MAX_JOBS", "6")

import torch  # noqa: E402
from flashinfer.prefill import gen_fmha_cutlass_sm100a_module  # noqa: E402

for hd in (128, 64):
    spec = gen_fmha_cutlass_sm100a_module(torch.bfloat16, torch.bfloat16, torch.bfloat16, torch.int32, hd, hd, 0, False,
                                          False)
    print("[flashinfer] building", spec.name, flush=True)
    try:
        spec.build(verbose="-v" in sys.argv)
        print("[flashinfer] ok:", spec.get_library_path() if hasattr(spec, "get_library_path") else "", flush=True)
    except Exception as e:  # noqa: BLE001
        print("[flashinfer] FAILED:", repr(e)[:2000], flush=True)
