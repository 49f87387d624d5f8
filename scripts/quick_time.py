"""Quick kernel timing sweep (CUDA events, L2 flush between reps). Not the bench contract -- a dev tool."""
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
import tfa_ctypes  # noqa: E402

PEAK = 1709.7e12


def time_cfg(B, H, S, D, causal, dtype=torch.bfloat16, reps=10, warm=3):
    q, k, v = (torch.empty(B, H, S, D, dtype=dtype, device="cuda").normal_(0, 0.5) for _ in range(3))
    out = torch.empty_like(q)
    lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    scale = 1 / math.sqrt(D)
    for _ in range(warm):
        tfa_ctypes.fwd(q, k, v, causal, scale, out=out, lse=lse)
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tfa_ctypes.fwd(q, k, v, causal, scale, out=out, lse=lse)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    t = ts[len(ts) // 2]
    F = 2.0 * B * H * S * S * D
    Fstd = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
    return {"B": B, "H": H, "S": S, "D": D, "causal": causal, "ms_med": t * 1e3, "ms_min": ts[0] * 1e3,
            "tflops_F": F / t / 1e12, "tflops_std": Fstd / t / 1e12, "frac_std_of_peak": Fstd / t / PEAK}


if __name__ == "__main__":
    cfgs = [(4, 16, 2048, 64, False), (4, 32, 4096, 128, True), (1, 32, 16384, 128, True), (8, 32, 4096, 128, True),
            (4, 32, 4096, 128, False), (4, 32, 4096, 64, True), (2, 32, 8192, 128, False)]
    if len(sys.argv) > 1:
        cfgs = json.loads(sys.argv[1])
    for c in cfgs:
        r = time_cfg(*c)
        print("QT " + json.dumps(r), flush=True)
