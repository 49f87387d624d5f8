"""Headline sweep of SURVEY.md 8(d): D in {64,128} x S in {512..16384} x causal in {F,T}, B*H chosen so that
B*H*S = 2^19 tokens (>= 1 wave), bf16, scale 1/sqrt(D); CUDA events, median of 10, 256 MB L2 flush between reps.
Prints a markdown table (TFLOP/s with BASELINE.json's F = 2*B*H*S^2*D and with the standard count)."""
import math
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
sys.path.insert(0, ROOT)
import tfa_ctypes  # noqa: E402
from bench import ClockSampler  # noqa: E402  (NVML clock / throttle-reason sampling during the timed region)

PEAK = 1709.7


def run(B, H, S, D, causal, dtype=torch.bfloat16, reps=15):
    q, k, v = (torch.empty(B, H, S, D, dtype=dtype, device="cuda").normal_(0, 0.5) for _ in range(3))
    out = torch.empty_like(q)
    lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    sc = 1 / math.sqrt(D)
    for _ in range(3):
        tfa_ctypes.fwd(q, k, v, causal, sc, out=out, lse=lse)
    torch.cuda.synchronize()
    time.sleep(0.7)            # let the power limiter recover: back-to-back heavy configs otherwise bias the next one
    ts = []
    with ClockSampler() as clk:
      for _ in range(reps):
        FLUSH.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        tfa_ctypes.fwd(q, k, v, causal, sc, out=out, lse=lse)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e-3)
    ts.sort()
    c = clk.summary()
    return ts[len(ts) // 2], ts[0], ("%s MHz %s" % (c["sm_mhz"], ",".join(c["reasons"]) or "-"))


if __name__ == "__main__":
    torch.manual_seed(20)
    FLUSH = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
    print("| D | S | B*H | causal | ms (median) | ms (min) | TFLOP/s (F=2BHS^2D) | TFLOP/s (std count) | std / 1709.7 | clocks |")
    print("|---|---|---|---|---|---|---|---|---|---|")
    for D in (64, 128):
        for causal in (False, True):
            for S in (512, 1024, 2048, 4096, 8192, 16384):
                BH = (1 << 19) // S
                H = min(32, BH)
                B = BH // H
                t, tmin, clocks = run(B, H, S, D, causal)
                F = 2.0 * B * H * S * S * D
                Fstd = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
                print(f"| {D} | {S} | {B}x{H} | {'yes' if causal else 'no'} | {t * 1e3:.3f} | {tmin * 1e3:.3f} | "
                      f"{F / t / 1e12:.0f} | {Fstd / t / 1e12:.0f} | {Fstd / t / 1e12 / PEAK:.3f} | {clocks} |", flush=True)
    # fp16 spot checks of the two BASELINE shapes
    for (B, H, S, D, causal) in ((4, 32, 4096, 128, True), (4, 16, 2048, 64, False)):
        t, tmin, clocks = run(B, H, S, D, causal, dtype=torch.float16)
        Fstd = 4.0 * B * H * S * S * D * (0.5 if causal else 1.0)
        print(f"| {D} (fp16) | {S} | {B}x{H} | {'yes' if causal else 'no'} | {t * 1e3:.3f} | {tmin * 1e3:.3f} | "
              f"{2.0 * B * H * S * S * D / t / 1e12:.0f} | {Fstd / t / 1e12:.0f} | {Fstd / t / 1e12 / PEAK:.3f} | {clocks} |")
