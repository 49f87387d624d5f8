"""Minimal on-GPU check of the launch-order change: a few shapes against torch SDPA (fp32 math), then cfg3/cfg4 times."""
import math, os, sys, time
t0 = time.time()
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
import tfa_ctypes
print("import %.1fs" % (time.time() - t0), flush=True)
for (B, H, S, D, causal) in ((2, 5, 700, 128, True), (1, 3, 300, 64, False), (3, 11, 1024, 128, True)):
    q, k, v = (torch.empty(B, H, S, D, dtype=torch.bfloat16, device="cuda").normal_(0, 0.5) for _ in range(3))
    o, lse = tfa_ctypes.fwd(q, k, v, causal, 1 / math.sqrt(D))
    ref = torch.nn.functional.scaled_dot_product_attention(q.float(), k.float(), v.float(), is_causal=causal)
    print("CHECK", (B, H, S, D, causal), "max err %.2e" % (o.float() - ref).abs().max().item(), flush=True)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
for (B, H, S, D) in ((4, 32, 4096, 128), (1, 32, 16384, 128), (8, 32, 4096, 128)):
    q, k, v = (torch.empty(B, H, S, D, dtype=torch.bfloat16, device="cuda").normal_(0, 0.5) for _ in range(3))
    out = torch.empty_like(q); lse = torch.empty(B, H, S, dtype=torch.float32, device="cuda")
    for _ in range(3): tfa_ctypes.fwd(q, k, v, True, 1 / math.sqrt(D), out=out, lse=lse)
    ts = []
    for _ in range(8):
        flush.zero_(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); tfa_ctypes.fwd(q, k, v, True, 1 / math.sqrt(D), out=out, lse=lse); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    print("TIME", (B, H, S, D), "%.4f ms  %.0f TFLOP/s" % (ts[4], 2.0 * B * H * S * S * D / (ts[4] * 1e-3) / 1e12), flush=True)
