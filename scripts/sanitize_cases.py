"""Small forward cases for compute-sanitizer (memcheck / racecheck / synccheck): causal, ragged, split-KV, GQA, the
fused exchange with local "peer" buffers.  Each case is checked against the oracle so that a tool-induced slowdown
cannot hide a wrong result.  Run:  compute-sanitizer --tool memcheck python scripts/sanitize_cases.py"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tfa_ctypes as tfa  # noqa: E402
from helpers import ref_inputs  # noqa: E402
from oracle import oracle as orc  # noqa: E402


def want(q, k, v, causal, scale):
    return orc.attn_general(q.float().cpu().numpy(), k.float().cpu().numpy(), v.float().cpu().numpy(), causal, scale,
                            orc.ROUND_BF16, False)[0]


def main():
    bad = 0
    for (B, H, S, D, causal) in [(1, 2, 256, 128, True), (1, 1, 200, 64, True), (1, 2, 384, 128, False), (1, 1, 77, 64, False),
                                 (1, 3, 640, 128, True)]:
        q, k, v = ref_inputs(B, H, S, D, torch.bfloat16, seed=20, device="cuda")
        o, _ = tfa.fwd(q, k, v, causal, D ** -0.5, out_fp32=True)
        o16, _ = tfa.fwd(q, k, v, causal, D ** -0.5)
        torch.cuda.synchronize()
        w = want(q, k, v, causal, D ** -0.5)
        e = float(np.abs(o.cpu().numpy() - w).max())
        e16 = float(np.abs(o16.float().cpu().numpy() - w).max())
        print(f"fwd B{B} H{H} S{S} D{D} causal={causal}: max err fp32-out {e:.2e}, 16-bit {e16:.2e}")
        bad += e > 1.5e-3 or e16 > 1e-2
    # generalised problem: GQA + Sq != Sk + split-KV
    g = torch.Generator().manual_seed(3)
    mk = lambda *s: torch.empty(s).normal_(0, 0.5, generator=g).to(torch.bfloat16).cuda()
    q, k, v = mk(1, 4, 96, 128), mk(1, 2, 1000, 128), mk(1, 2, 1000, 128)
    for ns in (1, 3):
        o, _ = tfa.attn_fwd(q, k, v, True, 128 ** -0.5, num_splits=ns, out_fp32=True)
        torch.cuda.synchronize()
        e = float(np.abs(o.cpu().numpy() - want(q, k, v, True, 128 ** -0.5)).max())
        print(f"attn_fwd GQA Sq96 Sk1000 splits={ns}: max err {e:.2e}")
        bad += e > 1.5e-3
    # fused exchange with 3 local "peer" copies
    q, k, v = ref_inputs(1, 2, 384, 128, torch.bfloat16, seed=5, device="cuda")
    bufs = [torch.zeros(4, 2, 384, 128, dtype=torch.bfloat16, device="cuda") for _ in range(4)]
    sl = 2 * 384 * 128 * 2
    tfa.fwd_multi(q, k, v, True, 128 ** -0.5, bufs[2][2:3], [bufs[r].data_ptr() + 2 * sl for r in (0, 1, 3)])
    torch.cuda.synchronize()
    w = want(q, k, v, True, 128 ** -0.5)
    for r in range(4):
        e = float(np.abs(bufs[r][2:3].float().cpu().numpy() - w).max())
        print(f"fused copy {r}: max err {e:.2e}")
        bad += e > 1e-2
    print("SANITIZE_CASES_" + ("OK" if not bad else "BAD"))
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
