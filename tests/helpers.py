"""Shared helpers for the test-suite (seeded inputs in the reference's distribution, metrics)."""
import math
import os

import numpy as np
import torch

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    return np.load(os.path.join(GOLDEN, name))


def bits_to_f32(bits, kind="bf16"):
    t = torch.from_numpy(np.ascontiguousarray(bits)).view(torch.bfloat16 if kind == "bf16" else torch.float16)
    return t.float().numpy()


def bits_to_tensor(bits, kind="bf16"):
    return torch.from_numpy(np.ascontiguousarray(bits)).view(torch.bfloat16 if kind == "bf16" else torch.float16)


def ref_inputs(B, H, S, D, dtype=torch.bfloat16, seed=20, device="cpu"):
    """q,k,v ~ N(0, 0.5^2), drawn in that order after manual_seed(seed)
    (reference: flash_attention_cutlass/test.py:14-16, tiny_flash_attn_triton.py:221-224)."""
    g = torch.Generator(device="cpu")
    g.manual_seed(seed)
    out = []
    for _ in range(3):
        t = torch.empty((B, H, S, D), dtype=torch.float32).normal_(mean=0.0, std=0.5, generator=g).to(dtype)
        out.append(t.to(device))
    return out


def err_stats(x, ref, rtol=1e-3, atol=1e-3):
    x = np.asarray(x, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    diff = np.abs(x - ref)
    ok = diff <= atol + rtol * np.abs(ref)
    return {"max_abs": float(diff.max()), "pass_frac": float(ok.mean()),
            "max_rel": float((diff / np.maximum(np.abs(ref), 1e-6)).max())}


def bf16_ulp(x):
    """spacing of bf16 numbers at |x| (8 significant bits)."""
    ax = np.maximum(np.abs(np.asarray(x, dtype=np.float64)), 2.0 ** -126)
    return 2.0 ** (np.floor(np.log2(ax)) - 7)


def fp16_ulp(x):
    ax = np.maximum(np.abs(np.asarray(x, dtype=np.float64)), 2.0 ** -14)
    return 2.0 ** (np.floor(np.log2(ax)) - 10)


def default_scale(D):
    return 1.0 / math.sqrt(D)
