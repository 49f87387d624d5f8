"""GPU: BASELINE.json's full-size configurations, checked through size-independent properties and a
float32 torch statement of the same math on sampled heads (the CPU oracle would need minutes there)."""
import math

import pytest
import torch

from helpers import ref_inputs

pytestmark = pytest.mark.gpu

CONFIGS = {
    "cfg2": (4, 16, 2048, 64, False),
    "cfg3": (4, 32, 4096, 128, True),
    "cfg4": (1, 32, 16384, 128, True),
}


@pytest.fixture(scope="module")
def mod(built):
    import attention_cutlass
    return attention_cutlass


def torch_fp32_attention(q, k, v, causal, scale):
    """fp32 softmax(scale QK^T)V for a few heads; TF32 off so it is a true fp32 statement."""
    old = torch.backends.cuda.matmul.allow_tf32
    torch.backends.cuda.matmul.allow_tf32 = False
    try:
        qf, kf, vf = q.float(), k.float(), v.float()
        s = torch.matmul(qf, kf.transpose(-1, -2)) * scale
        if causal:
            S = q.shape[-2]
            s.masked_fill_(torch.ones(S, S, device=q.device, dtype=torch.bool).triu_(1), float("-inf"))
        lse = torch.logsumexp(s, dim=-1)
        p = torch.softmax(s, dim=-1)
        return torch.matmul(p, vf), lse
    finally:
        torch.backends.cuda.matmul.allow_tf32 = old


@pytest.mark.parametrize("cfg", list(CONFIGS))
def test_full_size_config_vs_fp32_on_sampled_heads(mod, cfg):
    B, H, S, D, causal = CONFIGS[cfg]
    q, k, v = ref_inputs(B, H, S, D, torch.bfloat16, seed=20, device="cuda")
    scale = 1.0 / math.sqrt(D)
    out, lse = mod.flash_attention_v2_cutlass(q, k, v, causal, scale)
    o32, _ = mod.flash_attention_v2_fp32out(q, k, v, causal, scale)
    torch.cuda.synchronize()
    assert out.shape == q.shape and out.dtype == q.dtype and lse.shape == (B, H, S) and lse.dtype == torch.float32
    for (b, h) in {(0, 0), (B - 1, H - 1), (B // 2, H // 3)}:
        want, want_lse = torch_fp32_attention(q[b, h], k[b, h], v[b, h], causal, scale)
        # kernel arithmetic (fp32 out): P is rounded to bf16 before PV, hence 1e-3 not 1e-6
        torch.testing.assert_close(o32[b, h], want, rtol=1e-3, atol=1e-3)
        torch.testing.assert_close(lse[b, h], want_lse, rtol=0, atol=2e-4)
        # shipped bf16 output: within one bf16 ulp (+1e-3) of the fp32 statement, and the reference's 1e-2 bar
        diff = (out[b, h].float() - want).abs()
        ulp = torch.pow(2.0, torch.floor(torch.log2(want.abs().clamp_min(1e-30))) - 7)
        assert bool((diff <= ulp + 1e-3).all())
        assert float(diff.max()) <= 1e-2
    assert bool(torch.isfinite(out.float()).all()) and bool(torch.isfinite(lse).all())


def test_v_of_ones_gives_ones(mod):
    """softmax rows sum to 1: with V == 1 every output element is exactly representable 1 +- rounding."""
    B, H, S, D = 2, 4, 4096, 128
    q, k, _ = ref_inputs(B, H, S, D, torch.bfloat16, seed=3, device="cuda")
    v = torch.ones_like(q)
    o32, _ = mod.flash_attention_v2_fp32out(q, k, v, True, 0.088)
    torch.cuda.synchronize()
    assert float((o32 - 1).abs().max()) < 2e-3          # sum_j rn_bf16(p_j) / sum_j p_j


def test_linearity_in_v(mod):
    B, H, S, D = 1, 8, 2048, 128
    q, k, v1 = ref_inputs(B, H, S, D, torch.bfloat16, seed=4, device="cuda")
    _, _, v2 = ref_inputs(B, H, S, D, torch.bfloat16, seed=5, device="cuda")
    f = lambda v: mod.flash_attention_v2_fp32out(q, k, v, True, 0.088)[0]
    # 2*v1 and v1+v2 stay exactly representable often enough only for the power-of-two case -> test scaling exactly
    assert torch.equal(f(v1 * 2), f(v1) * 2)
    lhs = f((v1.float() + v2.float()).to(torch.bfloat16))
    torch.testing.assert_close(lhs, f(v1) + f(v2), rtol=0, atol=2e-2)


def test_causal_prefix_property_bitwise(mod):
    """Causal attention of the first n rows does not depend on anything after them: the first n rows of the
    S-long problem equal the n-long problem bit for bit (n a multiple of the 256-row CTA tile)."""
    B, H, S, D = 1, 4, 2048, 128
    q, k, v = ref_inputs(B, H, S, D, torch.bfloat16, seed=6, device="cuda")
    full, lse_full = mod.flash_attention_v2_cutlass(q, k, v, True, 0.088)
    for n in (256, 1024):
        part, lse_part = mod.flash_attention_v2_cutlass(q[:, :, :n].contiguous(), k[:, :, :n].contiguous(),
                                                        v[:, :, :n].contiguous(), True, 0.088)
        assert torch.equal(part, full[:, :, :n])
        assert torch.equal(lse_part, lse_full[:, :, :n])


def test_key_permutation_invariance_noncausal(mod):
    B, H, S, D = 1, 4, 1024, 64
    q, k, v = ref_inputs(B, H, S, D, torch.bfloat16, seed=8, device="cuda")
    perm = torch.randperm(S, device="cuda")
    a, la = mod.flash_attention_v2_fp32out(q, k, v, False, 0.125)
    b, lb = mod.flash_attention_v2_fp32out(q, k[:, :, perm].contiguous(), v[:, :, perm].contiguous(), False, 0.125)
    torch.testing.assert_close(a, b, rtol=1e-3, atol=1e-3)
    torch.testing.assert_close(la, lb, rtol=0, atol=2e-4)


def test_lse_defines_softmax_normaliser(mod):
    """exp(scale*s_ij - lse_i) must sum to 1 over visible keys (checks LSE and masking at a ragged length)."""
    B, H, S, D = 1, 2, 333, 64
    q, k, v = ref_inputs(B, H, S, D, torch.float16, seed=10, device="cuda")
    _, lse = mod.flash_attention_v2_cutlass(q, k, v, True, 0.125)
    s = torch.matmul(q.float(), k.float().transpose(-1, -2)) * 0.125
    s.masked_fill_(torch.ones(S, S, device="cuda", dtype=torch.bool).triu_(1), float("-inf"))
    total = torch.exp(s - lse[..., None]).sum(-1)
    torch.testing.assert_close(total, torch.ones_like(total), rtol=0, atol=1e-3)
