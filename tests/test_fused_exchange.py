"""GPU (one device is enough): the fused compute + exchange entry point tfa_fwd_multi against the
CPU oracle.  On the 8-GPU job the `extra_out` pointers are the peers' mappings of the gathered output; the kernel does
not care where they live, so here they are further buffers on the SAME device -- the epilogue's peer-store path runs
exactly as it does over NVLink and every copy must be the oracle's attention, bit-identical to the local copy.

Reference arithmetic: flash_attention_c/csrc/attn.cpp:101-167 / flash_attention_cutlass/csrc/flash_attention.cu:373-685
through oracle.attn_exact (pinned by tests/test_oracle.py)."""
import numpy as np
import pytest
import torch

from helpers import bf16_ulp, fp16_ulp, ref_inputs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def tfa(built):
    import tfa_ctypes
    tfa_ctypes.lib()
    return tfa_ctypes


def _oracle(q, k, v, causal, scale, kind):
    from oracle import oracle as orc
    mode = orc.ROUND_BF16 if kind == "bf16" else orc.ROUND_FP16
    return orc.attn_exact(q.float().cpu().numpy(), k.float().cpu().numpy(), v.float().cpu().numpy(), causal, scale,
                          mode, False)


@pytest.mark.parametrize("B,H,S,D,causal,kind,n_extra", [
    (2, 3, 512, 128, True, "bf16", 7),      # the 8-GPU shape of the job: 7 peer copies
    (1, 2, 384, 64, False, "fp16", 3),
    (1, 2, 200, 128, True, "bf16", 1),      # ragged S: rows >= S must not be written to any copy
    (2, 2, 1024, 64, True, "bf16", 2),
])
def test_every_copy_is_the_oracles_attention(tfa, B, H, S, D, causal, kind, n_extra):
    dt = torch.bfloat16 if kind == "bf16" else torch.float16
    q, k, v = ref_inputs(B, H, S, D, dt, seed=20, device="cuda")
    scale = D ** -0.5
    want32, want_lse = _oracle(q, k, v, causal, scale, kind)
    # the gathered buffer of a (n_extra + 1)-rank job: this "rank" owns slice `me`; a guard row pattern detects stray writes
    world, me = n_extra + 1, min(1, n_extra)
    bufs = [torch.full((world * B, H, S, D), 7.0, dtype=dt, device="cuda") for _ in range(world)]
    slice_elems = B * H * S * D
    out_local = bufs[me][me * B:(me + 1) * B]
    extra = [bufs[r].data_ptr() + me * slice_elems * 2 for r in range(world) if r != me]
    _, lse = tfa.fwd_multi(q, k, v, causal, scale, out_local, extra)
    torch.cuda.synchronize()
    ulp = bf16_ulp(want32) if kind == "bf16" else fp16_ulp(want32)
    for r in range(world):
        got = bufs[r][me * B:(me + 1) * B]
        d = np.abs(got.float().cpu().numpy() - want32)
        # same criterion as tests/test_fwd_parity.py / test_general_attn.py: 1e-3 + half a 16-bit ulp; on causal shapes
        # the early rows (1-3 visible keys, P ~ 0.5 each) can flip ONE 16-bit rounding of P and move O by ~1e-3 more
        # (SURVEY.md A.3), so a handful of elements may exceed the line by that much
        excess = d - (1e-3 + 1e-3 * np.abs(want32) + 0.505 * ulp)
        if causal:
            assert (excess <= 0).mean() >= 0.9995 and excess.max() < 2e-3, f"copy {r}: max err {d.max():.3e}, max excess {excess.max():.3e}"
        else:
            assert np.all(excess <= 0), f"copy {r}: max err {d.max():.3e}, max excess {excess.max():.3e}"
        assert torch.equal(got, out_local), f"copy {r} differs from the local copy"
        other = torch.cat([bufs[r][:me * B], bufs[r][(me + 1) * B:]])
        assert bool((other == 7.0).all()), f"copy {r}: wrote outside its slice"
    assert np.abs(lse.cpu().numpy() - want_lse).max() <= 2e-4
    # and the plain entry point gives the same bits
    o_plain, lse_plain = tfa.fwd(q, k, v, causal, scale)
    torch.cuda.synchronize()
    assert torch.equal(o_plain, out_local) and torch.equal(lse_plain, lse)


def test_fused_rejects_what_it_cannot_do(tfa):
    q, k, v = ref_inputs(1, 1, 128, 64, torch.bfloat16, seed=1, device="cuda")
    out = torch.empty_like(q)
    with pytest.raises(tfa.TfaError):
        tfa.fwd_multi(q, k, v, False, 0.125, out, [out.data_ptr() + 8])          # misaligned peer pointer
    with pytest.raises(tfa.TfaError):
        tfa.fwd_multi(q, k, v, False, 0.125, out, [out.data_ptr()] * 8)          # more than 7 peers
