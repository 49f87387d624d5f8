"""CPU, world_size 2, gloo: the multi-GPU host logic (shard over batch*head, all-gather O).
The compute is injected: here the CPU oracle stands in for the CUDA operator (tests only)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))

from sharded import shard_batch, shard_bounds, sharded_forward  # noqa: E402


def test_shard_bounds_cover_and_balance():
    for n in (1, 7, 8, 2048, 2049):
        for world in (1, 2, 3, 8):
            spans = [shard_bounds(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            for (a, b), (c, d) in zip(spans, spans[1:]):
                assert b == c
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(4, 2, 2)
    assert shard_batch(64, 3, 8) == (24, 32)      # BASELINE config 5: 8 batches (x32 heads) per GPU


def _oracle_attn(q, k, v, causal, scale):
    from oracle import oracle as orc
    o, lse = orc.attn_exact(q.float().numpy(), k.float().numpy(), v.float().numpy(), causal, scale, orc.ROUND_NONE)
    return torch.from_numpy(o), torch.from_numpy(lse)


def _worker(rank, world, port, n_chunks, ret):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        torch.set_num_threads(2)
        B, H, S, D = 4, 2, 64, 64
        g = torch.Generator().manual_seed(20)
        q, k, v = (torch.empty(B, H, S, D).normal_(0, 0.5, generator=g) for _ in range(3))
        lo, hi = shard_batch(B, rank, world)
        out_full, lse_local = sharded_forward(q[lo:hi], k[lo:hi], v[lo:hi], True, 0.125, _oracle_attn,
                                              n_chunks=n_chunks)
        want, want_lse = _oracle_attn(q, k, v, True, 0.125)
        ok = bool(torch.allclose(out_full, want, atol=1e-6)) and bool(
            torch.allclose(lse_local, want_lse[lo:hi], atol=1e-6)) and out_full.shape == want.shape
        ret[rank] = ok
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("n_chunks", [1, 2])
def test_sharded_forward_gloo_world2(n_chunks):
    world = 2
    port = 29500 + (os.getpid() % 2000) + n_chunks
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, n_chunks, ret), nprocs=world, join=True)
    assert all(ret.get(r) for r in range(world)), dict(ret)


def test_single_process_no_gather():
    q = torch.randn(2, 1, 32, 64)
    o, lse = sharded_forward(q, q, q, False, 0.125, _oracle_attn, gather=False)
    assert o.shape == q.shape and lse.shape == (2, 1, 32)
    assert np.isfinite(o.numpy()).all()
