"""Multi-GPU driver of the forward: shard the independent (batch*head) problems across ranks.

The reference has no distributed code at all; BASELINE.json's north_star defines the
one exchange: every rank computes its contiguous slice of the flattened (B*H) heads -- the same
independence the reference's grid uses (blockIdx.y = b*H+h, flash_attention.cu:382,409) -- and the
outputs O are all-gathered (NCCL over NVLink/NVSwitch; gloo in the CPU tests).  LSE is not gathered.

`attn_fn(q, k, v, is_causal, scale) -> (out, lse)` is injected so the host logic can be exercised
on CPU with gloo (tests/test_sharding.py); the product wiring passes the CUDA operator.
"""
from __future__ import annotations

from typing import Callable, Tuple


def shard_bounds(n_units: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced [lo, hi) slice of `n_units` for `rank` (first n%world ranks get one extra)."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_units, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def shard_batch(B: int, rank: int, world: int) -> Tuple[int, int]:
    """Shard on the batch axis so each rank owns whole (H,S,D) slabs and O stays contiguous."""
    return shard_bounds(B, rank, world)


def sharded_forward(q_local, k_local, v_local, is_causal: bool, scale: float, attn_fn: Callable,
                    group=None, gather: bool = True, n_chunks: int = 1, out_full=None):
    """Run attn_fn on this rank's (B_local,H,S,D) shard and all-gather O along the batch axis.

    Requires equal B_local on every rank (all_gather_into_tensor).  With n_chunks > 1 the local
    batch is processed in chunks and each chunk's all-gather is issued asynchronously so that the
    collective of chunk i overlaps the kernel of chunk i+1.
    Returns (o_full or o_local, lse_local).
    """
    import torch
    import torch.distributed as dist

    world = dist.get_world_size(group) if (gather and dist.is_initialized()) else 1
    Bl = q_local.shape[0]
    if not gather or world == 1:
        return attn_fn(q_local, k_local, v_local, is_causal, scale)
    rank = dist.get_rank(group)
    n_chunks = max(1, min(n_chunks, Bl))
    if out_full is None:
        out_full = torch.empty((world * Bl,) + tuple(q_local.shape[1:]), dtype=q_local.dtype,
                               device=q_local.device)
    lses, works, stage = [], [], []
    for c in range(n_chunks):
        lo, hi = shard_bounds(Bl, c, n_chunks)
        o_c, lse_c = attn_fn(q_local[lo:hi], k_local[lo:hi], v_local[lo:hi], is_causal, scale)
        lses.append(lse_c)
        if n_chunks == 1:
            works.append(dist.all_gather_into_tensor(out_full, o_c.contiguous(), group=group, async_op=True))
        else:
            # gather chunk c of every rank into a staging buffer, then scatter to the right batch rows
            buf = torch.empty((world * (hi - lo),) + tuple(o_c.shape[1:]), dtype=o_c.dtype, device=o_c.device)
            works.append(dist.all_gather_into_tensor(buf, o_c.contiguous(), group=group, async_op=True))
            stage.append((buf, lo, hi))
    for w in works:
        w.wait()
    for buf, lo, hi in stage:
        out_full.view(world, Bl, *q_local.shape[1:])[:, lo:hi].copy_(buf.view(world, hi - lo, *q_local.shape[1:]))
    return out_full, (lses[0] if len(lses) == 1 else torch.cat(lses, dim=0))


class FusedGather:
    """Fused attention + all-gather of O over NVLink peer memory.

    Every rank owns a symmetric (B_total,H,S,D) buffer (torch symmetric memory: CUDA VMM allocations mapped into every
    rank of the node).  forward() launches ONE kernel per rank that writes its O tiles into its own slice of ALL ranks'
    buffers -- the epilogue's coalesced 128-bit stores are simply repeated for the peer mappings, so the NVLink traffic
    overlaps the attention math tile by tile -- then a stream-ordered symmetric-memory barrier makes every rank's
    buffer complete.  No NCCL collective on the data path.
    """

    def __init__(self, B_total, H, S, D, dtype, device, group=None):
        import torch
        import torch.distributed as dist
        import torch.distributed._symmetric_memory as symm

        self.group = group if group is not None else dist.group.WORLD
        self.rank, self.world = dist.get_rank(self.group), dist.get_world_size(self.group)
        if self.world > 8:
            raise ValueError("FusedGather supports up to 8 ranks (one NVSwitch domain)")
        if B_total % self.world:
            raise ValueError("batch must divide evenly across ranks")
        self.shape = (B_total, H, S, D)
        try:
            symm.enable_symm_mem_for_group(self.group.group_name)
        except Exception:  # noqa: BLE001  (newer torch enables it implicitly)
            pass
        self.buf = symm.empty(self.shape, dtype=dtype, device=device)
        self.hdl = symm.rendezvous(self.buf, self.group.group_name)
        self.peer_ptrs = [int(x) for x in self.hdl.buffer_ptrs]
        self.Bl = B_total // self.world
        self.slice_bytes = self.Bl * H * S * D * self.buf.element_size()

    def launch(self, q_local, k_local, v_local, is_causal, scale, lse=None):
        """The kernel only (asynchronous on the current stream): attention + peer stores."""
        import tfa_ctypes as tfa

        off = self.rank * self.slice_bytes
        out_local = self.buf[self.rank * self.Bl:(self.rank + 1) * self.Bl]
        extra = [self.peer_ptrs[r] + off for r in range(self.world) if r != self.rank]
        _, lse = tfa.fwd_multi(q_local, k_local, v_local, is_causal, scale, out_local, extra, lse=lse)
        return lse

    def barrier(self):
        """Stream-ordered cross-rank barrier: after it, every peer's stores into my buffer have landed."""
        self.hdl.barrier()

    def forward(self, q_local, k_local, v_local, is_causal, scale, lse=None):
        """Returns (gathered O, local LSE).  The gathered tensor IS the symmetric buffer: consume or copy it before the
        next forward() -- the next call's peer stores overwrite it.  The barrier in FRONT of the launch orders those
        stores after every rank's work queued so far (a fast rank must not overwrite a slice a slower peer is still
        reading from the previous step); the barrier BEHIND it makes every rank's buffer complete."""
        self.barrier()
        lse = self.launch(q_local, k_local, v_local, is_causal, scale, lse=lse)
        self.barrier()
        return self.buf, lse
