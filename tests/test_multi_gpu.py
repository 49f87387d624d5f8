"""GPU, >= 2 devices (skipped on the 1-GPU box): the fused attention + all-gather (peer stores from the kernel epilogue)
must equal kernel + ncclAllGather bit for bit.  Run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`."""
import os
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, os.path.join(os.environ["TFA_ROOT"], "tiny-flash-attention_b200"))
import tfa_ctypes as tfa
from sharded import FusedGather, shard_batch
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(int(os.environ["LOCAL_RANK"]))
dev = torch.device("cuda", int(os.environ["LOCAL_RANK"]))
dist.init_process_group("nccl", device_id=dev)
B, H, S, D = 2 * world, 4, 1024, 128
g = torch.Generator(device=dev).manual_seed(20)
q, k, v = (torch.empty(B, H, S, D, dtype=torch.bfloat16, device=dev).normal_(0, 0.5, generator=g) for _ in range(3))
lo, hi = shard_batch(B, rank, world)
fg = FusedGather(B, H, S, D, torch.bfloat16, dev)
buf, lse = fg.forward(q[lo:hi], k[lo:hi], v[lo:hi], True, D ** -0.5)
want, want_lse = tfa.fwd(q, k, v, True, D ** -0.5)          # every rank computes the whole job for reference
torch.cuda.synchronize()
ok = torch.equal(buf, want) and torch.equal(lse, want_lse[lo:hi])
t = torch.tensor([1 if ok else 0], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MIN)
if rank == 0: print("FUSED_GATHER_OK" if int(t.item()) else "FUSED_GATHER_MISMATCH", flush=True)
dist.destroy_process_group()
'''


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs >= 2 GPUs")
def test_fused_gather_matches_single_gpu_result(built, tmp_path):
    n = min(torch.cuda.device_count(), 8)
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    env = dict(os.environ, TFA_ROOT=ROOT)
    p = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={n}",
                        "--master-addr", "127.0.0.1", "--master-port", "29541", str(script)],
                       capture_output=True, text=True, timeout=600, env=env)
    assert "FUSED_GATHER_OK" in p.stdout, p.stdout[-2000:] + p.stderr[-3000:]
