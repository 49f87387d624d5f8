"""Run the split-KV configuration bench.py reports (B1 H8 Sq128 Sk65536 D128, non-causal) a few times -- the
target of the ncu captures of the combine kernel.  python scripts/splitkv_case.py [num_splits]"""
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
import tfa_ctypes  # noqa: E402

ns = int(sys.argv[1]) if len(sys.argv) > 1 else 0
q = torch.empty(1, 8, 128, 128, dtype=torch.bfloat16, device="cuda").normal_(0, 0.5)
k = torch.empty(1, 8, 65536, 128, dtype=torch.bfloat16, device="cuda").normal_(0, 0.5)
v = torch.empty(1, 8, 65536, 128, dtype=torch.bfloat16, device="cuda").normal_(0, 0.5)
for _ in range(5):
    o, lse, used = tfa_ctypes.attn_fwd(q, k, v, False, 1 / math.sqrt(128), num_splits=ns, return_splits=True)
torch.cuda.synchronize()
print("splits used:", used)
