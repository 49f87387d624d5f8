import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
import torch
import attention_cutlass as m
q = torch.zeros(1, 2, 256, 128, dtype=torch.bfloat16, device="cuda")
m.flash_attention_v2_cutlass(q, q, q, False, 0.1)
torch.cuda.synchronize()
print("launch ok", flush=True)
try:
    m.flash_attention_v2_cutlass(q.transpose(1, 2), q, q, False, 0.1)
except RuntimeError as e:
    print("raise1 ok:", str(e).splitlines()[0], flush=True)
x = torch.zeros(1, 1, 128, 96, dtype=torch.float16, device="cuda")
try:
    m.flash_attention_v2_cutlass(x, x, x, False, 0.1)
except RuntimeError as e:
    print("raise2 ok:", str(e).splitlines()[0], flush=True)
print("done", flush=True)
