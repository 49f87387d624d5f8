"""Run ONE forward case in this process and print a JSON verdict incl. the device watchdog record.
   python scripts/one_case.py '{"B":1,"H":2,"S":128,"D":64,"causal":false,"kind":"bf16","out_fp32":true}'"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
sys.path.insert(0, os.path.join(ROOT, "tests"))
import tfa_ctypes  # noqa: E402
from helpers import ref_inputs  # noqa: E402


def main():
    kw = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {}
    B, H, S, D = kw.get("B", 1), kw.get("H", 2), kw.get("S", 128), kw.get("D", 64)
    causal, kind, f32 = kw.get("causal", False), kw.get("kind", "bf16"), kw.get("out_fp32", True)
    dt = torch.bfloat16 if kind == "bf16" else torch.float16
    q, k, v = ref_inputs(B, H, S, D, dt, seed=20, device="cuda")
    scale = D ** -0.5
    res = {"case": kw}
    try:
        o, lse = tfa_ctypes.fwd(q, k, v, causal, scale, out_fp32=f32)
        torch.cuda.synchronize()
    except Exception as e:  # noqa: BLE001
        res.update(ok=False, error=repr(e)[:200], debug=tfa_ctypes.debug_record())
        print("ONE_CASE " + json.dumps(res))
        return
    from oracle import oracle as orc
    if S * S * B * H <= 64 * 1024 * 1024:
        want, want_lse = orc.attn_exact(q.float().cpu().numpy(), k.float().cpu().numpy(), v.float().cpu().numpy(),
                                        causal, scale, orc.ROUND_BF16 if kind == "bf16" else orc.ROUND_FP16, False)
        d = np.abs(o.float().cpu().numpy() - want)
        res.update(ok=bool(d.max() < (1e-3 if f32 else 1e-2)), max_abs=float(d.max()),
                   lse_max_abs=float(np.abs(lse.cpu().numpy() - want_lse).max()),
                   nan=int(np.isnan(o.float().cpu().numpy()).sum()))
        if not res["ok"]:
            bad = np.argwhere(d > 1e-2)
            res["n_bad"] = int(len(bad))
            res["first_bad"] = [int(x) for x in bad[0]] if len(bad) else None
            res["bad_rows"] = sorted(set(int(b[2]) for b in bad))[:20]
            res["bad_cols"] = sorted(set(int(b[3]) for b in bad))[:20]
    else:
        res.update(ok=bool(torch.isfinite(o.float()).all()))
    res["debug"] = tfa_ctypes.debug_record()
    print("ONE_CASE " + json.dumps(res))


if __name__ == "__main__":
    main()
