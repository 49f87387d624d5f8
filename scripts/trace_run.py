"""Timeline of one CTA from a -DTFA_TRACE build (TFA_LIB=.../libtfa_b200_trace.so).
   python scripts/trace_run.py '{"B":4,"H":32,"S":4096,"D":128,"causal":true,"block":100}'"""
import ctypes
import json
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
import tfa_ctypes  # noqa: E402

ROLE = {0: "softmax0", 1: "softmax1", 2: "mma", 3: "loader"}
EV_SM = {1: "start", 2: "S_ready", 3: "ld_done", 4: "max/rescale_done", 5: "exp/pack/st_issued", 6: "P_arrived", 7: "O_ready", 8: "epi_done"}
EV_MMA = {1: "start", 2: "K0_ready", 3: "Q_ready", 4: "S(0)_issued", 5: "V_ready", 6: "P0_ready", 7: "P1_ready",
          8: "PV0_issued", 9: "PV1_issued", 10: "K_ready", 12: "S0_next_issued", 13: "S1_next_issued",
          14: "S0_HOISTED(next item)", 15: "S1_HOISTED(next item)"}
EV_LD = {1: "start", 2: "slot_free"}


def main():
    kw = json.loads(sys.argv[1]) if len(sys.argv) > 1 else {}
    B, H, S, D = kw.get("B", 4), kw.get("H", 32), kw.get("S", 4096), kw.get("D", 128)
    causal, block = kw.get("causal", True), kw.get("block", 0)
    L = tfa_ctypes.lib()
    L.tfa_internal_set_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
    L.tfa_internal_set_trace.restype = None
    buf = torch.zeros(4 * 512, dtype=torch.int64, device="cuda")
    q, k, v = (torch.empty(B, H, S, D, dtype=torch.bfloat16, device="cuda").normal_(0, 0.5) for _ in range(3))
    for _ in range(2):
        tfa_ctypes.fwd(q, k, v, causal, 1 / math.sqrt(D))
    torch.cuda.synchronize()
    L.tfa_internal_set_trace(buf.data_ptr(), block)
    tfa_ctypes.fwd(q, k, v, causal, 1 / math.sqrt(D))
    torch.cuda.synchronize()
    L.tfa_internal_set_trace(None, 0)
    raw = buf.cpu().view(4, 512)
    evs = []
    for role in range(4):
        n = int(raw[role, 0])
        for i in range(n):
            x = int(raw[role, 1 + i]) & ((1 << 64) - 1)
            evs.append((x >> 8, role, x & 0xff))
    if not evs:
        print("no events (is TFA_LIB the trace build?)")
        return
    t0 = min(e[0] for e in evs)
    evs.sort()
    names = {0: EV_SM, 1: EV_SM, 2: EV_MMA, 3: EV_LD}
    limit = kw.get("limit", 400)
    last = {}
    for t, role, ev in evs[:limit]:
        dt = t - last.get(role, t)
        last[role] = t
        print(f"{t - t0:9d}  (+{dt:6d})  {ROLE[role]:9s} {names[role].get(ev, ev)}")
    print(f"total span {evs[-1][0] - t0} cycles, {len(evs)} events")
    # steady-state summary per softmax role: mean interval between consecutive events of one KV-tile iteration
    for role in (0, 1):
        seq = [(t, ev) for t, r, ev in evs if r == role]
        iters, cur = [], {}
        for t, ev in seq:
            if ev == 2:
                if cur:
                    iters.append(cur)
                cur = {2: t}
            elif cur and ev in (3, 4, 5, 6):
                cur[ev] = t
        if cur:
            iters.append(cur)
        mid = [it for it in iters[2:-2] if all(e in it for e in (2, 3, 4, 5, 6))]
        if len(mid) < 2:
            continue
        def mean(f):
            xs = [f(a) for a in mid]
            return sum(xs) / len(xs)
        period = (mid[-1][2] - mid[0][2]) / (len(mid) - 1)
        print(f"{ROLE[role]}: iters {len(mid)}  period {period:.0f}  S_ready->ld+mask+max(3) {mean(lambda a: a[3]-a[2]):.0f}  "
              f"->rescale chk(4) {mean(lambda a: a[4]-a[3]):.0f}  ->last early hand-off(5) {mean(lambda a: a[5]-a[4]):.0f}  "
              f"->p_full(6) {mean(lambda a: a[6]-a[5]):.0f}  softmax total {mean(lambda a: a[6]-a[2]):.0f}  "
              f"p_full->next S_ready {period - mean(lambda a: a[6]-a[2]):.0f}")


if __name__ == "__main__":
    main()
