"""Pipe / UMMA micro-benchmarks on the real B200 (dev tool; see csrc/tfa_microbench.cu)."""
import ctypes
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tiny-flash-attention_b200"))
import tfa_ctypes  # noqa: E402

L = tfa_ctypes.lib()
vp, ci = ctypes.c_void_p, ctypes.c_int
L.tfa_microbench_pipe.argtypes = [ci, ci, ci, ci, vp, vp, vp]
L.tfa_microbench_pipe.restype = ci
L.tfa_microbench_umma.argtypes = [ci, ci, ci, ci, ci, vp, vp]
L.tfa_microbench_umma.restype = ci
L.tfa_microbench_softmax.argtypes = [ci, ci, ci, ci, ci, vp, vp, vp, vp]
L.tfa_microbench_softmax.restype = ci
NAMES = ["MUFU.EX2", "FFMA", "FFMA2", "FADD2", "FMNMX3", "F2FP.bf16x2", "ex2_poly2(pair)", "FFMA2+2xMUFU(pair)"]
sink = torch.zeros(4, device="cuda")
cyc = torch.zeros(1024, dtype=torch.int64, device="cuda")
out = {}
iters = 2000
if "umma2" in sys.argv[1:]:
    # 2-CTA (cta_group::2) MMA: cycles per M=256 instruction of a CTA pair, all 74 pairs issuing
    L.tfa_microbench_umma2.argtypes = [ci, ci, ci, ci, vp, vp]
    L.tfa_microbench_umma2.restype = ci
    for (N, form, fname) in ((128, 0, "SS M=256 N=128 (QK^T, two Q tiles' rows)"), (128, 1, "TS M=256 N=128 (PV D=128)"),
                             (64, 1, "TS M=256 N=64 (PV D=64)"), (256, 0, "SS M=256 N=256")):
        n_mma = 4096
        for _ in range(2):
            tfa_ctypes.check(L.tfa_microbench_umma2(148, n_mma, N, form, cyc.data_ptr(), None))
        torch.cuda.synchronize()
        c = cyc[:74].float().median().item() / n_mma
        flops = 2 * 256 * N * 16
        print(f"UMMA2 {fname:40s}: {c:7.2f} SM-cycles per MMA ({flops / c / 2:7.0f} flop/clk/SM)")
    sys.exit(0)
if "softmax" in sys.argv[1:]:
    # exponential phase of the softmax in isolation: cycles per 128-element row per warp
    inp = torch.empty(4100, device="cuda").normal_(0, 2.0)
    inp[4096], inp[4097] = 0.1275, -0.3
    for emu in (0, 1, 2, 3):
        for cw, sw in ((4, 0), (4, 4), (4, 8), (8, 0), (8, 4)):
            for _ in range(2):
                tfa_ctypes.check(L.tfa_microbench_softmax(emu, 148, cw, sw, 500, inp.data_ptr(), sink.data_ptr(),
                                                          cyc.data_ptr(), None))
            torch.cuda.synchronize()
            c = cyc[:148].float().median().item() / 500
            print(f"SOFTMAX emu={emu}/8 compute warps/SMSP={cw // 4} spinning warps/SMSP={sw // 4}: "
                  f"{c:7.1f} cycles per 128-key row per warp")
    sys.exit(0)
for which, name in enumerate(NAMES):
    for warps_per_smsp in (1, 2, 4):
        nthreads = 128 * warps_per_smsp
        for _ in range(2):
            tfa_ctypes.check(L.tfa_microbench_pipe(which, 148, nthreads, iters, sink.data_ptr(), cyc.data_ptr(), None))
        torch.cuda.synchronize()
        c = cyc[:148].float().median().item()
        n_inst = iters * 16 if which not in (2, 3, 6, 7) else iters * 8   # warp-instructions (pairs for the x2 forms)
        per = c / n_inst
        # cycles per warp-instruction per SMSP (divide by warps sharing the SMSP)
        out[f"{name} w/smsp={warps_per_smsp}"] = round(per / warps_per_smsp, 3)
        print(f"PIPE {name:22s} warps/SMSP={warps_per_smsp}: {per:7.2f} cyc/instr/warp -> {per / warps_per_smsp:6.2f} cyc per instr per SMSP")
for (N, form, fname) in ((128, 0, "SS N=128 (QK^T)"), (128, 1, "TS N=128 (PV D=128)"), (64, 1, "TS N=64 (PV D=64)"), (64, 0, "SS N=64")):
    for uniform in (0, 1):
        n_mma = 4096
        for _ in range(2):
            tfa_ctypes.check(L.tfa_microbench_umma(148, n_mma, N, form, uniform, cyc.data_ptr(), None))
        torch.cuda.synchronize()
        c = cyc[:148].float().median().item()
        out[f"UMMA {fname} uniform={uniform}"] = round(c / n_mma, 2)
        flops = 2 * 128 * N * 16
        print(f"UMMA {fname:22s} issue={'elect' if uniform else 'lane0-branch'}: {c / n_mma:7.2f} SM-cycles per MMA "
              f"({flops / (c / n_mma):7.0f} flop/clk/SM)")
print("MICRO " + json.dumps(out))
