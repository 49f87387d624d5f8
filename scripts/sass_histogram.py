#!/usr/bin/env python
"""SASS opcode histogram of the shipped kernels (cuobjdump -sass on the in-tree libtfa_b200.so) -- the Blackwell-native
evidence the profiling guide asks for: UTCHMMA (tcgen05.mma), UTCBAR (tcgen05.commit), LDTM/STTM (tcgen05.ld/st),
UTMALDG/UTMASTG (TMA load/store), SYNCS (mbarrier); and the absence of HMMA / LDSM / LDGSTS (legacy mma.sync path).
   python scripts/sass_histogram.py [kernel-name-regex] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tiny-flash-attention_b200", "libtfa_b200.so")
KEY = ["UTCHMMA", "UTCBAR", "UTCATOMSWS", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAPF", "UTMACCTL", "UBLKCP", "SYNCS",
       "USETMAXREG", "MUFU", "FFMA2", "FADD2", "FMNMX3", "F2FP", "HMMA", "LDSM", "LDGSTS", "STG", "LDG", "STS", "LDS", "STL", "LDL"]


def main():
    pat = re.compile(sys.argv[1] if len(sys.argv) > 1 else r"fa_fwd_sm100_persist_kernel|splitkv_combine|empty_rows_fill")
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    cur, hists, sizes = None, collections.OrderedDict(), {}
    for ln in out.splitlines():
        m = re.match(r"\s+Function : (\S+)", ln)
        if m:
            cur = m.group(1)
            if pat.search(cur):
                hists[cur] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", ln)
        if m and cur in hists:
            hists[cur][m.group(1) + ("." + m.group(2).split(".")[1] if m.group(1) in ("UTMALDG", "UTMASTG", "LDTM", "STTM") and m.group(2) else "")] += 1
    print(f"# SASS opcode histogram, {os.path.relpath(LIB, ROOT)} (cuobjdump -sass), kernels matching /{pat.pattern}/\n")
    tot = collections.Counter()
    for fn, h in hists.items():
        demangled = subprocess.run(["cu++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        n = sum(h.values())
        print(f"## {demangled[:150]}\n   {n} instructions ({n * 16 // 1024} KB)")
        keys = [k for k in h if any(k.startswith(x) for x in KEY)]
        print("   " + "  ".join(f"{k}={h[k]}" for k in sorted(keys)))
        tot.update(h)
    print("\n## all listed kernels together")
    for k in sorted(tot, key=lambda k_: -tot[k_])[:60]:
        print(f"   {k:24s} {tot[k]}")
    legacy = {k: tot[k] for k in tot if k.startswith(("HMMA", "LDSM", "LDGSTS", "HGMMA"))}
    print(f"\nlegacy tensor path (HMMA / LDSM / LDGSTS / HGMMA): {legacy if legacy else 'none'}")


if __name__ == "__main__":
    main()
