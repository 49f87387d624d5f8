#!/usr/bin/env python
"""profiles/traffic.json from an `ncu --set full` capture of the dominant kernel: DRAM bytes per (batch*head) problem,
stamped with the commit and the digest of the kernel sources it was taken from (bench.py drops the traffic claim when the
sources change).   python scripts/make_traffic_json.py gpurun_out/prof.ncu-rep B H S D"""
import csv
import hashlib
import io
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SOURCES = ["tiny-flash-attention_b200/csrc/fa_fwd_sm100_persist.cuh", "tiny-flash-attention_b200/csrc/fa_fwd_sm100.cuh",
           "tiny-flash-attention_b200/csrc/ptx_sm100.cuh"]


def main():
    rep, B, H, S, D = sys.argv[1], *map(int, sys.argv[2:6])
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(out)))
    hdr, units = rows[0], rows[1]
    d = dict(zip(hdr, rows[2]))
    u = dict(zip(hdr, units))

    def to_bytes(key):
        v = float(d[key].replace(",", ""))
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u[key]]
    rd, wr = to_bytes("dram__bytes_read.sum"), to_bytes("dram__bytes_write.sum")
    h = hashlib.sha256()
    for f in SOURCES:
        h.update(open(os.path.join(ROOT, f), "rb").read())
    commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    rec = {"capture": os.path.basename(rep), "kernel": d.get("Kernel Name"), "shape": [B, H, S, D], "commit": commit,
           "dram_read_bytes": rd, "dram_write_bytes": wr, "dram_bytes_per_head": (rd + wr) / (B * H),
           "algorithmic_bytes_per_head": 4 * S * D * 2 + 4 * S, "sources": SOURCES, "sources_sha256": h.hexdigest()}
    json.dump(rec, open(os.path.join(ROOT, "profiles", "traffic.json"), "w"), indent=1)
    print(json.dumps(rec, indent=1))


if __name__ == "__main__":
    main()
