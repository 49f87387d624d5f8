"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count, total time, share."""
import csv
import sys
from collections import defaultdict

rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 14 and r[0].isdigit()]
tot = defaultdict(lambda: [0, 0.0])
for r in rows:
    name = r[4].split("(")[0][-90:]
    val = float(r[14].replace(",", ""))
    unit = r[13]
    us = val / 1000.0 if unit in ("ns", "nsecond") else (val * 1000.0 if unit in ("ms", "msecond") else val)
    tot[name][0] += 1
    tot[name][1] += us
total = sum(v[1] for v in tot.values())
print(f"# {len(rows)} launches, total {total:.1f} us (serialised, cold-cache: compare SHARES)")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{us:12.1f} us  {100 * us / total:6.2f}%  n={n:4d}  avg={us / n:10.1f} us  {name}")
