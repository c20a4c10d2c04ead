"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel shares (profiles/*_summary.txt)."""
import csv, re, sys
from collections import defaultdict
path = sys.argv[1]
rows = list(csv.reader(l for l in open(path, errors="replace") if l.startswith('"')))
hdr = rows[0]
ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
acc = defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index("Metric Name")] != "gpu__time_duration.sum":
        continue
    name = re.sub(r"<.*", "", r[ki]).replace("void ", "").replace("fz::", "")
    t = float(r[vi].replace(",", ""))
    t_us = t / 1e3 if r[ui] in ("nsecond", "ns") else (t if r[ui] in ("usecond", "us") else t * 1e3)
    acc[name][0] += 1
    acc[name][1] += t_us
tot = sum(v[1] for v in acc.values())
print(f"{'kernel':42s} {'launches':>8s} {'total ms':>10s} {'share':>7s} {'avg us':>9s}")
for k, (n, t) in sorted(acc.items(), key=lambda kv: -kv[1][1]):
    print(f"{k[:42]:42s} {n:8d} {t / 1e3:10.3f} {100 * t / tot:6.1f}% {t / n:9.1f}")
print(f"{'TOTAL':42s} {sum(v[0] for v in acc.values()):8d} {tot / 1e3:10.3f}")
