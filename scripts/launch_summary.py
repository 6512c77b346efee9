"""Summarise an ncu launch list (`--metrics gpu__time_duration.sum --csv`): launches, total and share per kernel.
usage: launch_summary.py launches.csv "<command line that produced it>" """
import csv, re, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki, vi, mi = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Name")
ui = hdr.index("Metric Unit")
agg = {}
for r in rows[1:]:
    if r[mi] != "gpu__time_duration.sum":
        continue
    us = float(r[vi].replace(",", "")) * {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(r[ui], 1e-3)
    m = re.search(r"cmbl::(\w+<[^>]*>)", r[ki])
    name = m.group(1) if m else "torch/other: " + r[ki][:50]
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
print(sys.argv[2] if len(sys.argv) > 2 else "")
print("(per-launch times are cold-cache and serialised under ncu: compare SHARES, not absolutes)\n")
print(f"{'kernel':70s} {'launches':>8s} {'total us':>12s} {'avg us':>9s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:70s} {v[0]:8d} {v[1]:12.1f} {v[1]/v[0]:9.1f} {100*v[1]/tot:6.1f}%")
