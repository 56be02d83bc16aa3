#!/usr/bin/env python
"""Per-kernel shares of an `ncu --metrics gpu__time_duration.sum,dram__bytes_*` launch list
(cold-cache, serialised: compare SHARES, not absolutes).  usage: launch_shares.py csv [skip]
`skip` = launches to drop from the front (initialisation / warm-up)."""
import collections
import csv
import sys

lines = [l for l in open(sys.argv[1]) if not l.startswith("==")]
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
rows = collections.OrderedDict()
for row in csv.DictReader(lines):
    rows.setdefault((int(row["ID"]), row["Kernel Name"]), {})[row["Metric Name"]] = (
        float(row["Metric Value"].replace(",", "")), row["Metric Unit"])
agg = collections.OrderedDict()
tot = 0.0
for (idx, k), m in rows.items():
    if idx < skip:
        continue
    t, u = m["gpu__time_duration.sum"]
    t = t / 1e3 if u == "ns" else (t * 1e3 if u == "ms" else t)

    def gb(x):
        v, un = m[x]
        return v * {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1}[un]
    name = k.split("(")[0].replace("void ", "")[:60]
    a = agg.setdefault(name, [0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += t
    a[2] += gb("dram__bytes_read.sum")
    a[3] += gb("dram__bytes_write.sum")
    tot += t
print(f"launches {sum(a[0] for a in agg.values())} (first {skip} skipped), total {tot:.1f} us")
print(f"{'kernel':62s} {'n':>4s} {'us/launch':>10s} {'share':>7s} {'GB rd/l':>8s} {'GB wr/l':>8s}")
for name, (n, t, rd, wr) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{name:62s} {n:4d} {t / n:10.1f} {100 * t / tot:6.1f}% {rd / n:8.3f} {wr / n:8.3f}")
