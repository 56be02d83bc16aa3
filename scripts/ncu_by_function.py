#!/usr/bin/env python
"""Attribute executed warp instructions of an `ncu --page source --csv --print-source cuda,sass`
export to source regions (file + line ranges named below): where the instruction budget of the
stage kernels goes.  usage: ncu_by_function.py export.csv zones_per_launch"""
import csv
import re
import sys
from collections import OrderedDict, defaultdict

path, zones = sys.argv[1], float(sys.argv[2])
rows = list(csv.reader(open(path)))

# (file suffix, first line, last line, label) -- resolved from the sources at run time
import os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
dev = open(os.path.join(ROOT, "artemis_b200/csrc/ab200_dev.cuh")).read().splitlines()


def span(pattern, end_pattern=r"^}"):
    """line range [start, end] of the first top-level construct matching `pattern`"""
    for i, l in enumerate(dev):
        if re.search(pattern, l):
            for j in range(i + 1, len(dev)):
                if re.match(end_pattern, dev[j]):
                    return i + 1, j + 1
    return None


regions = OrderedDict()
for label, pat in (("division / sqrt helpers", r"AB_D double drcp"), ("ppm_iface", r"AB_D double ppm_iface"),
                   ("ppm_mono", r"AB_D void ppm_mono"), ("plm", r"AB_D void plm\("),
                   ("ppm4", r"AB_D void ppm4")):
    sp = span(pat)
    if sp:
        regions[label] = sp
# the Riemann solvers: from "struct Riemann" to the end of the file section
for i, l in enumerate(dev):
    if re.search(r"struct Riemann", l):
        regions.setdefault("Riemann solvers", (i + 1, len(dev)))
        break
# ddiv/dsqrt are three small functions in a row
if "division / sqrt helpers" in regions:
    s0 = regions["division / sqrt helpers"][0]
    regions["division / sqrt helpers"] = (s0, s0 + 40)

kernels, cur_file, cur_fn = OrderedDict(), None, None
hdr = None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1]
        continue
    if r[0] == "Function Name":
        cur_fn = r[1]
        continue
    if r[0] == "Line No":
        hdr = {h: i for i, h in enumerate(r)}
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    try:
        n = int(r[hdr["Instructions Executed"]])
    except ValueError:
        continue
    line = int(r[0])
    label = os.path.basename(cur_file) + " (kernel body)"
    if cur_file.endswith("ab200_dev.cuh"):
        label = "ab200_dev.cuh other (Coords, set_aux, ...)"
        for k, (a, b) in regions.items():
            if a <= line <= b:
                label = k
                break
    kernels.setdefault(cur_fn, defaultdict(int))[label] += n

for fn, d in kernels.items():
    tot = sum(d.values())
    if tot == 0:
        continue
    print("=" * 100)
    print(fn[:110])
    print("warp instructions %.1f M = %.0f thread instructions per zone" % (tot / 1e6, 32 * tot / zones))
    for k, v in sorted(d.items(), key=lambda x: -x[1]):
        print("  %-45s %8.1f M  %5.1f%%  %6.0f /zone" % (k, v / 1e6, 100 * v / tot, 32 * v / zones))
