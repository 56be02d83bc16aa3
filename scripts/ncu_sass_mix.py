#!/usr/bin/env python
"""Opcode mix + stall samples per opcode from `ncu --page source --csv` (one kernel per block)."""
import collections
import csv
import sys


def main(path, which=0):
    kernels = []
    cur = None
    for row in csv.reader(open(path)):
        if row and row[0] == "Kernel Name":
            cur = {"name": row[1], "rows": [], "hdr": None}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = row
        elif cur is not None and row:
            cur["rows"].append(row)
    k = kernels[which]
    h = k["hdr"]
    isrc, iex, ismp = h.index("Source"), h.index("Instructions Executed"), h.index("# Samples")
    ilsb, iwait, ibar, issb = (h.index("stall_long_sb"), h.index("stall_wait"),
                               h.index("stall_barrier"), h.index("stall_short_sb"))
    mix = collections.Counter()
    smp = collections.Counter()
    lsb = collections.Counter()
    tot = 0
    for r in k["rows"]:
        op = r[isrc].split()
        if not op:
            continue
        o = op[0] if not op[0].startswith("@") else op[1]
        o = o.split(".")[0]
        n = int(r[iex] or 0)
        mix[o] += n
        smp[o] += int(r[ismp] or 0)
        lsb[o] += int(r[ilsb] or 0)
        tot += n
    print(k["name"][:90], "SASS lines:", len(k["rows"]), "warp-inst:", tot)
    st = sum(smp.values())
    for o, n in mix.most_common(28):
        print(f"  {o:10s} {n:12d} {100*n/tot:5.1f}%   samples {100*smp[o]/max(st,1):5.1f}%  long_sb {lsb[o]}")


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 0)
