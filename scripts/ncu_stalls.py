#!/usr/bin/env python
"""Aggregate `ncu --page source --csv` output: stall reasons and the hottest SASS lines."""
import csv
import sys


def main(path, ntop=40):
    rows = list(csv.reader(open(path)))
    kernels, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1], "hdr": None, "data": []}
            kernels.append(cur)
        elif cur is not None and cur["hdr"] is None:
            cur["hdr"] = r
        elif cur is not None and len(r) == len(cur["hdr"]):
            cur["data"].append(r)
    for kn in kernels:
        hdr, data = kn["hdr"], kn["data"]
        ix = {h: i for i, h in enumerate(hdr)}
        tot = sum(int(r[ix["# Samples"]]) for r in data)
        inst = sum(int(r[ix["Instructions Executed"]]) for r in data)
        print("=" * 100)
        print(kn["name"][:100])
        print("samples", tot, "warp instructions", inst)
        stalls = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
        agg = {s: sum(int(r[ix[s]]) for r in data) for s in stalls}
        for k, v in sorted(agg.items(), key=lambda x: -x[1])[:9]:
            print(f"  {k:28s} {v:8d} {100 * v / tot:5.1f}%")
        ops = {}
        for r in data:
            op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
            if op.startswith("@"):
                op = r[ix["Source"]].split()[1]
            op = op.split(".")[0]
            o = ops.setdefault(op, [0, 0])
            o[0] += int(r[ix["Instructions Executed"]])
            o[1] += int(r[ix["# Samples"]])
        print("  opcode            executed   %inst  samples")
        for op, (n, s) in sorted(ops.items(), key=lambda x: -x[1][0])[:16]:
            print(f"  {op:14s} {n:12d} {100 * n / inst:6.1f}% {s:8d}")
        top = sorted(data, key=lambda r: -int(r[ix["# Samples"]]))[:ntop]
        for r in top:
            st = {s: int(r[ix[s]]) for s in stalls if int(r[ix[s]]) > 0}
            st = sorted(st.items(), key=lambda x: -x[1])[:3]
            print(r[ix["# Samples"]].rjust(6), r[ix["Instructions Executed"]].rjust(9),
                  r[ix["Source"]][:56].ljust(56), st)


if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 40)
