#!/usr/bin/env python
"""Instruction histogram of the kernels in a .cu / .cubin / .so (no GPU needed).

  python scripts/sass_count.py scripts/probes/riemann_probe.cu [-DAB200_FAST_MATH ...] [-k regex]

Compiles a .cu for sm_100a (same flags as artemis_b200/build.py), disassembles with cuobjdump
and prints, per kernel whose name matches -k, the static instruction count by mnemonic group.
Static counts of straight-line device functions (Riemann solvers, limiters) are their dynamic
cost per call; for kernels with loops pass --loop to count only the largest backward-branch
body."""
import collections
import os
import re
import subprocess
import sys
import tempfile


def disasm(path, flags):
    if path.endswith(".cu"):
        cub = os.path.join(tempfile.mkdtemp(), "probe.cubin")
        cmd = ["nvcc", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
               "--expt-relaxed-constexpr", "-I", "include", "-cubin", "-o", cub, path] + flags
        subprocess.check_call(cmd)
        path = cub
    return subprocess.check_output(["cuobjdump", "-sass", path], text=True)


def kernels(sass):
    cur, out = None, collections.OrderedDict()
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            out[cur] = []
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(@!?U?P\d+\s+)?([A-Z0-9_.]+)(.*?);", line)
        if m and cur:
            out[cur].append((int(m.group(1), 16), m.group(3), m.group(4)))
    return out


GROUPS = [("fp64 arith", r"^(DFMA|DMUL|DADD)"), ("fp64 cmp/minmax", r"^(DSETP|DMNMX)"),
          ("mufu", r"^MUFU"), ("select", r"^(FSEL|SEL)"), ("ld/st global", r"^(LDG|STG|LD\.|ST\.)"),
          ("ld/st shared", r"^(LDS|STS|LDSM)"), ("int/addr", r"^(IMAD|IADD|LEA|SHF|LOP|ISETP|IABS|I2F|F2I|PRMT|MOV|UMOV|UIADD|ULEA|UIMAD|R2UR|S2R|S2UR|CS2R|ULOP|USHF|UISETP|USEL|VIADD|PLOP|UPLOP|R2P|P2R)"),
          ("branch/sync", r"^(BRA|BSSY|BSYNC|EXIT|CALL|RET|BAR|WARPSYNC|NANOSLEEP|SYNCS|BMOV|YIELD|NOP|DEPBAR|ERRBAR|MEMBAR|FENCE|CCTL)")]


def main():
    args = sys.argv[1:]
    path = args[0]
    kre, flags, loop = ".", [], False
    it = iter(args[1:])
    for a in it:
        if a == "-k":
            kre = next(it)
        elif a == "--loop":
            loop = True
        else:
            flags.append(a)
    for name, ins in kernels(disasm(path, flags)).items():
        dem = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
        if not re.search(kre, dem):
            continue
        if loop:  # largest backward branch body
            best = (0, 0)
            for addr, op, rest in ins:
                if op.startswith("BRA"):
                    m = re.search(r"0x([0-9a-f]+)", rest)
                    if m and int(m.group(1), 16) < addr and addr - int(m.group(1), 16) > best[1] - best[0]:
                        best = (int(m.group(1), 16), addr)
            ins = [i for i in ins if best[0] <= i[0] <= best[1]]
        hist = collections.Counter()
        for _, op, _ in ins:
            for g, rx in GROUPS:
                if re.match(rx, op):
                    hist[g] += 1
                    break
            else:
                hist["other:" + op.split(".")[0]] += 1
        print(f"{dem[:110]}\n   total {len(ins)}  " + "  ".join(f"{k}={v}" for k, v in hist.most_common()))


if __name__ == "__main__":
    main()
