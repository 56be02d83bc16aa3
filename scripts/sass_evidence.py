#!/usr/bin/env python
"""SASS evidence of the hot kernels in the BUILT library (no GPU needed): per kernel the count
of TMA (UTMALDG / UTMAPF), mbarrier (SYNCS), named-barrier (BAR), fp64 and MUFU instructions, plus a
short excerpt around the first TMA instruction.  usage: sass_evidence.py [libartemis_b200.so]"""
import collections
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else "artemis_b200/lib/libartemis_b200.so"
KERNELS = [  # config-2 instantiations (gas, HLLC, PPM, Cartesian)
    ("k_xchunk_pass  (x1 pass, default path)", r"_ZN5ab20013k_xchunk_passILi0ELi0ELi0ELi2EEE"),
    ("k_march_pass<2> (x2 pass, default path)", r"_ZN5ab20012k_march_passILi0ELi0ELi0ELi2ELi2ELb0EEE"),
    ("k_march_pass<3,LAST> (x3 pass + C2P/P2C/dt, default path)", r"_ZN5ab20012k_march_passILi0ELi0ELi0ELi2ELi3ELb1EEE"),
    ("k_sweep_stage MODE 0 (single-pass stage, one role)", r"_ZN5ab20013k_sweep_stageILi0ELi0ELi2ELi0EEE"),
    ("k_trio_stage MODE 0 (single-pass stage, warp-specialised)", r"_ZN5ab20012k_trio_stageILi0ELi0ELi2ELi0EEE"),
    ("k_fill_ghosts (ghost fill)", r"_ZN5ab20013k_fill_ghostsILi0ELi0ELb0EEE"),
]
names = subprocess.check_output(["cuobjdump", "-sass", lib, "-lelf"], text=True, stderr=subprocess.DEVNULL) \
    if False else ""
sym = subprocess.check_output("cuobjdump -elf %s | grep -oE '_ZN5ab200[A-Za-z0-9_]+' | sort -u" % lib,
                              shell=True, text=True).split()
for label, pat in KERNELS:
    full = [s for s in sym if re.match(pat, s)]
    if not full:
        print(f"== {label}: not found ({pat})")
        continue
    sass = subprocess.check_output(["cuobjdump", "-sass", "-fun", full[0], lib], text=True)
    ops = re.findall(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", sass)
    c = collections.Counter(o.split(".")[0] for o in ops)
    print(f"== {label}\n   {full[0]}")
    print("   instructions %d | UTMALDG %d UTMAPF %d | SYNCS (mbarrier) %d | BAR %d | LDGSTS (cp.async) %d | "
          "DFMA+DMUL+DADD %d DSETP %d FSEL %d MUFU %d | LDS %d STS %d LDG %d STG %d LDL+STL %d" % (
              len(ops), c["UTMALDG"], c["UTMAPF"], c["SYNCS"], c["BAR"], c["LDGSTS"],
              c["DFMA"] + c["DMUL"] + c["DADD"], c["DSETP"], c["FSEL"], c["MUFU"], c["LDS"], c["STS"],
              c["LDG"], c["STG"], c["LDL"] + c["STL"]))
    lines = sass.splitlines()
    for i, l in enumerate(lines):
        if "UTMALDG" in l:
            print("   first TMA load in context:")
            for x in lines[max(0, i - 3):i + 4]:
                m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(.*?);", x)
                if m:
                    print("      " + m.group(1).strip())
            break
