"""Experiment builds: recompile ONLY the Cartesian fused unit (fused.cu, AB_GEOM=0) of the
default build with extra nvcc flags and link it with the cached objects of the default build into
artemis_b200/lib/libartemis_b200_<name>.so (select with AB200_VARIANT=<name>).
usage: python scripts/build_exp.py <name> [nvcc flags...]"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from artemis_b200 import build as B  # noqa: E402

name, extra = sys.argv[1], sys.argv[2:]
fast = os.path.join(B.OBJDIR, "fast")
out_dir = os.path.join("/tmp/ab200_exp", name)
os.makedirs(out_dir, exist_ok=True)
obj = os.path.join(out_dir, "fused_g0.o")
t0 = time.time()
cmd = [B.NVCC, *B.ARCH, *B.COMMON, "-DAB200_FAST_MATH", "-DAB_GEOM=0", *extra, "-Xptxas", "-v", "-c",
       os.path.join(B.CSRC, "fused.cu"), "-o", obj]
r = subprocess.run(cmd, capture_output=True, text=True)
open(os.path.join(out_dir, "ptxas.log"), "w").write(r.stderr)
if r.returncode:
    sys.exit(r.stderr[-3000:])
objs = [os.path.join(fast, f) for f in sorted(os.listdir(fast)) if f.endswith(".o") and f != "fused_g0.o"]
lib = os.path.join(B.LIBDIR, f"libartemis_b200_{name}.so")
subprocess.check_call([B.NVCC, *B.ARCH, "-shared", "-o", lib, obj, *objs, "-lcudart", "-ldl"])
print(lib, f"{time.time() - t0:.0f}s")
