"""Builds libartemis_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

Two variants are produced from the same sources:
  * artemis_b200/lib/libartemis_b200.so         -- default (FMA contraction on)
  * artemis_b200/lib/libartemis_b200_strict.so  -- --fmad=false, operation-for-operation
    identical to the reference's CPU arithmetic (used by the bit-exactness tests)
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIBDIR = os.path.join(HERE, "lib")
OBJDIR = os.environ.get("AB200_OBJDIR", "/tmp/ab200_build_objs")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
COMMON = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "-Xcompiler", "-O2",
          "--expt-relaxed-constexpr", "-I", os.path.join(HERE, "..", "include")]

# (source, extra defines, object suffix)
UNITS = [("api.cu", [], ""), ("tasks.cu", [], ""), ("halo.cu", [], ""),
         ("fused_dispatch.cu", [], ""), ("host_path.cu", [], ""), ("tma_maps.cu", [], "")]
UNITS += [("sweep_host.cu", [], ""), ("refine.cu", [], ""), ("comm.cu", [], ""), ("sources.cu", [], ""), ("history.cu", [], ""), ("diffusion.cu", [], ""), ("multilevel.cu", [], "")]
UNITS += [("sweep.cu", [f"-DAB_RS={r}"], f"_r{r}") for r in range(3)]
UNITS += [("trio.cu", [f"-DAB_RS={r}"], f"_r{r}") for r in range(3)]
UNITS += [("fused.cu", [f"-DAB_GEOM={g}"], f"_g{g}") for g in range(6)]
UNITS += [("tasks_flux.cu", [f"-DAB_GEOM={g}"], f"_g{g}") for g in range(6)]


def _deps():
    return [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [
        os.path.join(HERE, "..", "include", "ab200.h")]


def _stamp(flags):
    h = hashlib.sha256()
    for p in _deps():
        with open(p, "rb") as fh:
            h.update(fh.read())
    h.update(" ".join(flags).encode())
    return h.hexdigest()


def _closure(path, seen=None):
    """Transitive set of quoted #include files of a source (for per-object rebuild stamps)."""
    import re
    seen = seen if seen is not None else set()
    if path in seen or not os.path.exists(path):
        return seen
    seen.add(path)
    with open(path) as fh:
        for inc in re.findall(r'#include\s+"([^"]+)"', fh.read()):
            _closure(os.path.normpath(os.path.join(os.path.dirname(path), inc)), seen)
    return seen


def _compile(args):
    src, defs, suffix, variant, flags = args
    obj = os.path.join(OBJDIR, variant, os.path.splitext(src)[0] + suffix + ".o")
    os.makedirs(os.path.dirname(obj), exist_ok=True)
    extra = os.environ.get("AB200_EXTRA_NVCC_FLAGS", "").split()
    cmd = [NVCC, *ARCH, *COMMON, *flags, *extra, *defs, "-c", os.path.join(CSRC, src), "-o", obj]
    h = hashlib.sha256(" ".join(cmd).encode())
    for dep in sorted(_closure(os.path.join(CSRC, src))):
        with open(dep, "rb") as fh:
            h.update(fh.read())
    stamp = h.hexdigest()
    if os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == stamp:
        return obj, ""
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed for {src}{suffix} [{variant}]:\n{r.stdout}\n{r.stderr}")
    with open(obj + ".stamp", "w") as fh:
        fh.write(stamp)
    return obj, r.stderr


def build_variant(variant: str, flags, verbose=False, jobs=None):
    out = os.path.join(LIBDIR, f"libartemis_b200{'' if variant == 'fast' else '_' + variant}.so")
    stamp_file = out + ".stamp"
    stamp = _stamp(flags + os.environ.get("AB200_EXTRA_NVCC_FLAGS", "").split())
    if os.path.exists(out) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return out
    os.makedirs(LIBDIR, exist_ok=True)
    units = [u for u in UNITS if os.path.exists(os.path.join(CSRC, u[0]))]
    jobs = jobs or min(len(units), os.cpu_count() or 4)
    with cf.ThreadPoolExecutor(jobs) as ex:
        res = list(ex.map(_compile, [(s, d, x, variant, flags) for s, d, x in units]))
    if verbose:
        for _, err in res:
            if err.strip():
                print(err, file=sys.stderr)
    objs = [o for o, _ in res]
    cmd = [NVCC, *ARCH, "-shared", "-o", out, *objs, "-lcudart", "-ldl"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return out


def build_all(verbose=False):
    fast = build_variant("fast", ["-DAB200_FAST_MATH"], verbose)
    if os.environ.get("AB200_BUILD_ONLY_FAST"):
        return fast, None
    strict = build_variant("strict", ["--fmad=false"], verbose)
    return fast, strict


if __name__ == "__main__":
    print(build_all(verbose="-v" in sys.argv))
