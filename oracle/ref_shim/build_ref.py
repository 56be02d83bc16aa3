#!/usr/bin/env python
"""Builds oracle/_ref/libartemis_ref.so: the reference's own hot-path sources, compiled where
they lie under /root/reference/src against the mock Parthenon in oracle/ref_shim/include.
Only runs where /root/reference is mounted (this container); the GPU box uses the prebuilt
.so that travels with the snapshot.  Flags mirror the reference's Release build
(-O3 -DNDEBUG -std=c++17 -fopenmp, no -march => no FMA contraction)."""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("ARTEMIS_REFERENCE", "/root/reference")
OUT = os.path.join(HERE, "..", "_ref")


def build(force=False):
    lib = os.path.join(OUT, "libartemis_ref.so")
    if not os.path.isdir(os.path.join(REF, "src")):
        return lib if os.path.exists(lib) else None
    srcs = [os.path.join(HERE, "ref_api.cpp"), os.path.join(HERE, "ref_refine.cpp")] + [
        os.path.join(dp, f) for dp, _, fs in os.walk(os.path.join(HERE, "include")) for f in fs]
    if (not force and os.path.exists(lib)
            and all(os.path.getmtime(lib) >= os.path.getmtime(s) for s in srcs)):
        return lib
    os.makedirs(OUT, exist_ok=True)
    refine = os.path.join(REF, "src", "utils", "refinement")
    cmd = ["g++", "-O3", "-DNDEBUG", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC",
           "-shared", "-w", "-I", os.path.join(HERE, "include"), "-I", os.path.join(REF, "src"),
           # the two multilevel operator headers, named explicitly (the include path resolves
           # "utils/refinement/*.hpp" to the mocks of the uniform-mesh translation unit)
           '-DAR_PROLONGATION_HPP="%s"' % os.path.join(refine, "prolongation.hpp"),
           '-DAR_RESTRICTION_HPP="%s"' % os.path.join(refine, "restriction.hpp"),
           os.path.join(HERE, "ref_api.cpp"), os.path.join(HERE, "ref_refine.cpp"), "-o", lib]
    subprocess.check_call(cmd)
    return lib


if __name__ == "__main__":
    print(build(force="-f" in sys.argv))
