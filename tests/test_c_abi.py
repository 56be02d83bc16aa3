"""The drop-in boundary is a C ABI: a plain C99 translation unit (tests/c/abi_roundtrip.c)
compiled against include/ab200.h and LINKED to libartemis_b200.so binds a 2-block mesh, runs
one rk2 cycle (ab200_fused_stage + ab200_fill_ghosts + the device dt bookkeeping) and prints a
checksum of the raw result bits; the ctypes path must produce the same bits.  ("no
C/C++ TU ever links and calls the .so".)"""
import math
import os
import re
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "c", "abi_roundtrip.c")
LIBDIR = os.path.join(ROOT, "artemis_b200", "lib")


def _build(tmp_path):
    exe = str(tmp_path / "abi_roundtrip")
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-Werror", "-I", os.path.join(ROOT, "include"), SRC,
           "-L", LIBDIR, "-lartemis_b200", f"-Wl,-rpath,{LIBDIR}", "-lm", "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_c_program_compiles_links_and_fails_loudly_without_a_gpu(tmp_path):
    """CPU half: the TU compiles with -Wall -Wextra -Werror as C99, every symbol it uses
    resolves at link time, the library loads, and with no device ab200_create reports
    AB200_ECUDA (exit 77) instead of falling back to anything."""
    from artemis_b200 import build as b
    b.build_variant("fast", ["-DAB200_FAST_MATH"])
    exe = _build(tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True)
    assert r.returncode in (0, 77), (r.returncode, r.stdout, r.stderr)
    assert re.search(r"abi_version (\d+)", r.stdout)
    if r.returncode == 77:
        assert "no CUDA device" in r.stdout and "no CPU fallback" in r.stdout


def test_glue_tu_compiles():
    """tests/c/ab200_glue.cpp -- the file INTEGRATION.md tells a maintainer to add to Artemis --
    compiles against include/ab200.h and a mock that declares exactly the public Parthenon
    accessors it uses (tests/c/parthenon_glue_mock.hpp, each cited to its reference file:line)."""
    cmd = ["g++", "-std=c++17", "-fsyntax-only", "-Wall", "-Wextra", "-Werror",
           "-DAB200_GLUE_SYNTAX_CHECK", "-I", os.path.join(ROOT, "include"),
           "-I", os.path.join(ROOT, "tests", "c"), os.path.join(ROOT, "tests", "c", "ab200_glue.cpp")]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr


def _ctypes_twin(path):
    """Same mesh, state and call sequence as tests/c/abi_roundtrip.c, through ctypes."""
    from artemis_b200.enums import Coordinates, Fluid, ReconstructionMethod, RSolver
    from artemis_b200.mesh import UniformMesh
    from artemis_b200.meshdata import MeshData
    from artemis_b200.params import FluidParams
    NB, NX, NG = 2, (8, 6, 4), 4
    mesh = UniformMesh(nx=(16, 6, 4), xmin=(0, 0, 0), xmax=(1.6, 0.6, 0.4), block_nx=NX, nghost=NG)
    dx = [0.8 / 8, 0.6 / 6, 0.4 / 4]
    for b in range(NB):   # the C program's per-block UniformCartesian, bit for bit
        mesh.blk_dx[b] = dx
        mesh.blk_xmin[b] = [0.8 * b - NG * dx[0], -NG * dx[1], -NG * dx[2]]
    gamma = 1.4
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod.ppm, RSolver.hllc,
                     cfl=0.3, nspecies=1, dfloor=1e-10, gamma=gamma, siefloor=1e-10)
    ni, nj, nk = mesh.ni, mesh.nj, mesh.nk
    prim = np.zeros((NB, 6, nk, nj, ni))
    PI2 = 6.283185307179586
    for b in range(NB):
        xm = mesh.blk_xmin[b]
        for k in range(nk):
            for j in range(nj):
                for i in range(ni):
                    x = xm[0] + (i + 0.5) * dx[0]
                    y = xm[1] + (j + 0.5) * dx[1]
                    z = xm[2] + (k + 0.5) * dx[2]
                    s = math.sin(PI2 * x / 1.6) * math.cos(PI2 * y / 0.6)
                    c = math.cos(PI2 * z / 0.4)
                    prim[b, 0, k, j, i] = 1.0 + 0.2 * s * c
                    prim[b, 1, k, j, i] = 0.3 * c
                    prim[b, 2, k, j, i] = -0.2 * s
                    prim[b, 3, k, j, i] = 0.1 * s * c
                    prim[b, 5, k, j, i] = 1.5 + 0.3 * s
                    prim[b, 4, k, j, i] = (gamma - 1.0) * prim[b, 0, k, j, i] * prim[b, 5, k, j, i]
    md = MeshData(mesh, gas=gp, materialize_fluxes=False)
    md.set_stage_path(path)
    md.gas.prim.set(prim)
    big = float(np.finfo(np.float64).max)
    md.call("ab200_prim_to_cons")
    md.call("ab200_cons_to_prim")
    md.call("ab200_fill_ghosts")
    md.set_time_state(big)
    md.call("ab200_estimate_timestep_device")
    md.call("ab200_set_global_timestep_device", big, 0)
    for s, (g0, g1, be) in enumerate(((0.0, 1.0, 1.0), (0.5, 0.5, 0.5))):
        md.call("ab200_fused_stage", g0, g1, be, 0.0, 0, int(s == 0), 1 | 4 | (2 if s == 1 else 0))
        md.call("ab200_fill_ghosts")
    md.call("ab200_set_global_timestep_device", big, 1)
    md.call("ab200_sync_prim")
    ts = md.time_state()
    h = 1469598103934665603
    for a in (md.gas.u0.get(), md.gas.prim.get()):
        for byte in a.tobytes():
            h = ((h ^ byte) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    md.close()
    return h, ts


@pytest.mark.gpu
@pytest.mark.parametrize("path", ["three_pass", "single_pass", "role_split"])
def test_c_host_and_ctypes_host_produce_the_same_bits(tmp_path, path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, path], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, (r.stdout, r.stderr)
    out = dict(line.split(" ", 1) for line in r.stdout.strip().splitlines())
    assert int(out["launches"]) > 0
    code = {"three_pass": 1, "single_pass": 2, "role_split": 3}[path]
    assert int(out["stage_path"]) == code
    h, ts = _ctypes_twin(path)
    assert out["checksum"] == f"{h:016x}"
    m = re.match(r"(\d+) time (\S+) dt (\S+)", out["ncycle"])
    assert int(m.group(1)) == 1 and float(m.group(2)) == ts[2] and float(m.group(3)) == ts[0]
