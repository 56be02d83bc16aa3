#!/usr/bin/env python
"""Generates tests/golden/strat_bc_vectors.npz from the reference's OWN shearing-box boundary
functions (strat::ExtrapInnerX1 ... ExtrapOuterX3, src/pgen/strat.hpp:154-666, sliced and compiled
by oracle/ref_shim/strat/build_strat_ref.py into oracle/_ref/libstrat_ref.so).  Run in the
container that mounts /root/reference; the .npz is committed so the oracle stays pinned where the
reference tree is absent (tests/test_strat_bc_pin.py::test_oracle_reproduces_the_golden_vectors).

Each case: a seeded gas (6 entries) + dust (4 entries) block with ghost zones, the six faces
applied in Parthenon's order on fine arrays (case 0, 1) or on the coarse buffer (case 2)."""
import ctypes as C
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
CASES = [  # nx (fine interior), ng, coarse, block origin
    dict(nx=(6, 4, 6), ng=2, coarse=0, xmin=(-0.2, -0.5, -0.4)),
    dict(nx=(8, 8, 1), ng=2, coarse=0, xmin=(0.1, -0.5, 0.0)),
    dict(nx=(8, 8, 8), ng=2, coarse=1, xmin=(-0.3, 0.0, 0.2)),
]
Q, OM0, SD = 1.5, 0.7, 1
DXF = (0.05, 0.04, 0.03)


def geometry(case):
    nx, ng, c = case["nx"], case["ng"], 2 if case["coarse"] else 1
    ndim = 2 + (nx[2] > 1)
    n = [nx[d] // c + 2 * ng if d < ndim else 1 for d in range(3)]
    s = [ng if d < ndim else 0 for d in range(3)]
    e = [s[d] + (nx[d] // c if d < ndim else 1) - 1 for d in range(3)]
    dx = [DXF[d] * (c if (d == 0 or d < ndim) else 1) for d in range(3)]
    x0 = [case["xmin"][d] - s[d] * dx[d] for d in range(3)]
    return ndim, n, s, e, dx, x0


def state(seed, nvar, shape, S):
    rng = np.random.default_rng(seed)
    a = rng.standard_normal((nvar,) + shape)
    a[:S] = rng.uniform(0.5, 2.0, (S,) + shape)
    return a


if __name__ == "__main__":
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "build_strat_ref", os.path.join(ROOT, "oracle", "ref_shim", "strat", "build_strat_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    lib = mod.build()
    assert lib and os.path.exists(lib), "needs /root/reference (or a prebuilt oracle/_ref/libstrat_ref.so)"
    L = C.CDLL(lib)
    L.ar_strat_bc.restype = None
    DP, I3, D3 = C.POINTER(C.c_double), C.c_int * 3, C.c_double * 3
    out = {"q": Q, "om0": OM0, "ncases": len(CASES)}
    for c, case in enumerate(CASES):
        ndim, n, s, e, dx, x0 = geometry(case)
        shape = (n[2], n[1], n[0])
        gas, dust = state(100 + c, 6, shape, 1), state(200 + c, 4 * SD, shape, SD)
        out[f"gas_in_{c}"], out[f"dust_in_{c}"] = gas.copy(), dust.copy()
        for face in range(2 * ndim):
            L.ar_strat_bc(case["ng"], I3(*case["nx"]), D3(*x0), D3(*dx), case["coarse"], 1,
                          gas.ctypes.data_as(DP), SD, dust.ctypes.data_as(DP), face,
                          C.c_double(Q), C.c_double(OM0))
        out[f"gas_out_{c}"], out[f"dust_out_{c}"] = gas, dust
    path = os.path.join(HERE, "strat_bc_vectors.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))
