"""Shearing-box user boundary conditions (inputs/ssheet/ssheet.in, config 5): the
oracle's ao_strat_bc (oracle/artemis_oracle.c) against the reference's OWN six functions --
strat::ExtrapInnerX1 / ExtrapOuterX1 / ShearInnerX2 / ShearOuterX2 / ExtrapInnerX3 /
ExtrapOuterX3 are sliced out of src/pgen/strat.hpp:154-666 at build time (together with
strat::StratParams and parthenon::IndexShape) and compiled against a mock of the few types they
touch (oracle/ref_shim/strat).  Bit for bit, on random states with both signs of every velocity
(so each inflow clip fires and does not fire), blocks on either side of x1 = 0 (both branches of
the shear inflow), fine arrays and coarse buffers, gas and gas + dust, 2-D and 3-D, nghost 2 / 4,
applied face after face in Parthenon's order so corners inherit the earlier faces' fills."""
import ctypes as C
import os

import numpy as np
import pytest

from oracle import oracle_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "oracle", "_ref", "libstrat_ref.so")
needs_ref = pytest.mark.skipif(not os.path.exists(LIB), reason="oracle/_ref/libstrat_ref.so not built")
_DP = C.POINTER(C.c_double)
KIND = {0: 5, 1: 5, 2: 6, 3: 6, 4: 5, 5: 5}   # strat: extrap on x1 / x3, inflow on x2


def _state(rng, nvar, shape, S):
    a = rng.standard_normal((nvar,) + shape)
    a[:S] = rng.uniform(0.5, 2.0, (S,) + shape)     # densities > 0 (the x3 faces take a power)
    return a


@needs_ref
@pytest.mark.parametrize("coarse", [0, 1])
@pytest.mark.parametrize("ng", [2, 4])
@pytest.mark.parametrize("nx,xmin", [((8, 12, 1), (-0.3, -0.5, 0.0)), ((8, 6, 10), (-0.2, -0.5, -0.4)),
                                     ((12, 8, 8), (0.1, 0.0, 0.0)), ((8, 8, 8), (-1.0, 0.0, 0.2))])
@pytest.mark.parametrize("Sd", [0, 3])
def test_strat_user_bcs_equal_the_references_own_functions(nx, xmin, ng, coarse, Sd):
    L = C.CDLL(LIB)
    L.ar_strat_bc.restype = None
    ndim = 2 + (nx[2] > 1)
    dxf = [0.05, 0.04, 0.03]
    c = 2 if coarse else 1
    # the index space the arrays live in: fine block, or its coarse buffer
    n = [nx[d] // c + 2 * ng if d < ndim else 1 for d in range(3)]
    s = [ng if d < ndim else 0 for d in range(3)]
    e = [s[d] + (nx[d] // c if d < ndim else 1) - 1 for d in range(3)]
    dx = [dxf[d] * (c if (d == 0 or d < ndim) else 1) for d in range(3)]
    x0 = [xmin[d] - s[d] * dx[d] for d in range(3)]
    shape = (n[2], n[1], n[0])
    q, om0 = 1.5, 0.7
    rng = np.random.default_rng(7 + 13 * ng + coarse + Sd)
    gas = _state(rng, 6, shape, 1)
    dust = _state(rng, 4 * Sd, shape, Sd) if Sd else None
    want_g, want_d = gas.copy(), (dust.copy() if Sd else None)
    got_g, got_d = gas.copy(), (dust.copy() if Sd else None)
    I3, D3 = C.c_int * 3, C.c_double * 3
    for face in range(2 * ndim):
        L.ar_strat_bc(ng, I3(*nx), D3(*x0), D3(*dx), coarse, 1, want_g.ctypes.data_as(_DP), Sd,
                      want_d.ctypes.data_as(_DP) if Sd else None, face, C.c_double(q), C.c_double(om0))
        oracle_py.strat_bc(got_g, 0, x0, dx, s, e, 0, 1, face, KIND[face], q, om0)
        if Sd:
            oracle_py.strat_bc(got_d, 0, x0, dx, s, e, 1, Sd, face, KIND[face], q, om0)
    assert not np.array_equal(want_g, gas)
    assert np.array_equal(got_g, want_g)
    # the pressure (entry 4) is no FillGhost field: neither side touches it
    assert np.array_equal(want_g[4], gas[4])
    if Sd:
        assert np.array_equal(got_d, want_d)


def test_strat_registers_extrap_on_x1_x3_and_inflow_on_x2_only():
    a = np.ones((6, 6, 6, 6))
    for face, kind in ((2, 5), (3, 5), (0, 6), (1, 6), (4, 6), (5, 6)):
        with pytest.raises(ValueError):
            oracle_py.strat_bc(a, 0, (0, 0, 0), (1, 1, 1), (2, 2, 2), (3, 3, 3), 0, 1, face, kind)


def test_oracle_reproduces_the_golden_vectors():
    """tests/golden/strat_bc_vectors.npz = outputs of the reference's own functions
    (tests/golden/make_strat_bc_vectors.py): the pin that survives where neither
    /root/reference nor oracle/_ref exists"""
    import importlib.util
    here = os.path.dirname(os.path.abspath(__file__))
    spec = importlib.util.spec_from_file_location(
        "make_strat_bc_vectors", os.path.join(here, "golden", "make_strat_bc_vectors.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gold = np.load(os.path.join(here, "golden", "strat_bc_vectors.npz"))
    assert int(gold["ncases"]) == len(gen.CASES)
    for c, case in enumerate(gen.CASES):
        ndim, n, s, e, dx, x0 = gen.geometry(case)
        gas, dust = gold[f"gas_in_{c}"].copy(), gold[f"dust_in_{c}"].copy()
        for face in range(2 * ndim):
            oracle_py.strat_bc(gas, 0, x0, dx, s, e, 0, 1, face, KIND[face], float(gold["q"]),
                               float(gold["om0"]))
            oracle_py.strat_bc(dust, 0, x0, dx, s, e, 1, gen.SD, face, KIND[face], float(gold["q"]),
                               float(gold["om0"]))
        assert np.array_equal(gas, gold[f"gas_out_{c}"])
        assert np.array_equal(dust, gold[f"dust_out_{c}"])
