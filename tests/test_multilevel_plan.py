"""Multilevel ghost-exchange bookkeeping:
  * `calc_indices` against parthenon::CalcIndices ITSELF -- the function is sliced out of
    external/parthenon/src/bvals/comms/bnd_info.cpp:105-252 at build time and compiled against a
    mock of the few types it touches (oracle/ref_shim/bnd) -- over every offset, level relation,
    block parity, index-range type and prolongation / restriction flag;
  * the neighbour rule of Tree::FindNeighborsImpl on refined lattices (each ghost region is
    claimed by exactly one neighbour, messages pair up);
  * the whole exchange (restrict -> copy -> restrict -> coarse BCs -> prolongate -> fine BCs) on
    the CPU executor: a globally linear field must be reproduced in EVERY ghost zone (same-level
    copies are exact, volume-weighted restriction of a linear Cartesian field is the value at
    the coarse centroid, minmod prolongation of a linear field is exact), and no ghost zone may
    be left unfilled."""
import ctypes as C
import itertools
import os

import numpy as np
import pytest

from artemis_b200.enums import BoundaryFlag
from artemis_b200.multilevel import (BOUNDARY_EXTERIOR_RECV, BOUNDARY_INTERIOR_SEND, INTERIOR_RECV,
                                     INTERIOR_SEND, MultilevelMesh, calc_indices, exchange_plan)
from oracle import multilevel_py

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BND_LIB = os.path.join(ROOT, "oracle", "_ref", "libbnd_ref.so")
needs_bnd = pytest.mark.skipif(not os.path.exists(BND_LIB), reason="oracle/_ref/libbnd_ref.so not built")


@needs_bnd
@pytest.mark.parametrize("nx,ng", [((8, 8, 8), 4), ((16, 8, 4), 2), ((8, 6, 1), 2), ((12, 1, 1), 4)])
def test_calc_indices_equals_parthenons_own_function(nx, ng):
    L = C.CDLL(BND_LIB)
    ndim = sum(n > 1 for n in nx)
    offs = [(-1, 0, 1) if d < ndim else (0,) for d in range(3)]
    n = 0
    for o in itertools.product(*offs):
        for dlev in (-1, 0, 1):                      # neighbour coarser / same / finer
            for my_par in itertools.product(*[(0, 1) if d < ndim else (0,) for d in range(3)]):
                for nb_par in itertools.product(*[(0, 1) if d < ndim else (0,) for d in range(3)]):
                    my_l = tuple(4 + p for p in my_par)
                    nb_l = tuple(6 + p for p in nb_par)
                    for ir in (BOUNDARY_INTERIOR_SEND, BOUNDARY_EXTERIOR_RECV, INTERIOR_SEND, INTERIOR_RECV):
                        for prores in (0, 1):
                            # cell-centred field, and the flux of one on the face normal to a
                            # face offset (GetFluxCorrectionElements, bnd_info.cpp:72-84)
                            els = [0]
                            if sum(abs(v) for v in o) == 1:
                                els.append([abs(v) for v in o].index(1) + 1)
                            for el in els:
                                out = (C.c_int * 6)()
                                L.ar_calc_indices(ng, (C.c_int * 3)(*nx), 2, (C.c_longlong * 3)(*my_l),
                                                  2 + dlev, (C.c_longlong * 3)(*nb_l),
                                                  (C.c_int * 3)(*o), ir, prores, el, out)
                                want = tuple((out[2 * d], out[2 * d + 1]) for d in range(3))
                                got = calc_indices(ng, nx, 2, my_l, 2 + dlev, nb_l, o, ir,
                                                   bool(prores), flux_el=el)
                                assert got == want, (o, dlev, my_l, nb_l, ir, prores, el)
                                n += 1
    assert n >= 288


def _mesh(ndim, bcs, refine, root=(4, 4, 4), bnx=(8, 8, 8), ng=4):
    rb = tuple(root[d] if d < ndim else 1 for d in range(3))
    bn = tuple(bnx[d] if d < ndim else 1 for d in range(3))
    ref = tuple(tuple(r[d] if d < ndim else 0 for d in range(3)) for r in refine)
    return MultilevelMesh(root_blocks=rb, block_nx=bn, xmin=(0.0, -1.0, 2.0), xmax=(2.0, 1.5, 4.0),
                          refine=ref, nghost=ng, bcs=bcs)


PER = (BoundaryFlag.periodic,) * 6
OUT = (BoundaryFlag.outflow,) * 6
CASES = [
    (3, PER, [(1, 1, 1)]),                                  # one refined root block
    (3, PER, [(1, 1, 1), (2, 1, 1), (1, 2, 1), (2, 2, 2)]),  # an L-shaped refined region
    (3, OUT, [(0, 0, 0), (3, 3, 3), (1, 2, 0)]),            # refined blocks on physical boundaries
    (2, PER, [(1, 1, 0), (2, 2, 0)]),                       # corner-touching refined blocks
    (2, OUT, [(0, 1, 0), (3, 3, 0)]),
    (1, OUT, [(1, 0, 0), (3, 0, 0)]),
]


@pytest.mark.parametrize("ndim,bcs,refine", CASES)
def test_neighbour_lists_pair_up(ndim, bcs, refine):
    m = _mesh(ndim, bcs, refine)
    assert m.nb == np.prod(m.root_blocks) + len(refine) * (2 ** ndim - 1)
    for b in range(m.nb):
        seen = set()
        for nb in m.neighbors[b]:
            assert abs(nb.level - m.leaves[b][0]) <= 1          # 2:1 nesting
            key = (nb.gid, nb.offsets)
            assert key not in seen                               # no duplicates
            seen.add(key)
            back = [q for q in m.neighbors[nb.gid]
                    if q.gid == b and q.offsets == tuple(-o for o in nb.offsets)]
            assert len(back) == 1
    plan = exchange_plan(m)
    assert len(plan.copies) == sum(len(v) for v in m.neighbors)
    if refine:
        assert plan.prolongate and plan.restrict_send and plan.restrict_set


def _linear(mesh, b, dims, coef):
    """cell-centre values of a + c.x on the fine array of block b (or its coarse buffer)"""
    shape = tuple(dims[::-1])
    out = np.full(shape, coef[0])
    return out


def _centres(lo, dx, n, g):
    return lo + (np.arange(n) - g + 0.5) * dx


@pytest.mark.parametrize("ndim,bcs,refine", CASES)
def test_linear_field_is_reproduced_in_every_ghost_zone(ndim, bcs, refine):
    m = _mesh(ndim, bcs, refine)
    plan = exchange_plan(m)
    nvar = 3
    coef = np.array([[0.7, 0.3, -0.2, 0.5], [-1.1, 0.05, 0.4, -0.3], [2.0, -0.6, 0.1, 0.25]])
    if bcs[0] == BoundaryFlag.periodic:      # periodic + linear: use a constant so wrapping is exact
        coef[:, 1:] = 0.0
    fine = np.full(m.shape(nvar), np.nan)
    coarse = np.full(m.coarse_shape(nvar), np.nan)
    exact = np.zeros(m.shape(nvar))
    for b in range(m.nb):
        x = [_centres(m.blk_lo[b, d], m.blk_dx[b, d], (m.ni, m.nj, m.nk)[d], m.ngd[d]) for d in range(3)]
        X, Y, Z = x[0][None, None, :], x[1][None, :, None], x[2][:, None, None]
        for v in range(nvar):
            exact[b, v] = coef[v, 0] + coef[v, 1] * X + coef[v, 2] * Y + coef[v, 3] * Z
        fine[(b, slice(None)) + m.interior()] = exact[(b, slice(None)) + m.interior()]
    kinds = ["periodic" if f == BoundaryFlag.periodic else "outflow" for f in bcs]
    multilevel_py.run_plan(m, plan, fine, coarse, range(nvar), [0] * nvar, kinds)
    assert not np.isnan(fine).any(), "a ghost zone was never filled"
    # ghost zones inside the domain carry the linear field; those beyond an outflow boundary
    # hold the value of the last interior zone along the face normal
    for b in range(m.nb):
        faces = m.physical_faces(b)
        idx = [np.arange(n) for n in (m.ni, m.nj, m.nk)]
        clip = []
        for d in range(3):
            lo = m.ngd[d] if 2 * d in faces else 0
            hi = m.ngd[d] + m.block_nx[d] - 1 if 2 * d + 1 in faces else idx[d][-1]
            clip.append(np.clip(idx[d], lo, hi))
        want = exact[b][:, clip[2][:, None, None], clip[1][None, :, None], clip[0][None, None, :]]
        diff = np.abs(fine[b] - want)
        if any(nb.level < m.leaves[b][0] for nb in m.neighbors[b]):
            # a prolongated zone whose coarse parent touches an outflow boundary sees a zero
            # gradient on one side, so minmod drops the transverse slope there (by design):
            # leave the first / last coarse cell along a physical face out of the exactness check
            for d in range(3):
                sl = [slice(None)] * 4
                if 2 * d in faces:
                    sl[3 - d] = slice(0, m.ngd[d] + 2)
                    diff[tuple(sl)] = 0.0
                if 2 * d + 1 in faces:
                    sl[3 - d] = slice(m.ngd[d] + m.block_nx[d] - 2, None)
                    diff[tuple(sl)] = 0.0
        assert diff.max() <= 1e-13, (b, m.leaves[b], diff.max())


# ---- flux correction (AddFluxCorrectionTasks, boundary_communication.cpp:454-461) ---------------
def _total(mesh, u, var=0):
    """volume integral of one conserved variable over all leaf blocks (Cartesian)"""
    vol = np.prod(mesh.blk_dx[:, :mesh.ndim], axis=1)
    return float(sum(vol[b] * u[(b, var) + mesh.interior()].sum() for b in range(mesh.nb)))


@pytest.mark.parametrize("ndim,refine", [(3, [(1, 1, 1), (2, 1, 1)]), (2, [(1, 1, 0), (2, 2, 0)]),
                                         (1, [(1, 0, 0)])])
def test_flux_correction_makes_the_refined_mesh_conservative(ndim, refine):
    """periodic refined mesh, random gas + dust state with jumps: with the fine fluxes restricted
    onto every shared coarse face the volume integrals of mass, momentum and energy are
    conserved to rounding over whole rk2 cycles; without the correction they are not"""
    from artemis_b200.enums import Coordinates
    from oracle.oracle_py import OracleSim
    from tests.helpers import dust_params, gas_params, random_prim
    drift = {}
    for fc in (True, False):
        m = _mesh(ndim, PER, refine)
        gp = gas_params(Coordinates.cartesian, "plm", "hlle")
        dp = dust_params(Coordinates.cartesian, "plm", "hlle", S=1)
        sim = OracleSim(m, gas=gp, dust=dp)
        sim.flux_correction = fc
        sim.gas.prim[:] = random_prim(m, gp, seed=5)
        sim.dust.prim[:] = random_prim(m, dp, seed=6)
        sim.nlim = 3
        sim.initialize()
        # mass, momentum, total energy (the gas's internal-energy entry is not a conservation law)
        nv = [5, dp.nvar]
        before = [[_total(m, f.u0, v) for v in range(n)] for f, n in zip(sim.fluids, nv)]
        scale = [[_total(m, np.abs(f.u0), v) for v in range(n)] for f, n in zip(sim.fluids, nv)]
        sim.run()
        after = [[_total(m, f.u0, v) for v in range(n)] for f, n in zip(sim.fluids, nv)]
        drift[fc] = max(abs(a - b) / s for fa, fb, fs in zip(after, before, scale)
                        for a, b, s in zip(fa, fb, fs))
    assert drift[True] <= 5e-15, drift
    assert drift[False] >= 1e-8, drift


@pytest.mark.parametrize("ndim,refine", [(3, [(1, 1, 1)]), (2, [(1, 1, 0), (2, 2, 0)])])
def test_flux_correction_covers_the_diffusion_fluxes(ndim, refine):
    """gas.diff.momentum / gas.diff.energy carry Metadata::WithFluxes (src/gas/gas.cpp:277-285),
    so AddFluxCorrectionTasks restricts them onto the shared coarse faces as well: with viscosity
    and conduction on, a periodic refined mesh conserves momentum and total energy to rounding
    only if the diffusion fluxes are corrected together with the hydrodynamic ones"""
    from artemis_b200.enums import Coordinates
    from oracle import multilevel_py
    from oracle.oracle_py import OracleSim, make_diffusion
    from tests.helpers import gas_params, random_prim
    drift = {}
    for with_diff_fc in (True, False):
        m = _mesh(ndim, PER, refine)
        gp = gas_params(Coordinates.cartesian, "plm", "hlle")
        sim = OracleSim(m, gas=gp)
        sim.diffusion = make_diffusion(visc=("constant", 2e-2), cond=("conductivity", 3e-2))
        sim.gas.prim[:] = random_prim(m, gp, seed=9)
        sim.nlim = 3
        sim.initialize()
        if not with_diff_fc:   # hydrodynamic fluxes corrected, diffusion fluxes left alone
            orig = multilevel_py.flux_correct
            multilevel_py.flux_correct = (
                lambda mesh, plan, flux, _o=orig: None if flux is sim.dflx else _o(mesh, plan, flux))
        try:
            before = [_total(m, sim.gas.u0, v) for v in range(5)]
            scale = [_total(m, np.abs(sim.gas.u0), v) for v in range(5)]
            sim.run()
        finally:
            if not with_diff_fc:
                multilevel_py.flux_correct = orig
        after = [_total(m, sim.gas.u0, v) for v in range(5)]
        drift[with_diff_fc] = max(abs(a - b) / s for a, b, s in zip(after, before, scale))
    assert drift[True] <= 5e-15, drift
    assert drift[False] >= 1e-10, drift


def test_flux_correction_plan_covers_every_fine_coarse_face_once():
    m = _mesh(3, PER, [(1, 1, 1), (2, 1, 1), (1, 2, 2)])
    from artemis_b200.multilevel import flux_correction_plan
    plan = flux_correction_plan(m)
    # every fine block has 3 outward faces of its parent; those shared with another refined
    # root block are same-level and need no correction
    n_face = sum(1 for b in range(m.nb) for nb in m.neighbors[b]
                 if nb.level == m.leaves[b][0] - 1 and sum(map(abs, nb.offsets)) == 1)
    assert len(plan) == n_face > 0
    h = [m.block_nx[d] // 2 for d in range(3)]
    seen = set()
    for fb, cb, d, rbox, dbox in plan:
        ext = [rbox[q][1] - rbox[q][0] + 1 for q in range(3)]
        assert ext == [1 if q == d else h[q] for q in range(3)]          # one coarse face layer
        assert [dbox[q][1] - dbox[q][0] + 1 for q in range(3)] == ext
        key = (cb, d) + tuple(dbox[q][0] for q in range(3))
        assert key not in seen                                           # written exactly once
        seen.add(key)
