"""Multilevel operators: the oracle's restatement of
ArtemisUtils::RestrictAverage<GEOM> / ProlongateSharedMinMod<GEOM> against the REFERENCE'S OWN
headers (src/utils/refinement/{restriction,prolongation}.hpp compiled into oracle/_ref), bit
for bit, on every geometry; plus the properties the operators must have (conservation of the
volume integral, exactness on constants, restriction o prolongation = identity)."""
import numpy as np
import pytest

from artemis_b200.enums import BoundaryFlag, Coordinates
from oracle import oracle_py, ref_py
from tests.helpers import GEOM_DOMAINS, make_mesh

ALL = [Coordinates.cartesian, Coordinates.cylindrical, Coordinates.axisymmetric,
       Coordinates.spherical1D, Coordinates.spherical2D, Coordinates.spherical3D]


def _setup(coords, seed=0, nvar=3, bnx=(8, 6, 4)):
    mesh = make_mesh(coords, 3, nblk=(2, 1, 1), bnx=bnx, bcs=(BoundaryFlag.outflow,) * 6)
    r = oracle_py.refine_geom(mesh, b=1)
    rng = np.random.default_rng(seed)
    fine = 1.0 + rng.random((nvar, mesh.nk, mesh.nj, mesh.ni))
    coarse = 1.0 + rng.random((nvar, r.cnk, r.cnj, r.cni))
    return mesh, r, fine, coarse


def _interior_box(r, mesh, grow=0):
    """coarse interior (+grow ghost cells in every active direction)"""
    box = []
    for d, (cs, n) in enumerate(((r.cib_s, mesh.block_nx[0]), (r.cjb_s, mesh.block_nx[1]),
                                 (r.ckb_s, mesh.block_nx[2]))):
        if d < mesh.ndim:
            box += [cs - grow, cs + n // 2 - 1 + grow]
        else:
            box += [0, 0]
    return box


needs_ref = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("coords", ALL)
def test_restriction_restatement_is_bit_identical_to_reference_code(coords):
    mesh, r, fine, coarse = _setup(coords, seed=1)
    box = _interior_box(r, mesh)
    a, b = coarse.copy(), coarse.copy()
    oracle_py.restrict_average(oracle_py.lib(), r, fine, a, box)
    oracle_py.restrict_average(ref_py.lib(), r, fine, b, box, prefix="ar")
    assert np.array_equal(a, b)
    assert not np.array_equal(a, coarse)


@needs_ref
@pytest.mark.parametrize("el", [1, 2, 3])
@pytest.mark.parametrize("coords", ALL)
def test_face_restriction_restatement_is_bit_identical_to_reference_code(coords, el):
    """flux correction restricts Metadata::Flux fields: RestrictAverage<GEOM>::Do<DIM, F1|F2|F3>"""
    mesh, r, fine, coarse = _setup(coords, seed=4 + el)
    if el > mesh.ndim:
        pytest.skip("no faces in a collapsed direction")
    box = _interior_box(r, mesh)
    box[2 * (el - 1) + 1] += 1          # a face field has one more element along its direction
    a, b = coarse.copy(), coarse.copy()
    oracle_py.restrict_average_face(oracle_py.lib(), r, fine, a, box, el)
    oracle_py.restrict_average_face(ref_py.lib(), r, fine, b, box, el, prefix="ar")
    assert np.array_equal(a, b)
    assert not np.array_equal(a, coarse)


@needs_ref
@pytest.mark.parametrize("coords", ALL)
def test_prolongation_restatement_is_bit_identical_to_reference_code(coords):
    mesh, r, fine, coarse = _setup(coords, seed=2)
    # jumps and extrema so both minmod branches and the sign logic are exercised
    coarse[:, :, :, ::3] *= -1.0
    box = _interior_box(r, mesh, grow=1)   # prolongation also fills the first fine ghost pairs
    a, b = fine.copy(), fine.copy()
    oracle_py.prolongate_minmod(oracle_py.lib(), r, coarse, a, box)
    oracle_py.prolongate_minmod(ref_py.lib(), r, coarse, b, box, prefix="ar")
    assert np.array_equal(a, b)
    assert not np.array_equal(a, fine)


@pytest.mark.parametrize("coords", ALL)
def test_restriction_conserves_the_volume_integral(coords):
    mesh, r, fine, coarse = _setup(coords, seed=3, nvar=1)
    box = _interior_box(r, mesh)
    oracle_py.restrict_average(oracle_py.lib(), r, fine, coarse, box)
    # volumes from the oracle's geometry (ao_geom_cell is pinned by the flux tests)
    L = oracle_py.lib()
    import ctypes as C
    out = (C.c_double * 32)()

    def vol(xmin, dx, k, j, i):
        L.ao_geom_cell(int(mesh.coords), (C.c_double * 3)(*xmin), (C.c_double * 3)(*dx), k, j, i,
                       out)
        return out[6]

    xmin = [r.xmin[d] for d in range(3)]
    dx = [r.dx[d] for d in range(3)]
    tot_f = tot_c = 0.0
    nd = mesh.ndim
    for ck in range(box[4], box[5] + 1):
        for cj in range(box[2], box[3] + 1):
            for ci in range(box[0], box[1] + 1):
                i = (ci - r.cib_s) * 2 + r.ib_s
                j = (cj - r.cjb_s) * 2 + r.jb_s if nd > 1 else r.jb_s
                k = (ck - r.ckb_s) * 2 + r.kb_s if nd > 2 else r.kb_s
                vs = 0.0
                for ok in range(2 if nd > 2 else 1):
                    for oj in range(2 if nd > 1 else 1):
                        for oi in range(2):
                            v = vol(xmin, dx, k + ok, j + oj, i + oi)
                            vs += v
                            tot_f += v * fine[0, k + ok, j + oj, i + oi]
                tot_c += vs * coarse[0, ck, cj, ci]
    assert abs(tot_f - tot_c) <= 1e-13 * abs(tot_f)


@pytest.mark.parametrize("coords", ALL)
def test_prolongation_is_exact_on_constants_and_inverted_by_restriction(coords):
    mesh, r, fine, coarse = _setup(coords, seed=4)
    box = _interior_box(r, mesh)
    const = np.full_like(coarse, 2.5)
    f = fine.copy()
    oracle_py.prolongate_minmod(oracle_py.lib(), r, const, f, box)
    sl = mesh.interior()
    assert np.array_equal(f[(slice(None),) + sl], np.full_like(f[(slice(None),) + sl], 2.5))
    # restriction of the prolonged field returns the coarse field in Cartesian geometry
    # (the limited slope is antisymmetric about the coarse centre); curvilinear: to rounding
    f = fine.copy()
    oracle_py.prolongate_minmod(oracle_py.lib(), r, coarse, f, box)
    back = coarse.copy()
    oracle_py.restrict_average(oracle_py.lib(), r, f, back, box)
    csl = tuple(slice(box[2 * d], box[2 * d + 1] + 1) for d in (2, 1, 0))
    tol = 1e-14 if coords == Coordinates.cartesian else 0.2
    err = np.abs(back[(slice(None),) + csl] - coarse[(slice(None),) + csl]).max()
    assert err <= tol * np.abs(coarse).max()


@needs_ref
@pytest.mark.parametrize("ndim", [1, 2])
def test_lower_dimensional_cartesian_meshes_match_reference_code(ndim):
    """DIM = 1 and 2 instantiations of both stencils (unused directions contribute exact zeros)."""
    mesh = make_mesh(Coordinates.cartesian, ndim, nblk=(2, 1, 1), bnx=(8, 6, 4),
                     bcs=(BoundaryFlag.outflow,) * 6)
    r = oracle_py.refine_geom(mesh, b=0)
    rng = np.random.default_rng(9)
    fine = 1.0 + rng.random((2, mesh.nk, mesh.nj, mesh.ni))
    coarse = rng.normal(size=(2, r.cnk, r.cnj, r.cni))
    box = _interior_box(r, mesh)
    a, b = coarse.copy(), coarse.copy()
    oracle_py.restrict_average(oracle_py.lib(), r, fine, a, box)
    oracle_py.restrict_average(ref_py.lib(), r, fine, b, box, prefix="ar")
    assert np.array_equal(a, b) and not np.array_equal(a, coarse)
    box = _interior_box(r, mesh, grow=1)
    a, b = fine.copy(), fine.copy()
    oracle_py.prolongate_minmod(oracle_py.lib(), r, coarse, a, box)
    oracle_py.prolongate_minmod(ref_py.lib(), r, coarse, b, box, prefix="ar")
    assert np.array_equal(a, b) and not np.array_equal(a, fine)
