"""GPU parity of the multilevel operators through the C ABI:
ab200_restrict == ArtemisUtils::RestrictAverage<GEOM>, ab200_prolongate ==
ArtemisUtils::ProlongateSharedMinMod<GEOM>.  The oracle (pinned bit for bit to the reference's
own headers in tests/test_refine_oracle.py) is the checker: strict build bit-identical, fast
build within 1e-14 (restriction is a ratio of sums, prolongation has two divisions)."""
import ctypes as C

import numpy as np
import pytest

from artemis_b200 import capi
from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid
from artemis_b200.meshdata import MeshData
from oracle import oracle_py
from tests.helpers import dust_params, gas_params, make_mesh, rel_err

pytestmark = pytest.mark.gpu
ALL = [Coordinates.cartesian, Coordinates.cylindrical, Coordinates.axisymmetric,
       Coordinates.spherical1D, Coordinates.spherical2D, Coordinates.spherical3D]


class _Coarse:
    """Device coarse buffers [nb][nvar][cnk][cnj][cni] of one fluid + their descriptors."""

    def __init__(self, md, ff, fluid, kind, box, var0=0, nvar=None):
        dims = (C.c_int * 6)()
        md.call("ab200_coarse_shape", dims)
        self.cn = (dims[2], dims[1], dims[0])
        self.cs = (dims[5], dims[4], dims[3])
        nb = md.mesh.nb
        self.nvar = nvar or ff.fp.nvar
        self.shape = (nb, self.nvar) + self.cn
        self.nbytes = int(np.prod(self.shape)) * 8
        p = C.c_void_p()
        capi.check(md.L, md.L.ab200_malloc(md.ctx, C.byref(p), self.nbytes), "ab200_malloc")
        self.ptr, self.md = p.value, md
        per = int(np.prod(self.shape[1:])) * 8
        self.descs = (capi.RefineDesc * nb)()
        for b in range(nb):
            self.descs[b] = capi.RefineDesc(int(fluid), b, var0, self.nvar, kind, box[0], box[1],
                                            box[2], box[3], box[4], box[5], self.ptr + b * per)

    def set(self, a):
        a = np.ascontiguousarray(a)
        capi.check(self.md.L, self.md.L.ab200_memcpy_h2d(self.md.ctx, self.ptr, a.ctypes.data,
                                                          self.nbytes), "h2d")

    def get(self):
        out = np.empty(self.shape)
        capi.check(self.md.L, self.md.L.ab200_memcpy_d2h(self.md.ctx, out.ctypes.data, self.ptr,
                                                          self.nbytes), "d2h")
        return out

    def free(self):
        self.md.L.ab200_free(self.md.ctx, self.ptr)


def _box(mesh, r, grow):
    box = []
    for d, cs in enumerate((r.cib_s, r.cjb_s, r.ckb_s)):
        if d < mesh.ndim:
            box += [cs - grow, cs + mesh.block_nx[d] // 2 - 1 + grow]
        else:
            box += [0, 0]
    return box


@pytest.mark.parametrize("variant", ["strict", "fast"])
@pytest.mark.parametrize("coords", ALL)
def test_restrict_and_prolongate_match_the_oracle(coords, variant):
    mesh = make_mesh(coords, 3, nblk=(2, 1, 1), bnx=(8, 6, 4), bcs=(BoundaryFlag.outflow,) * 6)
    gp = gas_params(coords, "plm", "hlle", S=2)
    dp = dust_params(coords, "plm", "hlle", S=1)
    md = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
    rng = np.random.default_rng(11)
    L = oracle_py.lib()
    for fluid, ff, kind in ((Fluid.gas, md.gas, 0), (Fluid.dust, md.dust, 1)):
        nv = ff.fp.nvar
        fine = 1.0 + rng.random(mesh.shape(nv))
        arr = ff.prim if kind == 0 else ff.u0
        arr.set(fine)
        r0 = oracle_py.refine_geom(mesh, 0)
        # ---- restriction over the coarse interior --------------------------------------------
        box = _box(mesh, r0, 0)
        cb = _Coarse(md, ff, fluid, kind, box)
        c0 = 1.0 + rng.random(cb.shape)
        cb.set(c0)
        md.call("ab200_restrict", cb.descs, mesh.nb)
        got = cb.get()
        want = c0.copy()
        for b in range(mesh.nb):
            oracle_py.restrict_average(L, oracle_py.refine_geom(mesh, b), fine[b], want[b], box)
        if variant == "strict":
            assert np.array_equal(got, want), (coords, fluid)
        else:
            assert rel_err(got, want) <= 1e-14, (coords, fluid)
        cb.free()
        # ---- prolongation over the interior + one coarse ghost layer ---------------------------
        box = _box(mesh, r0, 1)
        cb = _Coarse(md, ff, fluid, kind, box)
        c0 = 1.0 + rng.random(cb.shape)
        c0[..., ::3] *= -1.0   # extrema and sign changes: both minmod branches
        cb.set(c0)
        md.call("ab200_prolongate", cb.descs, mesh.nb)
        got = arr.get()
        want = fine.copy()
        for b in range(mesh.nb):
            oracle_py.prolongate_minmod(L, oracle_py.refine_geom(mesh, b), c0[b], want[b], box)
        if variant == "strict":
            assert np.array_equal(got, want), (coords, fluid)
        else:
            assert rel_err(got, want) <= 1e-13, (coords, fluid)
        assert not np.array_equal(got, fine)
        cb.free()
    md.close()


def test_partial_variable_ranges_and_one_launch_per_list():
    """A descriptor may name a sub-range of the pack (Parthenon restricts variable by variable)
    and a whole list is ONE launch."""
    coords = Coordinates.cartesian
    mesh = make_mesh(coords, 3, nblk=(2, 2, 1), bnx=(8, 6, 4), bcs=(BoundaryFlag.periodic,) * 6)
    gp = gas_params(coords, "ppm", "hllc", S=2)
    md = MeshData(mesh, gas=gp, variant="strict", materialize_fluxes=False)
    rng = np.random.default_rng(5)
    fine = 1.0 + rng.random(mesh.shape(gp.nvar))
    md.gas.prim.set(fine)
    r0 = oracle_py.refine_geom(mesh, 0)
    box = _box(mesh, r0, 0)
    var0, nvar = 2, 6   # the velocity block of 2 species
    cb = _Coarse(md, md.gas, Fluid.gas, 0, box, var0=var0, nvar=nvar)
    c0 = np.zeros(cb.shape)
    cb.set(c0)
    n0 = md.launch_count()
    md.call("ab200_restrict", cb.descs, mesh.nb)
    assert md.launch_count() - n0 == 1
    got = cb.get()
    want = c0.copy()
    for b in range(mesh.nb):
        oracle_py.restrict_average(oracle_py.lib(), oracle_py.refine_geom(mesh, b),
                                   np.ascontiguousarray(fine[b, var0:var0 + nvar]), want[b], box)
    assert np.array_equal(got, want)
    cb.free()
    md.close()


def test_refine_rejects_bad_descriptors():
    coords = Coordinates.cartesian
    mesh = make_mesh(coords, 3, nblk=(1, 1, 1), bnx=(8, 6, 4), bcs=(BoundaryFlag.periodic,) * 6)
    gp = gas_params(coords, "ppm", "hllc")
    md = MeshData(mesh, gas=gp, materialize_fluxes=False)
    r0 = oracle_py.refine_geom(mesh, 0)
    box = _box(mesh, r0, 0)
    cb = _Coarse(md, md.gas, Fluid.gas, 0, box)
    bad = (capi.RefineDesc * 1)()
    for field, value, msg in (("cis", 0, "outside the coarse buffer"),   # stencil reads index -1
                              ("block", 7, "out of range"), ("nvar", 99, "out of range"),
                              ("kind", 5, "unknown array kind"), ("coarse", None, "null coarse")):
        bad[0] = cb.descs[0]
        setattr(bad[0], field, value)
        with pytest.raises(capi.AB200Error, match=msg):
            md.call("ab200_prolongate", bad, 1)
    dust = (capi.RefineDesc * 1)()
    dust[0] = cb.descs[0]
    dust[0].fluid = 1
    with pytest.raises(capi.AB200Error, match="unbound fluid"):
        md.call("ab200_restrict", dust, 1)
    cb.free()
    md.close()


@pytest.mark.parametrize("ndim", [1, 2])
def test_lower_dimensional_cartesian_meshes(ndim):
    """DIM = 1 and 2: the unused directions contribute exact zeros, like in the reference."""
    coords = Coordinates.cartesian
    mesh = make_mesh(coords, ndim, nblk=(2, 1, 1), bnx=(8, 6, 4), bcs=(BoundaryFlag.outflow,) * 6)
    gp = gas_params(coords, "plm", "hlle")
    md = MeshData(mesh, gas=gp, variant="strict", materialize_fluxes=False)
    rng = np.random.default_rng(3)
    fine = 1.0 + rng.random(mesh.shape(gp.nvar))
    md.gas.prim.set(fine)
    r0 = oracle_py.refine_geom(mesh, 0)
    L = oracle_py.lib()
    box = _box(mesh, r0, 0)
    cb = _Coarse(md, md.gas, Fluid.gas, 0, box)
    c0 = rng.normal(size=cb.shape)
    cb.set(c0)
    md.call("ab200_restrict", cb.descs, mesh.nb)
    want = c0.copy()
    for b in range(mesh.nb):
        oracle_py.restrict_average(L, oracle_py.refine_geom(mesh, b), fine[b], want[b], box)
    assert np.array_equal(cb.get(), want)
    cb.free()
    box = _box(mesh, r0, 1)
    cb = _Coarse(md, md.gas, Fluid.gas, 0, box)
    cb.set(c0)
    md.call("ab200_prolongate", cb.descs, mesh.nb)
    want = fine.copy()
    for b in range(mesh.nb):
        oracle_py.prolongate_minmod(L, oracle_py.refine_geom(mesh, b), c0[b], want[b], box)
    assert np.array_equal(md.gas.prim.get(), want)
    cb.free()
    md.close()
