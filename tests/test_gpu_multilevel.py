"""GPU execution of the multilevel ghost exchange: the plan of
artemis_b200.multilevel (CalcIndices pinned to the reference, tests/test_multilevel_plan.py)
run through ab200_restrict / ab200_box_copy / ab200_block_bcs / ab200_prolongate must equal the
CPU executor (oracle restriction / prolongation pinned to the reference's operators) BIT FOR BIT
in the strict build -- random gas + dust states with jumps, Cartesian and curvilinear meshes,
periodic / outflow / reflecting boundaries -- and to rounding in the default build."""
import numpy as np
import pytest

from artemis_b200.enums import BoundaryFlag, Coordinates
from artemis_b200.meshdata import MeshData
from artemis_b200.multilevel import (MultilevelExchange, MultilevelMesh, exchange_plan,
                                     fill_ghost_ranges)
from oracle import multilevel_py
from tests.helpers import GEOM_DOMAINS, dust_params, gas_params, random_prim

pytestmark = pytest.mark.gpu

B = BoundaryFlag
CASES = [
    (Coordinates.cartesian, 3, (B.periodic,) * 6, [(1, 1, 1), (2, 1, 1), (1, 2, 2)]),
    (Coordinates.cartesian, 3, (B.reflect, B.outflow, B.outflow, B.reflect, B.periodic, B.periodic),
     [(0, 0, 0), (3, 3, 1), (1, 2, 0)]),
    (Coordinates.cartesian, 2, (B.outflow, B.reflect, B.reflect, B.outflow, B.periodic, B.periodic),
     [(0, 1, 0), (3, 3, 0), (2, 2, 0)]),
    (Coordinates.cartesian, 1, (B.reflect, B.outflow) + (B.periodic,) * 4, [(1, 0, 0), (3, 0, 0)]),
    (Coordinates.cylindrical, 3, (B.outflow, B.outflow, B.periodic, B.periodic, B.reflect, B.outflow),
     [(1, 1, 1), (0, 2, 0)]),
    (Coordinates.spherical3D, 3, (B.reflect, B.outflow, B.outflow, B.outflow, B.periodic, B.periodic),
     [(1, 1, 1), (3, 0, 2)]),
    (Coordinates.axisymmetric, 2, (B.outflow,) * 4 + (B.periodic,) * 2, [(1, 1, 0), (2, 2, 0)]),
    (Coordinates.spherical2D, 2, (B.outflow,) * 4 + (B.periodic,) * 2, [(0, 0, 0), (2, 1, 0)]),
]


def _mesh(coords, ndim, bcs, refine):
    root = tuple(4 if d < ndim else 1 for d in range(3))
    bnx = tuple(8 if d < ndim else 1 for d in range(3))
    xmin, xmax = GEOM_DOMAINS[coords]
    return MultilevelMesh(root_blocks=root, block_nx=bnx, xmin=xmin, xmax=xmax,
                          refine=tuple(tuple(r) for r in refine), nghost=4, bcs=bcs, coords=coords)


def _vdir(fp):
    S = fp.nspecies
    out = []
    for var0, nc in fill_ghost_ranges(fp):
        out += [((v - S) % 3 + 1) if S <= v < 4 * S else 0 for v in range(var0, var0 + nc)]
    return out


def _vars(fp):
    return [v for var0, nc in fill_ghost_ranges(fp) for v in range(var0, var0 + nc)]


@pytest.mark.parametrize("coords,ndim,bcs,refine", CASES)
@pytest.mark.parametrize("variant", ["strict", "fast"])
def test_multilevel_exchange_equals_cpu_executor(coords, ndim, bcs, refine, variant):
    m = _mesh(coords, ndim, bcs, refine)
    plan = exchange_plan(m)
    gp, dp = gas_params(coords, "plm", "hlle", S=2), dust_params(coords, "plm", "hlle", S=2)
    md = MeshData(m, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
    ex = MultilevelExchange(md, plan)
    kinds = {B.periodic: "periodic", B.outflow: "outflow", B.reflect: "reflect"}
    for ff, seed in zip(md.fluids, (11, 12)):
        prim = random_prim(m, ff.fp, seed=seed)
        # poison every ghost zone: whatever the exchange does not fill shows up
        mask = np.ones(prim.shape[2:], dtype=bool)
        mask[m.interior()] = False
        prim[:, :, mask] = -777.0
        ff.prim.set(prim)
        ff._host = prim
        # the same poison in the coarse buffers on both sides: the comparison then covers every
        # coarse zone, including what the boundary conditions copy out of never-restricted ones
        ex.coarse[int(ff.fp.fluid_type)].set(np.full(m.coarse_shape(ff.fp.nvar), -555.0))
    ex.exchange()
    for ff in md.fluids:
        fine = ff._host.copy()
        coarse = np.full(m.coarse_shape(ff.fp.nvar), -555.0)
        multilevel_py.run_plan(m, plan, fine, coarse, _vars(ff.fp), _vdir(ff.fp),
                               [kinds[b] for b in bcs])
        got = ff.prim.get()
        vs = _vars(ff.fp)
        assert not (got[:, vs] == -777.0).any(), "a ghost zone was never filled"
        if variant == "strict":
            assert np.array_equal(got, fine)
        else:
            assert np.max(np.abs(got - fine)) <= 1e-13 * np.max(np.abs(fine))
        cgot = ex.coarse[int(ff.fp.fluid_type)].get()
        if variant == "strict":
            assert np.array_equal(cgot, coarse)
    assert md.launch_count() > 0
    ex.close()
    md.close()


CYCLE_CASES = [
    (Coordinates.cartesian, 3, (B.periodic,) * 6, [(1, 1, 1), (2, 2, 1)]),
    (Coordinates.cartesian, 2, (B.reflect, B.outflow, B.outflow, B.reflect, B.periodic, B.periodic),
     [(0, 1, 0), (2, 2, 0)]),
    (Coordinates.spherical3D, 3, (B.reflect, B.outflow, B.outflow, B.outflow, B.periodic, B.periodic),
     [(1, 1, 1)]),
]


def _cycle_pair(coords, ndim, bcs, refine, mode, variant, ncyc, integ="rk2", flux_correction=True):
    from artemis_b200.driver import ArtemisDriver
    from oracle.oracle_py import OracleSim
    m = _mesh(coords, ndim, bcs, refine)
    gp, dp = gas_params(coords, "ppm", "hllc"), dust_params(coords, "plm", "hlle", S=1)
    prim, dprim = random_prim(m, gp, seed=41), random_prim(m, dp, seed=42)
    osim = OracleSim(m, gas=gp, dust=dp, integrator=integ)
    osim.flux_correction = flux_correction
    osim.gas.prim[:] = prim
    osim.dust.prim[:] = dprim
    osim.nlim = ncyc
    osim.initialize()
    osim.run()
    md = MeshData(m, gas=gp, dust=dp, variant=variant, materialize_fluxes=(mode == "tasks"))
    md.gas.prim.set(prim)
    md.dust.prim.set(dprim)
    ex = MultilevelExchange(md)
    drv = ArtemisDriver(md, integ, mode=mode, nlim=ncyc, comm=ex, flux_correction=flux_correction)
    drv.Initialize()
    drv.Execute()
    return m, osim, md, drv, ex


@pytest.mark.parametrize("coords,ndim,bcs,refine", CYCLE_CASES)
def test_multilevel_task_cycles_strict_bit_identical(coords, ndim, bcs, refine):
    """whole rk2 cycles on a refined mesh as the reference runs them: fine and coarse blocks
    advance with the same global dt, ghost zones through the multilevel exchange, fluxes
    corrected on every fine-coarse face between CalculateFluxes and ApplyUpdate"""
    m, osim, md, drv, ex = _cycle_pair(coords, ndim, bcs, refine, "tasks", "strict", 2)
    assert ex.n["fc"] > 0
    assert drv.ncycle == osim.ncycle == 2 and drv.dt == osim.dt and drv.time == osim.time
    for ff, of in zip(md.fluids, osim.fluids):
        assert np.array_equal(ff.u0.get(), of.u0)
        assert np.array_equal(ff.prim.get(), of.prim)
    ex.close()
    md.close()


@pytest.mark.parametrize("coords,ndim,bcs,refine", CYCLE_CASES[:2])
def test_multilevel_fused_cycle_within_1e12(coords, ndim, bcs, refine):
    from tests.helpers import zone_rel_err
    # the fused path stores no fluxes: on a refined mesh it exists only as the explicitly
    # non-conservative variant, compared with the oracle run the same way
    m, osim, md, drv, ex = _cycle_pair(coords, ndim, bcs, refine, "fused", "fast", 1,
                                       flux_correction=False)
    for ff, of in zip(md.fluids, osim.fluids):
        assert zone_rel_err(ff.u0.get(), of.u0, of.fp, "cons") <= 1e-12
        assert zone_rel_err(ff.prim.get(), of.prim, of.fp, "prim") <= 1e-12
    ex.close()
    md.close()


def test_fused_path_refuses_a_refined_mesh_with_flux_correction():
    from artemis_b200.driver import ArtemisDriver
    m = _mesh(Coordinates.cartesian, 2, (B.outflow,) * 4 + (B.periodic,) * 2, [(1, 1, 0)])
    md = MeshData(m, gas=gas_params(Coordinates.cartesian, "plm", "hlle"), materialize_fluxes=False)
    ex = MultilevelExchange(md)
    with pytest.raises(ValueError, match="flux correction"):
        ArtemisDriver(md, "rk2", mode="fused", comm=ex)
    ex.close()
    md.close()


FC_CASES = [CASES[0], CASES[1], CASES[2], CASES[3], CASES[4], CASES[5], CASES[6]]


@pytest.mark.parametrize("coords,ndim,bcs,refine", FC_CASES)
@pytest.mark.parametrize("variant", ["strict", "fast"])
def test_flux_correct_equals_cpu_restriction(coords, ndim, bcs, refine, variant):
    """ab200_flux_correct on random flux fields == SetFluxCorrections on the CPU (the oracle's
    face restriction is pinned to RestrictAverage<GEOM>::Do<DIM, F1|F2|F3> of the reference):
    bit for bit in the strict build; faces that no finer block touches stay untouched"""
    from artemis_b200.multilevel import flux_correction_plan
    m = _mesh(coords, ndim, bcs, refine)
    gp, dp = gas_params(coords, "plm", "hlle", S=2), dust_params(coords, "plm", "hlle", S=2)
    md = MeshData(m, gas=gp, dust=dp, variant=variant, materialize_fluxes=True)
    ex = MultilevelExchange(md)
    plan = flux_correction_plan(m)
    assert ex.n["fc"] == 2 * len(plan) > 0
    rng = np.random.default_rng(77)
    host = []
    for ff in md.fluids:
        fl = [rng.standard_normal(m.shape(ff.fp.nvar)) if d < ndim else None for d in range(3)]
        pf = [rng.standard_normal(m.shape(ff.fp.nspecies))
              if (d < ndim and ff.pflux[d] is not None) else None for d in range(3)]
        for d in range(ndim):
            ff.flux[d].set(fl[d])
            if pf[d] is not None:
                ff.pflux[d].set(pf[d])
        host.append((fl, pf))
    ex.flux_correct()
    for ff, (fl, pf) in zip(md.fluids, host):
        want = [a.copy() if a is not None else None for a in fl]
        multilevel_py.flux_correct(m, plan, want)
        wantp = [a.copy() if a is not None else None for a in pf]
        if any(a is not None for a in pf):
            multilevel_py.flux_correct(m, plan, wantp)
        for d in range(ndim):
            got = ff.flux[d].get()
            assert not np.array_equal(got, fl[d])
            if variant == "strict":
                assert np.array_equal(got, want[d])
            else:
                assert np.max(np.abs(got - want[d])) <= 1e-13 * np.max(np.abs(want[d]))
            if pf[d] is not None:
                gotp = ff.pflux[d].get()
                if variant == "strict":
                    assert np.array_equal(gotp, wantp[d])
                else:
                    assert np.max(np.abs(gotp - wantp[d])) <= 1e-13 * np.max(np.abs(wantp[d]))
    ex.close()
    md.close()


def test_multilevel_task_cycles_conserve_mass_on_the_gpu():
    """the GPU's own conservation check (periodic refined mesh, 3 rk2 cycles, task path)"""
    m, osim, md, drv, ex = _cycle_pair(Coordinates.cartesian, 3, (B.periodic,) * 6,
                                       [(1, 1, 1), (2, 2, 1)], "tasks", "fast", 3)
    vol = np.prod(m.blk_dx, axis=1)
    for ff, of in zip(md.fluids, osim.fluids):
        u = ff.u0.get()
        for v in range(min(5, ff.fp.nvar)):
            tot = lambda a: sum(vol[b] * a[(b, v) + m.interior()].sum() for b in range(m.nb))
            assert abs(tot(u) - tot(of.u0)) <= 1e-13 * tot(np.abs(of.u0))
    ex.close()
    md.close()


def test_box_copy_and_block_bcs_reject_bad_descriptors():
    import ctypes as C
    from artemis_b200 import capi
    m = _mesh(Coordinates.cartesian, 2, (B.outflow,) * 4 + (B.periodic,) * 2, [(1, 1, 0)])
    gp = gas_params(Coordinates.cartesian, "plm", "hlle")
    md = MeshData(m, gas=gp, materialize_fluxes=False)
    bad = (capi.BoxDesc * 1)(capi.BoxDesc(0, 4, 0, 0, None, 1, 0, None, 0, 0, 0, 0, 0, 0, m.ni + 1, 1, 1))
    with pytest.raises(capi.AB200Error, match="outside the array"):
        md.call("ab200_box_copy", bad, 1)
    badbc = (capi.BlockBcDesc * 1)(capi.BlockBcDesc(0, 0, 0, 4, 5, 1, None))   # x3 face in 2-D
    with pytest.raises(capi.AB200Error, match="bad face"):
        md.call("ab200_block_bcs", badbc, 1)
    md.close()


@pytest.mark.parametrize("ndim,refine", [(3, [(1, 1, 1), (2, 2, 1)]), (2, [(0, 1, 0), (2, 2, 0)])])
def test_flux_correction_of_the_diffusion_fluxes_strict_bit_identical(ndim, refine):
    """gas.diff.momentum / gas.diff.energy are Metadata::WithFluxes fields
    (src/gas/gas.cpp:277-285): ab200_flux_correct restricts them onto every fine-coarse face
    together with the hydrodynamic fluxes.  rk2 cycles with constant viscosity + conduction on a
    periodic refined mesh: bit-identical to the oracle (whose corrected run conserves momentum
    and energy to rounding, tests/test_multilevel_plan.py), and the corrected face values
    themselves are compared"""
    from artemis_b200.driver import ArtemisDriver
    from oracle.oracle_py import OracleSim, make_diffusion
    from tests.test_gpu_diffusion import _flux, desc_of
    bcs = (B.periodic,) * 6
    m = _mesh(Coordinates.cartesian, ndim, bcs, refine)
    gp = gas_params(Coordinates.cartesian, "plm", "hlle")
    prim = random_prim(m, gp, seed=61)
    osim = OracleSim(m, gas=gp)
    osim.diffusion = make_diffusion(visc=("constant", 2e-2), cond=("conductivity", 3e-2))
    osim.gas.prim[:] = prim
    osim.nlim = 2
    osim.initialize()
    osim.run()
    md = MeshData(m, gas=gp, variant="strict", materialize_fluxes=True)
    md.gas.prim.set(prim)
    ex = MultilevelExchange(md)
    drv = ArtemisDriver(md, "rk2", mode="tasks", nlim=2, comm=ex,
                        diffusion=desc_of(osim.diffusion))
    drv.Initialize()
    drv.Execute()
    assert ex.n["fc"] > 0 and drv.ncycle == osim.ncycle == 2
    assert drv.dt == osim.dt and drv.time == osim.time
    assert np.array_equal(md.gas.u0.get(), osim.gas.u0)
    assert np.array_equal(md.gas.prim.get(), osim.gas.prim)
    # the diffusion flux arrays of the last stage, corrected faces included
    for d in range(ndim):
        assert np.array_equal(_flux(md, d, osim.dflx[d].shape), osim.dflx[d])
    ex.close()
    md.close()
