"""GPU parity of the diffusion operators: ab200_diffusion_flux /
ab200_diffusion_update / ab200_diffusion_timestep and the stage drivers that host them, against
the oracle (pinned bit for bit to the reference's own diffusion headers,
tests/test_diffusion_oracle.py).  Strict build: bit for bit wherever the coefficient law does not
call pow() with a non-trivial exponent; <= 1e-13 otherwise (CUDA's pow vs libm's)."""
import ctypes as C

import numpy as np
import pytest

from artemis_b200 import capi
from artemis_b200.driver import ArtemisDriver
from artemis_b200.enums import BoundaryFlag, Coordinates
from artemis_b200.meshdata import MeshData
from oracle.oracle_py import OracleSim, make_diffusion
from tests.helpers import gas_params, make_mesh, random_prim, zone_rel_err

pytestmark = pytest.mark.gpu

GEOMS = [(Coordinates.cartesian, 3), (Coordinates.cartesian, 2), (Coordinates.cartesian, 1),
         (Coordinates.cylindrical, 3), (Coordinates.spherical1D, 1), (Coordinates.spherical2D, 2),
         (Coordinates.spherical3D, 3), (Coordinates.axisymmetric, 2)]
# pow-free laws (exponents 0): the strict build must be bit-identical
EXACT = {
    "constant_viscosity_bulk": dict(visc=("constant", 3e-3, 0.0, 1.7)),
    "viscosity_harmonic_and_diffusivity": dict(visc=("constant", 3e-3, "harmonic"),
                                               cond=("diffusivity", 5e-3), cv=1.1),
    "conductivity_harmonic": dict(cond=("conductivity", 4e-3, "harmonic"), cv=1.3),
}
# laws that call pow(): <= 1e-13
POW = {
    "alpha_viscosity": dict(visc=("alpha", 1e-2, 0.3)),
    "powerlaw_viscosity_and_conductivity": dict(visc=("powerlaw", 2e-3, 0.5),
                                                cond=("conductivity", 4e-3, 0.5, -0.25), cv=1.3),
}


def desc_of(d):
    """oracle ao_diffusion -> ab200_diffusion_desc (same fields, same order)"""
    out = capi.DiffusionDesc()
    for (name, _), (oname, _) in zip(out._fields_, d._fields_):
        setattr(out, name, getattr(d, oname))
    return out


def _setup(coords, ndim, phys, variant, S=2, integ="rk2", mode="tasks", ncyc=0):
    bcs = (BoundaryFlag.outflow,) * 6 if coords != Coordinates.cartesian else None
    mesh = make_mesh(coords, ndim, bcs=bcs)
    gp = gas_params(coords, "plm", "hlle", S=S)
    prim = random_prim(mesh, gp, seed=31)
    osim = OracleSim(mesh, gas=gp, integrator=integ)
    osim.gas.prim[:] = prim
    osim.diffusion = make_diffusion(**phys)
    osim.nlim = ncyc
    osim.initialize()
    md = MeshData(mesh, gas=gp, variant=variant, materialize_fluxes=(mode == "tasks"))
    md.gas.prim.set(prim)
    drv = ArtemisDriver(md, integ, mode=mode, nlim=ncyc, diffusion=desc_of(osim.diffusion))
    drv.Initialize()
    return mesh, osim, md, drv


def _flux(md, d, shape):
    ptr, cnt = C.POINTER(C.c_double)(), C.c_size_t(0)
    md.call("ab200_diffusion_flux_array", d + 1, C.byref(ptr), C.byref(cnt))
    out = np.zeros(shape)
    assert cnt.value == out.size
    md.call("ab200_memcpy_d2h", out.ctypes.data_as(C.c_void_p), C.cast(ptr, C.c_void_p),
            out.nbytes)
    md.synchronize()
    return out


@pytest.mark.parametrize("coords,ndim", GEOMS)
@pytest.mark.parametrize("phys", sorted(EXACT))
def test_strict_fluxes_update_and_timestep_bit_identical(coords, ndim, phys):
    mesh, osim, md, drv = _setup(coords, ndim, EXACT[phys], "strict")
    assert drv.dt == osim.dt                       # hydro + diffusive limits
    osim.DiffusionFlux()
    md.call("ab200_diffusion_flux")
    for d in range(ndim):
        got = _flux(md, d, osim.dflx[d].shape)
        assert np.array_equal(got, osim.dflx[d]), f"x{d + 1} flux"
    osim.DiffusionUpdate(0.37 * osim.dt)
    md.call("ab200_diffusion_update", 0.37 * osim.dt)
    assert np.array_equal(md.gas.u0.get(), osim.gas.u0)
    dt = C.c_double(0.0)
    md.call("ab200_diffusion_timestep", C.byref(dt))
    assert dt.value == osim.DiffusionTimestep()
    md.close()


@pytest.mark.parametrize("coords,ndim", GEOMS)
@pytest.mark.parametrize("phys", sorted(POW))
@pytest.mark.parametrize("variant", ["strict", "fast"])
def test_pow_laws_within_1e13(coords, ndim, phys, variant):
    mesh, osim, md, drv = _setup(coords, ndim, POW[phys], variant)
    osim.DiffusionFlux()
    md.call("ab200_diffusion_flux")
    for d in range(ndim):
        got, want = _flux(md, d, osim.dflx[d].shape), osim.dflx[d]
        assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))
    dt = C.c_double(0.0)
    md.call("ab200_diffusion_timestep", C.byref(dt))
    assert abs(dt.value - osim.DiffusionTimestep()) <= 1e-13 * dt.value
    md.close()


@pytest.mark.parametrize("coords,ndim", [(Coordinates.cartesian, 3), (Coordinates.spherical3D, 3),
                                         (Coordinates.cylindrical, 3), (Coordinates.axisymmetric, 2)])
def test_task_cycles_strict_bit_identical(coords, ndim):
    mesh, osim, md, drv = _setup(coords, ndim, EXACT["viscosity_harmonic_and_diffusivity"],
                                 "strict", S=1, integ="vl2", ncyc=2)
    osim.run()
    drv.Execute()
    assert drv.ncycle == osim.ncycle == 2 and drv.dt == osim.dt and drv.time == osim.time
    assert np.array_equal(md.gas.u0.get(), osim.gas.u0)
    assert np.array_equal(md.gas.prim.get(), osim.gas.prim)
    md.close()


@pytest.mark.parametrize("coords,ndim", [(Coordinates.cartesian, 3), (Coordinates.spherical3D, 3)])
@pytest.mark.parametrize("phys", ["constant_viscosity_bulk", "alpha_viscosity"])
def test_split_fused_stage_with_diffusion_within_1e12(coords, ndim, phys):
    """fast build, fused passes + deferred C2P + diffusion, one cycle"""
    mesh, osim, md, drv = _setup(coords, ndim, {**EXACT, **POW}[phys], "fast", S=1, mode="fused",
                                 ncyc=1)
    osim.run()
    drv.Execute()
    assert zone_rel_err(md.gas.u0.get(), osim.gas.u0, osim.gas.fp, "cons") <= 1e-12
    assert zone_rel_err(md.gas.prim.get(), osim.gas.prim, osim.gas.fp, "prim") <= 1e-12
    assert abs(drv.dt - osim.dt) <= 1e-12 * osim.dt
    md.close()


@pytest.mark.parametrize("variant", ["strict", "fast"])
def test_device_resident_cycles_with_diffusion(variant):
    """ab200_configure_diffusion + ab200_run_cycles (diffusion flux / update every stage, the
    diffusive limits folded into the device timestep) == the host-driven fused loop."""
    coords = Coordinates.spherical3D
    phys = EXACT["viscosity_harmonic_and_diffusivity"]
    ncyc = 3
    mesh, osim, md1, d1 = _setup(coords, 3, phys, variant, S=1, mode="fused", ncyc=ncyc)
    d1.Execute()
    mesh, osim, md2, d2 = _setup(coords, 3, phys, variant, S=1, mode="fused")
    md2.set_time_state(d2.dt)
    md2.call("ab200_run_cycles", 1, ncyc, float(np.finfo(np.float64).max))
    ts = md2.time_state()
    assert int(ts[3]) == ncyc
    if variant == "strict":
        assert np.array_equal(md1.gas.u0.get(), md2.gas.u0.get())
        assert np.array_equal(md1.gas.prim.get(), md2.gas.prim.get())
        assert ts[0] == d1.dt and ts[2] == d1.time
    else:
        assert zone_rel_err(md2.gas.u0.get(), md1.gas.u0.get(), md1.gas.fp, "cons") <= 1e-12
    md1.close()
    md2.close()


def test_diffusion_entry_points_need_configuration():
    mesh = make_mesh(Coordinates.cartesian, 3)
    gp = gas_params(Coordinates.cartesian, "plm", "hlle")
    md = MeshData(mesh, gas=gp, materialize_fluxes=False)
    md.gas.prim.set(random_prim(mesh, gp, seed=3))
    with pytest.raises(capi.AB200Error, match="not configured"):
        md.call("ab200_diffusion_flux")
    dd = capi.DiffusionDesc()
    dd.visc_type, dd.nu, dd.r0 = 1, 1e-3, 1.0
    md.call("ab200_configure_diffusion", C.byref(dd))
    with pytest.raises(capi.AB200Error, match="no diffusion fluxes"):
        md.call("ab200_diffusion_update", 1e-3)
    dd.cond_type = 7
    with pytest.raises(capi.AB200Error, match="Invalid conductivity type"):
        md.call("ab200_configure_diffusion", C.byref(dd))
    md.close()
