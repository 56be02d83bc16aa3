"""Diffusion operators: the oracle's restatement of
Diffusion::{MomentumFluxImpl, ThermalFluxImpl, DiffusionUpdateImpl, EstimateTimestep,
ZeroDiffusionImpl} (src/utils/diffusion/*.hpp, dispatched by src/gas/gas.cpp:437-464, 524-642)
against the reference's OWN headers compiled in oracle/_ref -- bit for bit: face fluxes, the
update, the diffusive timestep, over whole cycles, in every geometry."""
import numpy as np
import pytest

from artemis_b200.enums import BoundaryFlag, Coordinates
from oracle import ref_py
from oracle.oracle_py import OracleSim, make_diffusion
from tests.helpers import gas_params, make_mesh, random_prim

needs_ref = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref not built")

GEOMS = [(Coordinates.cartesian, 3), (Coordinates.cartesian, 2), (Coordinates.cartesian, 1),
         (Coordinates.cylindrical, 3), (Coordinates.spherical1D, 1), (Coordinates.spherical2D, 2),
         (Coordinates.spherical3D, 3), (Coordinates.axisymmetric, 2)]
PHYSICS = {
    "constant_viscosity": dict(visc=("constant", 3e-3)),
    "powerlaw_viscosity_bulk_harmonic": dict(visc=("powerlaw", 2e-3, 0.5, 1.7, "harmonic")),
    "alpha_viscosity": dict(visc=("alpha", 1e-2, 0.3)),
    "conductivity": dict(cond=("conductivity", 4e-3, 0.5, -0.25), cv=1.3),
    "diffusivity_harmonic": dict(cond=("diffusivity", 5e-3, 0.0, 0.0, "harmonic"), cv=0.8),
    "viscosity_and_conduction": dict(visc=("constant", 3e-3), cond=("diffusivity", 5e-3), cv=1.1),
}


def _pair(coords, ndim, phys, integ="rk2", ncyc=1, S=1):
    bcs = (BoundaryFlag.outflow,) * 6 if coords != Coordinates.cartesian else None
    mesh = make_mesh(coords, ndim, bcs=bcs)
    gp = gas_params(coords, "plm", "hlle", S=S)
    sims = []
    for cls in (OracleSim, ref_py.RefSim):
        sim = cls(mesh, gas=gp, integrator=integ)
        sim.gas.prim[:] = random_prim(mesh, gp, seed=31)
        sim.diffusion = make_diffusion(**phys)
        sim.nlim = ncyc
        sim.initialize()
        sims.append(sim)
    return sims


@needs_ref
@pytest.mark.parametrize("coords,ndim", GEOMS)
@pytest.mark.parametrize("phys", sorted(PHYSICS))
def test_diffusion_fluxes_and_timestep_bit_identical_to_reference_code(coords, ndim, phys):
    o, r = _pair(coords, ndim, PHYSICS[phys], S=2)
    o.DiffusionFlux()
    r.DiffusionFlux()
    for d in range(ndim):
        assert np.array_equal(o.dflx[d], r.dflx[d]), f"x{d + 1} diffusion flux"
        assert np.any(o.dflx[d] != 0.0)
    assert o.DiffusionTimestep() == r.DiffusionTimestep()
    assert o.dt == r.dt


@needs_ref
@pytest.mark.parametrize("coords,ndim", GEOMS)
def test_cycles_with_viscosity_and_conduction_bit_identical_to_reference_code(coords, ndim):
    o, r = _pair(coords, ndim, PHYSICS["viscosity_and_conduction"], integ="vl2", ncyc=2)
    o.run()
    r.run()
    assert o.ncycle == r.ncycle == 2 and o.time == r.time
    assert np.array_equal(o.gas.u0, r.gas.u0) and np.array_equal(o.gas.prim, r.gas.prim)
    base = OracleSim(o.mesh, gas=o.gas.fp, integrator="vl2")
    base.gas.prim[:] = random_prim(o.mesh, o.gas.fp, seed=31)
    base.nlim = 2
    base.initialize()
    base.run()
    assert not np.array_equal(base.gas.u0, o.gas.u0)     # the operators really acted


@needs_ref
def test_alpha_viscosity_cycle_bit_identical_to_reference_code():
    """inputs/disk/disk_sph.in physics: alpha viscosity on a spherical mesh (both sides call the
    same libm pow)."""
    o, r = _pair(Coordinates.spherical3D, 3, PHYSICS["alpha_viscosity"], ncyc=2)
    o.run()
    r.run()
    assert np.array_equal(o.gas.u0, r.gas.u0) and np.array_equal(o.gas.prim, r.gas.prim)
