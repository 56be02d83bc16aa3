"""Pins the oracle (oracle/artemis_oracle.c, the CPU restatement every GPU parity test checks
against) to the reference's OWN code: oracle/_ref/libartemis_ref.so is built from the
unmodified sources under /root/reference/src (fluid_fluxes.hpp, reconstruction/*, riemann/*,
geometry/*, artemis_integrator.hpp, derived/fill_derived.cpp) against a mock Parthenon.
Same seeded inputs through both; results must be BIT-IDENTICAL (both are gcc, no FMA)."""
import numpy as np
import pytest

from artemis_b200 import pgen
from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
from artemis_b200.mesh import UniformMesh
from artemis_b200.params import FluidParams
from oracle import ref_py
from oracle.oracle_py import OracleSim
from tests.helpers import dust_params, gas_params, make_mesh, random_prim

pytestmark = pytest.mark.skipif(not ref_py.available(),
                                reason="oracle/_ref not built and /root/reference not mounted")

GEOMS = [Coordinates.cartesian, Coordinates.cylindrical, Coordinates.axisymmetric,
         Coordinates.spherical3D, Coordinates.spherical2D, Coordinates.spherical1D]


def _pair(mesh, gp, dp, integ="rk2", omf=0.0, seed=5):
    sims = []
    for cls in (OracleSim, ref_py.RefSim):
        s = cls(mesh, gas=gp, dust=dp, integrator=integ, omf=omf)
        if gp is not None:
            s.gas.prim[:] = random_prim(mesh, gp, seed=seed)
        if dp is not None:
            s.dust.prim[:] = random_prim(mesh, dp, seed=seed + 1)
        sims.append(s)
    return sims


@pytest.mark.parametrize("coords", GEOMS)
@pytest.mark.parametrize("recon", ["pcm", "plm", "ppm"])
@pytest.mark.parametrize("rs", ["hllc", "hlle", "llf"])
def test_every_task_bit_identical_to_reference_code(coords, recon, rs):
    """One stage, task by task, on seeded states with shocks: fluxes, interface pressure, face
    velocity, update, flux source (with a rotating frame), aux, C2P, P2C, dt."""
    bcs = (BoundaryFlag.outflow,) * 6
    mesh = make_mesh(coords, 3, bcs=bcs)
    gp = gas_params(coords, recon, rs, S=2, de_switch=0.02)
    dp = dust_params(coords, recon, "hlle" if rs == "hllc" else rs, S=2)
    o, r = _pair(mesh, gp, dp, omf=0.3)
    for s in (o, r):
        for fs in s.fluids:
            s.PrimToCons(fs)
            np.copyto(fs.u1, fs.u0)
    assert o.EstimateTimestep() == r.EstimateTimestep()
    dt = 0.37 * o.EstimateTimestep()
    for s in (o, r):
        for fs in s.fluids:
            s.CalculateFluxes(fs, False)
    for fo, fr in zip(o.fluids, r.fluids):
        for d in range(mesh.ndim):
            assert np.array_equal(fo.flux[d], fr.flux[d]), (d, "flux")
            if fo.pflux[d] is not None:
                assert np.array_equal(fo.pflux[d], fr.pflux[d]), (d, "pflux")
                assert np.array_equal(fo.vface[d], fr.vface[d]), (d, "vface")
    for s in (o, r):
        for fs in s.fluids:
            s.ApplyUpdate(fs, 0.25, 0.75, 0.25 * dt)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0), "ApplyUpdate"
    for s in (o, r):
        for fs in s.fluids:
            s.FluxSource(fs, 0.25 * dt)
    sl = mesh.interior()
    for fo, fr in zip(o.fluids, r.fluids):   # the reference also touches 3 ghost columns in i
        assert np.array_equal(fo.u0[(slice(None), slice(None)) + sl],
                              fr.u0[(slice(None), slice(None)) + sl]), "FluxSource"
        fr.u0[:] = fo.u0
    for s in (o, r):
        for fs in s.fluids:
            s.SetAuxillaryFields(fs)
            s.ConsToPrim(fs)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0), "SetAuxillaryFields"
        assert np.array_equal(fo.prim, fr.prim), "ConsToPrim"
    for s in (o, r):
        for fs in s.fluids:
            s.PrimToCons(fs)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim), "PrimToCons"


def test_pcm_flag_of_vl2_stage_one():
    mesh = make_mesh(Coordinates.cartesian, 3, bcs=(BoundaryFlag.outflow,) * 6)
    gp = gas_params(Coordinates.cartesian, "ppm", "hllc")
    o, r = _pair(mesh, gp, None)
    for s in (o, r):
        s.CalculateFluxes(s.gas, True)
    for d in range(3):
        assert np.array_equal(o.gas.flux[d], r.gas.flux[d])


@pytest.mark.parametrize("coords,integ,ncyc", [(Coordinates.cartesian, "rk2", 12),
                                               (Coordinates.spherical3D, "vl2", 6),
                                               (Coordinates.cylindrical, "rk3", 6)])
def test_whole_cycles_bit_identical(coords, integ, ncyc):
    """Full driver loop (reference kernels + the oracle's ghost exchange) vs the oracle."""
    bcs = (BoundaryFlag.outflow,) * 6 if coords != Coordinates.cartesian else None
    mesh = make_mesh(coords, 3, bcs=bcs)
    gp = gas_params(coords, "ppm", "hllc")
    dp = dust_params(coords, "plm", "hlle", S=2)
    o, r = _pair(mesh, gp, dp, integ=integ, seed=9)
    for s in (o, r):
        s.nlim = ncyc
        s.initialize()
        s.run()
    assert o.time == r.time and o.dt == r.dt
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim)


def test_reference_code_reproduces_the_published_linwave_number():
    """The mock-Parthenon build of the reference's kernels reproduces the RMS-L1 error the
    real Kokkos-OpenMP build printed (tests/golden/reference_linwave.json)."""
    res = 16
    mesh = UniformMesh(nx=(res, res // 2, res // 2), xmin=(0, 0, 0), xmax=(3.0, 1.5, 1.5),
                       block_nx=(res // 4,) * 3, nghost=4)
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod.ppm, RSolver.hllc,
                     cfl=0.9, nspecies=1, dfloor=1e-20, gamma=1.66666666667)
    prim, lw = pgen.linear_wave(mesh, gp.gamma, 0, 1e-6, 0.0)
    r = ref_py.RefSim(mesh, gas=gp)
    r.gas.prim[:] = prim
    r.tlim, r.nlim = lw.tlim, 1000
    r.initialize()
    r.run()
    rms, _ = pgen.linear_wave_errors(mesh, lw, r.gas.u0)
    assert r.ncycle == 18 and f"{rms:.6e}" == "3.780974e-07"
