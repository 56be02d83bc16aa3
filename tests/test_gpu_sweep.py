"""GPU parity of the single-pass stage kernel (csrc/sweep.cuh): 3-D Cartesian meshes run
reconstruct -> Riemann -> update -> C2P for all three directions in ONE kernel that ping-pongs
between two primitive sets.  Oracle = the pinned CPU restatement (oracle/)."""
import numpy as np
import pytest

from artemis_b200 import pgen
from artemis_b200.driver import ArtemisDriver
from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
from artemis_b200.mesh import UniformMesh
from artemis_b200.meshdata import MeshData
from artemis_b200.params import FluidParams
from oracle.oracle_py import OracleSim
from tests.helpers import dust_params, gas_params, random_prim, zone_rel_err

pytestmark = pytest.mark.gpu
C = Coordinates.cartesian


def _mesh(nx, bnx, bcs=None, ng=4):
    return UniformMesh(nx=nx, xmin=(0.0, 0.0, 0.0), xmax=(1.0, 0.8, 0.6), block_nx=bnx,
                       nghost=ng, bcs=tuple(bcs or (BoundaryFlag.periodic,) * 6), coords=C)


SINGLE_PASS_KERNELS = ["single_pass", "role_split"]   # sweep.cuh | trio.cuh


def _twin(mesh, gp, dp, variant, integ, ncyc, device_resident=False, seed=5, shocks=True,
          path="single_pass"):
    osim = OracleSim(mesh, gas=gp, dust=dp, integrator=integ)
    md = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
    md.set_stage_path(path)
    assert md.stage_path() == path
    for which, fp in ((Fluid.gas, gp), (Fluid.dust, dp)):
        if fp is None:
            continue
        p = random_prim(mesh, fp, seed=seed + int(which), shocks=shocks)
        (osim.gas if which == Fluid.gas else osim.dust).prim[:] = p
        md.fluid(which).prim.set(p)
    osim.nlim = ncyc
    osim.initialize()
    osim.run()
    drv = ArtemisDriver(md, integ, mode="fused", nlim=ncyc)
    drv.Initialize()
    if device_resident:
        md.set_time_state(drv.dt)
        code = {"rk1": 0, "rk2": 1, "vl2": 2, "rk3": 3}[integ]
        md.call("ab200_run_cycles", code, ncyc, float(np.finfo(np.float64).max))
    else:
        drv.Execute()
    res = [(df.u0.get(), df.prim.get(), of.u0, of.prim, of.fp)
           for of, df in zip(osim.fluids, md.fluids)]
    launches = md.launch_count()
    md.close()
    return res, launches


@pytest.mark.parametrize("path", SINGLE_PASS_KERNELS)
@pytest.mark.parametrize("rs", ["hllc", "hlle", "llf"])
@pytest.mark.parametrize("recon,integ", [("ppm", "rk2"), ("plm", "vl2"), ("ppm", "rk3"),
                                         ("plm", "rk1")])
def test_sweep_strict_is_bit_identical_to_oracle(recon, integ, rs, path):
    """The single-pass kernel sums the three flux differences and applies the sources in the
    reference's own order, so the strict build reproduces the oracle bit for bit.  32^3 blocks
    hold 2x2 tiles each (tile-to-tile halos inside a block); 2 gas + 2 dust species."""
    mesh = _mesh((64, 32, 32), (32, 32, 32))
    gp = gas_params(C, recon, rs, S=2)
    dp = dust_params(C, recon, "llf" if rs == "llf" else "hlle", S=2)
    res, _ = _twin(mesh, gp, dp, "strict", integ, 2, path=path)
    for u0, prim, ou0, oprim, _ in res:
        assert np.array_equal(u0, ou0)
        assert np.array_equal(prim, oprim)


@pytest.mark.parametrize("path", SINGLE_PASS_KERNELS)
@pytest.mark.parametrize("bnx", [(16, 16, 16), (8, 6, 4), (32, 16, 8), (24, 20, 12)])
@pytest.mark.parametrize("bc", ["periodic", "outflow", "reflect"])
def test_sweep_fast_within_1e12_one_cycle(bnx, bc, path):
    """Default (FMA, fast division) build, full and ragged tiles, every boundary kind."""
    B = BoundaryFlag
    bcs = {"periodic": (B.periodic,) * 6, "outflow": (B.outflow,) * 6,
           "reflect": (B.reflect,) * 6}[bc]
    mesh = _mesh(tuple(2 * b for b in bnx), bnx, bcs)
    gp = gas_params(C, "ppm", "hllc")
    dp = dust_params(C, "plm", "hlle", S=2)
    res, _ = _twin(mesh, gp, dp, "fast", "rk2", 1, path=path)
    for u0, prim, ou0, oprim, fp in res:
        assert zone_rel_err(u0, ou0, fp, "cons") <= 1e-12
        assert zone_rel_err(prim, oprim, fp, "prim") <= 1e-12


@pytest.mark.parametrize("path", SINGLE_PASS_KERNELS)
@pytest.mark.parametrize("integ", ["rk1", "rk2", "vl2", "rk3"])
def test_sweep_device_resident_pingpong_matches_oracle(integ, path):
    """ab200_run_cycles leaves the primitives in the alternate set between stages (no copy
    back); after 3 cycles (odd and even stage counts) the caller's arrays hold the result."""
    mesh = _mesh((32, 32, 32), (16, 16, 16), (BoundaryFlag.outflow,) * 6)
    gp = gas_params(C, "ppm", "hllc")
    res, launches = _twin(mesh, gp, None, "strict", integ, 3, device_resident=True, path=path)
    for u0, prim, ou0, oprim, _ in res:
        assert np.array_equal(u0, ou0)
        assert np.array_equal(prim, oprim)
    assert launches > 0


def test_sweep_blast_hundred_cycles_and_single_launch_per_stage():
    """Config-2 numerics at 32^3: 100 rk2 cycles within 1e-9 of the oracle, and the stage is
    ONE kernel: per cycle 2 stage kernels + 2 ghost fills + dt bookkeeping."""
    mesh = UniformMesh(nx=(32, 32, 32), xmin=(-1, -1, -1), xmax=(1, 1, 1), block_nx=(16, 16, 16),
                       nghost=4, bcs=(BoundaryFlag.outflow,) * 6)
    gp = FluidParams(Fluid.gas, C, ReconstructionMethod.ppm, RSolver.hllc, cfl=0.3, nspecies=1,
                     dfloor=1e-10, gamma=1.4, siefloor=1e-10)
    prim = pgen.blast(mesh, gp.gamma, d0=1.0, p0=1e-5, internal_energy=1.0, radius=0.2, samples=0)
    osim = OracleSim(mesh, gas=gp)
    osim.gas.prim[:] = prim
    osim.nlim = 100
    osim.initialize()
    osim.run()
    md = MeshData(mesh, gas=gp, materialize_fluxes=False)
    md.set_stage_path("single_pass")
    md.gas.prim.set(prim)
    drv = ArtemisDriver(md, "rk2", mode="fused")
    drv.Initialize()
    md.set_time_state(drv.dt)
    n0 = md.launch_count()
    md.call("ab200_run_cycles", 1, 100, float(np.finfo(np.float64).max))
    per_cycle = (md.launch_count() - n0) / 100.0
    assert zone_rel_err(md.gas.u0.get(), osim.gas.u0, gp, "cons") <= 1e-9
    assert zone_rel_err(md.gas.prim.get(), osim.gas.prim, gp, "prim") <= 1e-9
    ts = md.time_state()
    assert ts[3] == 100 and abs(ts[2] - osim.time) <= 1e-12 * osim.time
    assert per_cycle <= 8, per_cycle
    md.close()


@pytest.mark.parametrize("bc", ["reflect", "outflow"])
def test_paths_agree_and_can_be_switched_mid_run(bc):
    """ab200_set_stage_path: the single-pass kernel and the three directional passes agree to
    the parity bar (the directional passes round the flux divergence once per direction, the
    single-pass kernel sums it first like the reference, so only the latter is bit-identical
    to the oracle), also when the path is switched while the primitives sit in the alternate
    set.  Gas PPM+HLLC and 2 dust species, reflecting walls included."""
    B = BoundaryFlag
    bcs = (B.reflect,) * 6 if bc == "reflect" else (B.outflow,) * 6
    mesh = _mesh((32, 32, 32), (16, 16, 16), bcs)
    gp = gas_params(C, "ppm", "hllc")
    dp = dust_params(C, "plm", "hlle", S=2)
    out = {}
    for path in ("three_pass", "single_pass", "switch"):
        md = MeshData(mesh, gas=gp, dust=dp, variant="strict", materialize_fluxes=False)
        md.set_stage_path("single_pass" if path == "switch" else path)
        assert md.stage_path() == ("three_pass" if path == "three_pass" else "single_pass")
        md.gas.prim.set(random_prim(mesh, gp, seed=3))
        md.dust.prim.set(random_prim(mesh, dp, seed=4))
        drv = ArtemisDriver(md, "rk3", mode="fused")
        drv.Initialize()
        md.set_time_state(drv.dt)
        big = float(np.finfo(np.float64).max)
        if path == "switch":
            md.call("ab200_fused_stage", 0.0, 1.0, 1.0, 0.0, 0, 1, 1 | 4)  # prim left in the alternate set
            md.call("ab200_fill_ghosts")
            md.set_stage_path("three_pass")
            md.call("ab200_fused_stage", 0.25, 0.75, 0.25, 0.0, 0, 0, 1 | 4)
            md.call("ab200_fill_ghosts")
            md.set_stage_path("single_pass")
            md.call("ab200_fused_stage", 2.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0, 0.0, 0, 0, 1 | 2 | 4)
            md.call("ab200_fill_ghosts")
            md.call("ab200_set_global_timestep_device", big, 1)
            md.call("ab200_sync_prim")
        else:
            md.call("ab200_run_cycles", 3, 1, big)
        out[path] = [(f.u0.get(), f.prim.get(), f.fp) for f in md.fluids] + [md.time_state()]
        md.close()
    for path in ("single_pass", "switch"):
        for (u_a, p_a, fp), (u_b, p_b, _) in zip(out["three_pass"][:2], out[path][:2]):
            assert zone_rel_err(u_b, u_a, fp, "cons") <= 1e-12, path
            assert zone_rel_err(p_b, p_a, fp, "prim") <= 1e-12, path
        ta, tb = out["three_pass"][2], out[path][2]
        assert ta[3] == tb[3] == 1 and abs(ta[0] - tb[0]) <= 1e-13 * ta[0], path
