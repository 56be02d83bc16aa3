"""Multi-rank fused path on ONE GPU: every rank of a 2x2x2 / 2x1x1 / 1x2x2 rank lattice is its
own ab200 context holding one tile of the block lattice; an in-process loopback stands in for
NCCL (message buffers are copied device to device).  After whole cycles of
  ab200_fused_stage -> ab200_fill_ghosts_local -> pack / transfer / unpack sweeps ->
  ab200_finish_remote_ghosts -> dt min over ranks -> ab200_set_global_timestep_device
every rank's arrays (ghost zones included) must be BIT-IDENTICAL to the single-context run of
the undivided mesh through ab200_run_cycles (which is itself checked against the oracle in
test_gpu_cycle.py)."""
import numpy as np
import pytest

from artemis_b200.comm import HaloComm, rank_coords
from artemis_b200.driver import ArtemisDriver
from artemis_b200.enums import BoundaryFlag, Coordinates
from artemis_b200.mesh import UniformMesh
from artemis_b200.meshdata import MeshData
from tests.helpers import dust_params, gas_params, random_prim

pytestmark = pytest.mark.gpu
B = BoundaryFlag
BIG = float(np.finfo(np.float64).max)


class _NoDist:
    """HaloComm only needs a transport for transfer_sweep, which the test replaces."""


def _tile_bcs(bcs, lay, rl, periodic):
    tbc = [int(v) for v in bcs]
    for d in range(3):
        if lay[d] > 1 and (rl[d] > 0 or periodic[d]):
            tbc[2 * d] = 3
        if lay[d] > 1 and (rl[d] < lay[d] - 1 or periodic[d]):
            tbc[2 * d + 1] = 3
    return tbc


@pytest.mark.parametrize("lay,bc,with_dust,integ", [
    ((2, 2, 2), "outflow", False, "rk2"),
    ((2, 1, 1), "periodic", True, "vl2"),
    ((1, 2, 2), "reflect", True, "rk3"),
    ((2, 2, 1), "mixed", False, "rk2"),
])
@pytest.mark.parametrize("scheme", ["direct", "sweeps"])
def test_multirank_cycles_bit_identical_to_single_context(lay, bc, with_dust, integ, scheme):
    bcs = {"outflow": (B.outflow,) * 6, "periodic": (B.periodic,) * 6, "reflect": (B.reflect,) * 6,
           "mixed": (B.reflect, B.outflow, B.periodic, B.periodic, B.outflow, B.reflect)}[bc]
    periodic = tuple(bcs[2 * d] == B.periodic for d in range(3))
    nblk = tuple(2 * lay[d] for d in range(3))
    bnx = (8, 6, 4)
    gm = UniformMesh(nx=tuple(nblk[d] * bnx[d] for d in range(3)), xmin=(0, 0, 0),
                     xmax=(1.0, 0.8, 0.6), block_nx=bnx, nghost=4, bcs=bcs)
    gp = gas_params(Coordinates.cartesian, "ppm", "hllc")
    dp = dust_params(Coordinates.cartesian, "plm", "hlle", S=2) if with_dust else None
    prim = random_prim(gm, gp, seed=21)
    dprim = random_prim(gm, dp, seed=22) if with_dust else None
    ncyc = 3
    code = {"rk1": 0, "rk2": 1, "vl2": 2, "rk3": 3}[integ]

    # ---- single context, undivided mesh --------------------------------------------------
    md = MeshData(gm, gas=gp, dust=dp, materialize_fluxes=False)
    md.gas.prim.set(prim)
    if with_dust:
        md.dust.prim.set(dprim)
    drv = ArtemisDriver(md, integ, mode="fused")
    drv.Initialize()
    dt0 = drv.dt
    md.set_time_state(dt0)
    md.call("ab200_run_cycles", code, ncyc, BIG)
    want_ts = md.time_state()
    want = [(f.prim.get(), f.u0.get()) for f in md.fluids]
    md.close()

    # ---- one context per rank ----------------------------------------------------------------
    world = lay[0] * lay[1] * lay[2]
    ranks = []
    for r in range(world):
        rl = rank_coords(r, lay)
        nbt = tuple(gm.nrb[d] // lay[d] for d in range(3))
        lo = tuple(rl[d] * nbt[d] for d in range(3))
        tm = UniformMesh(nx=gm.nx, xmin=gm.xmin, xmax=gm.xmax, block_nx=gm.block_nx, nghost=4,
                         bcs=bcs, lattice_lo=lo, lattice_n=nbt)
        gid = [int(l[0] + gm.nrb[0] * (l[1] + gm.nrb[1] * l[2])) for l in tm.blk_loc]
        tmd = MeshData(tm, gas=gp, dust=dp, materialize_fluxes=False,
                       bcs=_tile_bcs(bcs, lay, rl, periodic))
        comm = HaloComm(tmd, lay, rl, r, world, periodic=periodic, dist=_NoDist())
        tmd.gas.prim.set(np.ascontiguousarray(prim[gid]))
        if with_dust:
            tmd.dust.prim.set(np.ascontiguousarray(dprim[gid]))
        ranks.append((tmd, comm, gid))

    def sweeps():
        for d in range(3):
            if all(c.plans[d][0] is None and c.plans[d][1] is None for _, c, _ in ranks):
                continue
            for _, c, _ in ranks:
                c.pack_sweep(d)
            for tmd, c, _ in ranks:
                tmd.synchronize()
            for _, c, _ in ranks:       # loopback "NCCL": my send buffer -> the peer's receive
                for side, p in enumerate(c.plans[d]):
                    if p is not None:
                        ranks[p.peer][1].bufs[d][1 - side][1].copy_(c.bufs[d][side][0])
            import torch
            torch.cuda.synchronize()
            for _, c, _ in ranks:
                c.unpack_sweep(d)

    def direct():
        """single-round scheme: every rank packs once, the loopback copies each peer slice into
        the peer's receive slice for this rank, every rank unpacks once"""
        import torch
        for _, c, _ in ranks:
            if c._dslices:
                c.backend.pack(c._dsend_items, c.dsend)
        for tmd, c, _ in ranks:
            tmd.synchronize()
        for r, (_, c, _) in enumerate(ranks):
            for peer, sb, _ in c._dslices:
                back = [rb for pr, _, rb in ranks[peer][1]._dslices if pr == r]
                assert len(back) == 1 and back[0].numel() == sb.numel()
                back[0].copy_(sb)
        torch.cuda.synchronize()
        for _, c, _ in ranks:
            if c._dslices:
                c.backend.unpack(c._drecv_items, c.drecv)

    if scheme == "direct":
        sweeps = direct  # noqa: F811

    # Mesh::Initialize: PrimToCons, exchange, PrimToCons (driver.py Initialize, split by phase)
    for tmd, _, _ in ranks:
        tmd.call("ab200_prim_to_cons")
        tmd.call("ab200_cons_to_prim")
        tmd.call("ab200_fill_ghosts_local")
    sweeps()
    for tmd, _, _ in ranks:
        tmd.call("ab200_finish_remote_ghosts")
        tmd.call("ab200_prim_to_cons")
        tmd.set_time_state(dt0)
    integ_tab = drv.integrator
    for _ in range(ncyc):
        for stage in range(1, integ_tab.nstages + 1):
            flags = 1 | (2 if stage == integ_tab.nstages else 0)
            pcm = int(stage == 1 and integ == "vl2")
            for tmd, _, _ in ranks:
                tmd.call("ab200_fused_stage", integ_tab.gam0[stage - 1], integ_tab.gam1[stage - 1],
                         integ_tab.beta[stage - 1], 0.0, pcm, int(stage == 1), flags)
                tmd.call("ab200_fill_ghosts_local")
            sweeps()
            for tmd, _, _ in ranks:
                tmd.call("ab200_finish_remote_ghosts")
        ts = [tmd.time_state() for tmd, _, _ in ranks]
        new_dt = min(t[1] for t in ts)      # the all-reduce(MIN)
        for (tmd, _, _), t in zip(ranks, ts):
            tmd.set_time_state(t[0], new_dt, t[2], int(t[3]))
            tmd.call("ab200_set_global_timestep_device", BIG, 1)
    for tmd, _, gid in ranks:
        ts = tmd.time_state()
        assert ts[3] == ncyc and ts[0] == want_ts[0] and ts[2] == want_ts[2]
        for f, (wp, wu) in zip(tmd.fluids, want):
            assert np.array_equal(f.prim.get(), wp[gid])
            assert np.array_equal(f.u0.get(), wu[gid])
        tmd.close()
