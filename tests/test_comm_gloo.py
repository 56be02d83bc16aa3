"""World-size-2 gloo test of the cross-rank ghost exchange planning + transport (CPU only).

Two processes each own one tile of a block lattice held in numpy; a numpy stand-in replaces
the two C-ABI kernels (pack/unpack of [comp][k][j][i] sub-boxes).  After
  same-rank neighbour fill -> HaloComm.exchange (3 sweeps over gloo) -> physical BCs
every ghost zone must equal the single-process oracle exchange on the undivided mesh, and the
dt all-reduce must return the global minimum."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from artemis_b200.comm import HaloComm, plan_direct, plan_sweeps, rank_coords
from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
from artemis_b200.mesh import UniformMesh
from artemis_b200.params import FluidParams


class NumpyBackend:
    """Test-only stand-in for ab200_halo_pack / ab200_halo_unpack."""

    def __init__(self, prims):
        self.prims = prims  # {fluid int: array [nb][nvar][nk][nj][ni]}

    def alloc(self, n):
        return torch.zeros(max(n, 1), dtype=torch.float64)

    def pack(self, items, tensor):
        buf = tensor.numpy()
        for fl, b, v0, nc, si, ei, sj, ej, sk, ek, off in items:
            box = self.prims[fl][b, v0:v0 + nc, sk:ek + 1, sj:ej + 1, si:ei + 1]
            buf[off:off + box.size] = box.ravel()

    def unpack(self, items, tensor):
        buf = tensor.numpy()
        for fl, b, v0, nc, si, ei, sj, ej, sk, ek, off in items:
            shp = (nc, ek - sk + 1, ej - sj + 1, ei - si + 1)
            n = int(np.prod(shp))
            self.prims[fl][b, v0:v0 + nc, sk:ek + 1, sj:ej + 1, si:ei + 1] = \
                buf[off:off + n].reshape(shp)

    def allreduce_min(self, dist_, value):
        t = torch.tensor([value], dtype=torch.float64)
        dist_.all_reduce(t, op=dist_.ReduceOp.MIN)
        return float(t.item())


class _FF:
    def __init__(self, fp):
        self.fp = fp


class _FakeMD:
    def __init__(self, mesh, fps):
        self.mesh = mesh
        self.fluids = [_FF(fp) for fp in fps]
        self.device = 0


def _fps():
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod.ppm, RSolver.hllc,
                     cfl=0.3, nspecies=2, dfloor=1e-10, gamma=1.4)
    dp = FluidParams(Fluid.dust, Coordinates.cartesian, ReconstructionMethod.plm, RSolver.hlle,
                     cfl=0.3, nspecies=1, dfloor=1e-10)
    return [gp, dp]


def _global_mesh(bcs, lay):
    return UniformMesh(nx=(8 * lay[0], 4 * lay[1], 4 * lay[2]), xmin=(0, 0, 0), xmax=(1, 1, 1),
                       block_nx=(4, 4, 4), nghost=2, bcs=bcs)


def _ghost_lists(fp):
    from oracle.oracle_py import FluidState
    S = fp.nspecies
    if fp.fluid_type == Fluid.gas:
        gv = list(range(0, 4 * S)) + list(range(5 * S, 6 * S))
    else:
        gv = list(range(0, 4 * S))
    vd = [((v - S) % 3 + 1) if S <= v < 4 * S else 0 for v in gv]
    return np.array(gv, dtype=np.int32), np.array(vd, dtype=np.int32)


def _oracle_exchange(mesh, fp, prim, bc, phases):
    import ctypes as C
    from oracle import oracle_py
    L = oracle_py.lib()
    g = oracle_py.make_grid(mesh)
    gv, vd = _ghost_lists(fp)
    bc = np.array([int(v) for v in bc], dtype=np.int32)
    IP = C.POINTER(C.c_int)
    L.ao_exchange_ghosts_phase(C.byref(g), *[int(v) for v in mesh.lattice_n],
                               bc.ctypes.data_as(IP), fp.nvar,
                               prim.ctypes.data_as(C.POINTER(C.c_double)), len(gv),
                               gv.ctypes.data_as(IP), vd.ctypes.data_as(IP), phases)


def _pure_remote_mask(tm, tbc):
    """[nb] boolean masks [nk][nj][ni] of the ghost cells the direct scheme delivers: ghost
    index only along directions whose face belongs to another rank, interior along the rest."""
    s = (tm.is_, tm.js, tm.ks)
    e = (tm.ie, tm.je, tm.ke)
    nt = (tm.ni, tm.nj, tm.nk)
    nbd = tuple(tm.lattice_n)
    masks = []
    for b in range(tm.nb):
        l = (b % nbd[0], (b // nbd[0]) % nbd[1], b // (nbd[0] * nbd[1]))
        ok = np.ones((nt[2], nt[1], nt[0]), dtype=bool)
        ghost_any = np.zeros_like(ok)
        for d in range(3):
            idx = np.arange(nt[d])
            shape = [1, 1, 1]
            shape[2 - d] = nt[d]
            lo_g = (idx < s[d]).reshape(shape)
            hi_g = (idx > e[d]).reshape(shape)
            lo_remote = l[d] == 0 and tbc[2 * d] == 3
            hi_remote = l[d] == nbd[d] - 1 and tbc[2 * d + 1] == 3
            ok &= ~(lo_g & (not lo_remote)) & ~(hi_g & (not hi_remote))
            ghost_any |= lo_g | hi_g
        masks.append(ok & ghost_any)
    return masks


def _worker(rank, world, lay, bc_name, port, q, scheme="sweeps"):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        bcs = (BoundaryFlag[bc_name],) * 6
        periodic = tuple(bc_name == "periodic" for _ in range(3))
        gm = _global_mesh(bcs, lay)
        fps = _fps()
        rng = np.random.default_rng(7)
        gprims = [rng.normal(size=gm.shape(fp.nvar)) for fp in fps]
        # global answer: the single-process oracle exchange
        want = [p.copy() for p in gprims]
        for fp, p in zip(fps, want):
            _oracle_exchange(gm, fp, p, bcs, 3)
        # this rank's tile
        rl = rank_coords(rank, lay)
        nbt = tuple(gm.nrb[d] // lay[d] for d in range(3))
        lo = tuple(rl[d] * nbt[d] for d in range(3))
        tm = UniformMesh(nx=gm.nx, xmin=gm.xmin, xmax=gm.xmax, block_nx=gm.block_nx, nghost=2,
                         bcs=bcs, lattice_lo=lo, lattice_n=nbt)
        gid = []
        for b in range(tm.nb):
            l = tm.blk_loc[b]
            gid.append(int(l[0] + gm.nrb[0] * (l[1] + gm.nrb[1] * l[2])))
        tprims = {}
        for fp, p in zip(fps, gprims):
            t = p[gid].copy()
            # poison the ghosts so nothing stale can pass
            mask = np.ones(t.shape[2:], dtype=bool)
            mask[tm.interior()] = False
            t[:, :, mask] = np.nan
            tprims[int(fp.fluid_type)] = np.ascontiguousarray(t)
        tbc = list(int(v) for v in bcs)
        for d in range(3):
            if lay[d] > 1 and (rl[d] > 0 or periodic[d]):
                tbc[2 * d] = 3
            if lay[d] > 1 and (rl[d] < lay[d] - 1 or periodic[d]):
                tbc[2 * d + 1] = 3
        md = _FakeMD(tm, fps)
        comm = HaloComm(md, lay, rl, rank, world, backend=NumpyBackend(tprims),
                        periodic=periodic, dist=dist)
        for fp in fps:   # same-rank neighbours
            _oracle_exchange(tm, fp, tprims[int(fp.fluid_type)], tbc, 1)
        ok = True
        if scheme == "direct":
            # single round; delivers exactly the pure-remote ghost cells (the rest is resolved
            # on the device by ab200_finish_remote_ghosts, covered by tests/test_gpu_multirank.py)
            comm.exchange_direct()
            masks = _pure_remote_mask(tm, tbc)
            for fp, w in zip(fps, want):
                got = tprims[int(fp.fluid_type)]
                gv, _ = _ghost_lists(fp)
                for b in range(tm.nb):
                    ok = ok and masks[b].any()
                    ok = ok and np.array_equal(got[b][gv][:, masks[b]], w[gid[b]][gv][:, masks[b]])
        else:
            comm.exchange()
            for fp in fps:   # physical boundaries
                _oracle_exchange(tm, fp, tprims[int(fp.fluid_type)], tbc, 2)
            for fp, w in zip(fps, want):
                got = tprims[int(fp.fluid_type)]
                gv, _ = _ghost_lists(fp)
                ok = ok and np.array_equal(got[:, gv], w[gid][:, gv])
        dtmin = comm.allreduce_min(1.0 + rank)
        q.put((rank, bool(ok), dtmin, comm.bytes_per_exchange if scheme == "sweeps"
               else comm.bytes_per_direct_exchange))
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("lay,bc_name", [((2, 1, 1), "outflow"), ((2, 1, 1), "periodic"),
                                         ((1, 2, 1), "reflect"), ((1, 1, 2), "periodic")])
def test_two_rank_exchange_matches_single_process_oracle(lay, bc_name, oracle_lib):
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000) + 7 * sum(lay) + len(bc_name)
    procs = [ctx.Process(target=_worker, args=(r, 2, lay, bc_name, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(2)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, dtmin, nbytes in res:
        assert ok, f"rank {rank}: ghost zones differ from the single-process exchange"
        assert dtmin == 1.0
        assert nbytes > 0


def test_plan_is_symmetric_between_peers():
    """What rank A packs for B has exactly the shape B expects to unpack from A."""
    lay = (2, 2, 2)
    fps = _fps()
    fl = [(fp.fluid_type, fp.nspecies) for fp in fps]
    plans = {}
    for r in range(8):
        rl = rank_coords(r, lay)
        gm = _global_mesh((BoundaryFlag.outflow,) * 6, lay)
        nbt = tuple(gm.nrb[d] // lay[d] for d in range(3))
        tm = UniformMesh(nx=gm.nx, xmin=gm.xmin, xmax=gm.xmax, block_nx=gm.block_nx, nghost=2,
                         bcs=gm.bcs, lattice_lo=tuple(rl[d] * nbt[d] for d in range(3)),
                         lattice_n=nbt)
        plans[r] = plan_sweeps(tm, fl, lay, rl)
    npeers = 0
    for r in range(8):
        for d in range(3):
            for side in (0, 1):
                p = plans[r][d][side]
                if p is None:
                    continue
                npeers += 1
                q = plans[p.peer][d][1 - side]
                assert q is not None and q.peer == r and q.nelem == p.nelem
                for a, b in zip(p.send, q.recv):
                    # same fluid/vars/box shape and buffer offset on both ends
                    assert a[0] == b[0] and a[2:4] == b[2:4] and a[10] == b[10]
                    assert [a[5] - a[4], a[7] - a[6], a[9] - a[8]] == \
                           [b[5] - b[4], b[7] - b[6], b[9] - b[8]]
    assert npeers == 8 * 3   # 2x2x2, non-periodic: every rank has exactly 3 face peers


@pytest.mark.parametrize("lay,bc_name", [((2, 1, 1), "periodic"), ((2, 2, 1), "outflow"),
                                         ((1, 2, 2), "periodic"), ((2, 2, 1), "periodic"),
                                         ((2, 2, 2), "reflect")])
def test_direct_exchange_delivers_every_pure_remote_ghost_cell(lay, bc_name, oracle_lib):
    """Single-round scheme over gloo with 2, 4 and 8 ranks: faces, rank edges (two remote
    directions), the rank corner (three) and the periodic wrap onto one peer from both sides."""
    world = lay[0] * lay[1] * lay[2]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000) + 11 * sum(lay) + len(bc_name)
    procs = [ctx.Process(target=_worker, args=(r, world, lay, bc_name, port, q, "direct"))
             for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=180) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, ok, dtmin, nbytes in res:
        assert ok, f"rank {rank}: delivered ghost zones differ from the single-process exchange"
        assert dtmin == 1.0 and nbytes > 0


def test_direct_plan_is_symmetric_between_peers():
    """2x2x2 lattice: 7 peers per rank (3 faces, 3 edges, 1 corner); what A packs for B is
    exactly what B expects from A, box by box."""
    lay = (2, 2, 2)
    fps = _fps()
    fl = [(fp.fluid_type, fp.nspecies) for fp in fps]
    plans = {}
    for r in range(8):
        rl = rank_coords(r, lay)
        gm = _global_mesh((BoundaryFlag.outflow,) * 6, lay)
        nbt = tuple(gm.nrb[d] // lay[d] for d in range(3))
        tm = UniformMesh(nx=gm.nx, xmin=gm.xmin, xmax=gm.xmax, block_nx=gm.block_nx, nghost=2,
                         bcs=gm.bcs, lattice_lo=tuple(rl[d] * nbt[d] for d in range(3)),
                         lattice_n=nbt)
        plans[r] = {p.peer: p for p in plan_direct(tm, fl, lay, rl)}
    for r in range(8):
        assert len(plans[r]) == 7
        for peer, p in plans[r].items():
            q = plans[peer][r]
            assert p.nsend == q.nrecv and p.nrecv == q.nsend and len(p.send) == len(q.recv)
            for a, b in zip(p.send, q.recv):
                assert a[0] == b[0] and a[2:4] == b[2:4] and a[10] == b[10]
                assert [a[5] - a[4], a[7] - a[6], a[9] - a[8]] == \
                       [b[5] - b[4], b[7] - b[6], b[9] - b[8]]
    # the single round moves less than the three forwarding sweeps (no tangential ghosts)
    rl = rank_coords(0, lay)
    sweep_elems = sum(p.nelem for pair in plan_sweeps(tm, fl, lay, rank_coords(7, lay)) for p in pair if p)
    assert sum(p.nsend for p in plans[7].values()) < sweep_elems
