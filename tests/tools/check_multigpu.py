"""N-rank check of the device-resident cycle on REAL GPUs (one process per GPU, NCCL): after a
few cycles every rank's tile must be bit-identical to the undivided mesh advanced by one
context with ab200_run_cycles.  Complements tests/test_gpu_multirank.py (loopback transport on
one GPU), which cannot see stream-ordering mistakes of the overlapped exchange.

  python -m torch.distributed.run --nproc-per-node N tests/tools/check_multigpu.py [--cycles 3]
"""
import argparse
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from artemis_b200.comm import HaloComm, NativeComm, rank_coords  # noqa: E402
from artemis_b200.driver import ArtemisDriver  # noqa: E402
from artemis_b200.enums import BoundaryFlag, Coordinates  # noqa: E402
from artemis_b200.mesh import UniformMesh  # noqa: E402
from artemis_b200.meshdata import MeshData  # noqa: E402
from tests.helpers import dust_params, gas_params, random_prim  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--cycles", type=int, default=3)
ap.add_argument("--bnx", type=int, default=16)
ap.add_argument("--transport", default="native", choices=["native", "torch"],
                help="native: ab200_run_cycles_mr over the C ABI's own NCCL transport (comm.cu); "
                     "torch: StepDevice driven from Python over torch.distributed")
ap.add_argument("--physics", action="store_true",
                help="also configure uniform gravity, gas-dust drag, viscosity and conduction "
                     "(split stages + diffusion operators on every rank)")
args = ap.parse_args()
world, rank, local = (int(os.environ[k]) for k in ("WORLD_SIZE", "RANK", "LOCAL_RANK"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
lay = {2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[world]
B = BoundaryFlag
bcs = (B.reflect, B.outflow, B.outflow, B.reflect, B.outflow, B.outflow)
nblk = tuple(2 * lay[d] for d in range(3))
bnx = (args.bnx,) * 3
gm = UniformMesh(nx=tuple(nblk[d] * bnx[d] for d in range(3)), xmin=(0, 0, 0), xmax=(1.0, 0.8, 0.6),
                 block_nx=bnx, nghost=4, bcs=bcs)
C = Coordinates.cartesian
gp, dp = gas_params(C, "ppm", "hllc"), dust_params(C, "plm", "hlle", S=1)
prim, dprim = random_prim(gm, gp, seed=21, shocks=True), random_prim(gm, dp, seed=22, shocks=False)
BIG = float(np.finfo(np.float64).max)



def configure_physics(m):
    """the same source terms and diffusion operators on the undivided mesh and on every tile"""
    if not args.physics:
        return
    import ctypes as C
    from artemis_b200 import capi
    sd = capi.SourcesDesc()
    sd.gravity, sd.g[0], sd.g[1], sd.g[2] = 1, 0.1, -0.2, 0.3
    sd.drag, sd.ntau, sd.tau[0] = 1, 1, 0.05
    m.call("ab200_configure_sources", C.byref(sd))
    dd = capi.DiffusionDesc()
    dd.visc_type, dd.nu, dd.r0, dd.eta_bulk = 1, 3e-3, 1.0, 0.4
    dd.cond_type, dd.kappa, dd.rho_ref, dd.t_ref, dd.cv = 2, 5e-3, 1.0, 1.0, 1.1
    m.call("ab200_configure_diffusion", C.byref(dd))


# the undivided mesh on this rank's own GPU (every rank computes it: no gather needed)
md = MeshData(gm, gas=gp, dust=dp, device=local, materialize_fluxes=False)
md.gas.prim.set(prim)
md.dust.prim.set(dprim)
configure_physics(md)
drv = ArtemisDriver(md, "rk2", mode="fused")
drv.Initialize()
dt0 = drv.dt
md.set_time_state(dt0)
md.call("ab200_run_cycles", 1, args.cycles, BIG)
want = [(f.prim.get(), f.u0.get()) for f in md.fluids]
want_ts = md.time_state()
md.close()

rl = rank_coords(rank, lay)
nbt = tuple(gm.nrb[d] // lay[d] for d in range(3))
tm = UniformMesh(nx=gm.nx, xmin=gm.xmin, xmax=gm.xmax, block_nx=gm.block_nx, nghost=4, bcs=bcs,
                 lattice_lo=tuple(rl[d] * nbt[d] for d in range(3)), lattice_n=nbt)
gid = [int(l[0] + gm.nrb[0] * (l[1] + gm.nrb[1] * l[2])) for l in tm.blk_loc]
tbc = [int(v) for v in bcs]
for d in range(3):
    if lay[d] > 1 and rl[d] > 0:
        tbc[2 * d] = 3
    if lay[d] > 1 and rl[d] < lay[d] - 1:
        tbc[2 * d + 1] = 3
tmd = MeshData(tm, gas=gp, dust=dp, device=local, materialize_fluxes=False, bcs=tbc)
comm = HaloComm(tmd, lay, rl, rank, world)
tmd.gas.prim.set(np.ascontiguousarray(prim[gid]))
tmd.dust.prim.set(np.ascontiguousarray(dprim[gid]))
configure_physics(tmd)
tdrv = ArtemisDriver(tmd, "rk2", mode="fused", comm=comm)
tdrv.Initialize()
assert tdrv.dt == dt0, (tdrv.dt, dt0)
host0 = [f.prim.get() for f in tmd.fluids]   # this rank's tile with consistent ghost zones
tmd.set_time_state(tdrv.dt)
if args.transport == "native":
    ncomm = NativeComm(tmd, lay, rank, world)
    tmd.call("ab200_run_cycles_mr", 1, args.cycles, BIG)
else:
    tdrv.BeginDeviceResident()
    for _ in range(args.cycles):
        tdrv.StepDevice()
    tdrv.EndDeviceResident()
tmd.synchronize()
torch.cuda.synchronize()
ts = tmd.time_state()
ok = bool(ts[3] == args.cycles and ts[0] == want_ts[0] and ts[2] == want_ts[2])
for f, (wp, wu) in zip(tmd.fluids, want):
    ok = ok and np.array_equal(f.prim.get(), wp[gid]) and np.array_equal(f.u0.get(), wu[gid])
host_ok = None
if args.transport == "native":
    # the host-buffer entry point on the split mesh: every rank uploads the INTERIOR zones of its
    # tile from pinned memory (ghost zones poisoned: rebuilt on the device, remote faces through
    # the library's transport), runs the same cycles, downloads interior zones only
    import ctypes as C
    DP = C.POINTER(C.c_double)
    ghost = np.ones(host0[0].shape[2:], dtype=bool)
    ghost[tm.interior()] = False
    bufs = []
    for a in host0:
        t = torch.empty(a.shape, dtype=torch.float64).pin_memory()
        t.numpy()[:] = a
        t.numpy()[:, :, ghost] = -9.0e9
        bufs.append(t)
    tmd.call("ab200_set_host_transfer", 1 | 2 | 4)
    dt_io = C.c_double(dt0)
    tmd.call("ab200_cycles_host", 1, args.cycles, C.byref(dt_io), C.cast(bufs[0].data_ptr(), DP),
             None, C.cast(bufs[1].data_ptr(), DP), None)
    tmd.synchronize()
    tmd.call("ab200_set_host_transfer", 0)
    inner = (slice(None), slice(None)) + tm.interior()
    host_ok = bool(dt_io.value == want_ts[0])
    for t, (wp, wu) in zip(bufs, want):
        host_ok = host_ok and np.array_equal(t.numpy()[inner], wp[gid][inner])
        host_ok = host_ok and bool(np.all(t.numpy()[:, :, ghost] == -9.0e9))
    ok = ok and host_ok
flag = torch.tensor([1.0 if ok else 0.0], device=f"cuda:{local}")
dist.all_reduce(flag, op=dist.ReduceOp.MIN)
if rank == 0:
    direct = int(tmd.L.ab200_comm_is_direct(tmd.ctx)) if args.transport == "native" else 0
    print("check_multigpu: world %d lattice %s cycles %d transport %s physics %s (peer-write over "
          "CUDA IPC: %s, overlap: %s; ab200_cycles_host on the split mesh, interior-only "
          "transfers: %s) -> %s" % (world, lay, args.cycles, args.transport,
                                  "gravity+drag+viscosity+conduction" if args.physics else "hydro",
                                  bool(direct), bool(os.environ.get("AB200_OVERLAP")),
                                  "not run" if host_ok is None else ("ok" if host_ok else "MISMATCH on rank 0"),
                                  "BIT-IDENTICAL" if flag.item() == 1.0 else "MISMATCH"))
tmd.close()
dist.destroy_process_group()
sys.exit(0 if flag.item() == 1.0 else 1)
