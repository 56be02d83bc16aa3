#!/usr/bin/env python
"""bench.py -- zone-cycles/s of the Artemis gas stage update on B200 (config 2 of BASELINE.json).

Workload (N=1): 3-D Sedov blast, 256^3 zones per GPU as 64 MeshBlocks of 64^3, nghost=4,
PPM + HLLC, fp64, rk2, gamma=1.4, cfl=0.3, outflow boundaries (inputs/blast/blast.in with the
SURVEY 8d overrides).  N>1: weak scaling, one 256^3 tile per rank (block-spatial partition),
halo exchange + dt all-reduce over NCCL.  A "step" is one full integrator cycle
(2 stages + CFL reduction).  State (3.4 GB/GPU) is far larger than L2, so no flush is needed.

  python bench.py --gpus N --steps K --warmup W [--impl reference]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

ALG_BYTES_PER_ZONE_STAGE = 240.0   # R W6 + u0 6 + u1 6, W u0 6 + W 6 doubles
METRIC = "zone-cycles/sec, 3D PPM+HLLC fp64, 1/2/4/8 B200; % HBM roofline"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tile", type=int, default=256, help="zones per GPU per direction")
    ap.add_argument("--block", type=int, default=64)
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample", type=int, default=256,
                    help="CPU baseline mesh size per direction (default: the GPU arm's own 256^3 "
                         "single-GPU mesh, so both arms run the SAME config; the sample is bounded "
                         "in cycles, not in mesh size)")
    ap.add_argument("--cpu-cycles", type=int, default=2)
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--e2e-transfer", default="interior",
                    choices=["full", "interior", "interior_zc", "interior_dma"],
                    help="N=1 e2e leg: which zones of the pinned host arrays cross PCIe "
                         "(ab200_set_host_transfer): whole arrays; interior zones, in by a copy "
                         "kernel on the pinned array and out by strided DMA (the measured best); "
                         "copy kernels both ways; strided DMA both ways")
    ap.add_argument("--config", type=int, default=2, choices=[2, 3, 4, 5],
                    help="BASELINE.json config: 2 = the headline (3-D blast, PPM+HLLC, 256^3 per "
                         "GPU, weak scaling); 3 = gas + 4 dust species, PLM+HLLE, periodic, "
                         "--mesh^3 zones IN TOTAL split over the GPUs (strong scaling); 4 = spherical "
                         "3-D disk-like gas + dust, PPM+HLLE (WENO5 does not exist in the reference), "
                         "32^3 MeshBlocks, outflow, curvilinear geometry + geometric source terms")
    ap.add_argument("--mesh", type=int, default=512, help="config 3: total zones per direction")
    ap.add_argument("--dust-species", type=int, default=4)
    ap.add_argument("--no-sources", action="store_true",
                    help="config 4: without the deck's point-mass gravity and rotating frame")
    ap.add_argument("--no-flux-correction", action="store_true",
                    help="config 5: fused stage kernels without Parthenon's flux correction")
    ap.add_argument("--bcs", default="deck", choices=["generic", "deck"],
                    help="config 5: 'deck' = inputs/ssheet/ssheet.in's own user conditions (extrap on "
                         "x1 / x3, inflow on x2: strat.hpp) on fine arrays and coarse buffers; "
                         "'generic' = outflow x1 / x3, periodic x2")
    ap.add_argument("--no-drag", action="store_true",
                    help="config 3: leave Drag::DragSource out (it then stays on the reference path)")
    ap.add_argument("--state", default="blast", choices=["blast", "shocked"],
                    help="config 2 initial state: 'blast' = the configured deck (ambient gas is "
                         "quiescent, limiter / wave-speed branches are warp-uniform); 'shocked' = a "
                         "seeded multi-mode field with jumps and noise in every MeshBlock, so the "
                         "branches diverge inside warps (a second number beside the headline)")
    ap.add_argument("--transport", default="native", choices=["native", "torch"],
                    help="N > 1: 'native' = the C ABI's own NCCL transport (ab200_run_cycles_mr), "
                         "'torch' = torch.distributed driven from Python (round-1 path)")
    ap.add_argument("--path", default="auto", choices=["auto", "three_pass", "single_pass", "role_split"],
                    help="ab200_set_stage_path: which stage kernels run (auto = library policy)")
    return ap.parse_args()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def rank_layout(n):
    return {1: (1, 1, 1), 2: (2, 1, 1), 4: (2, 2, 1), 8: (2, 2, 2)}[n]


def make_problem(args, rank, world):
    from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
    from artemis_b200.mesh import UniformMesh
    from artemis_b200.params import FluidParams
    lay = rank_layout(world)
    T, B = args.tile, args.block
    nx = tuple(T * lay[d] for d in range(3))
    nbt = T // B
    rl = (rank % lay[0], (rank // lay[0]) % lay[1], rank // (lay[0] * lay[1]))
    mesh = UniformMesh(nx=nx, xmin=tuple(-1.0 * lay[d] for d in range(3)),
                       xmax=tuple(1.0 * lay[d] for d in range(3)), block_nx=(B, B, B), nghost=4,
                       bcs=(BoundaryFlag.outflow,) * 6,
                       lattice_lo=tuple(rl[d] * nbt for d in range(3)), lattice_n=(nbt,) * 3)
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod.ppm, RSolver.hllc,
                     cfl=0.3, nspecies=1, dfloor=1e-10, gamma=1.4, siefloor=1e-10)
    return mesh, gp, lay, rl


class ClockSampler(threading.Thread):
    """nvidia-smi clocks + throttle reasons during the timed region (B200_PROFILING.md)."""

    def __init__(self, gpu_index):
        super().__init__(daemon=True)
        self.idx = gpu_index
        self.stop_evt = threading.Event()
        self.rows = []

    def run(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
             "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
             "clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={q}", "--format=csv,noheader,nounits", "-lms", "100",
                 "-i", str(self.idx)], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except OSError:
            return
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])
            if self.stop_evt.is_set():
                break
        self.proc.terminate()

    def finish(self):
        self.stop_evt.set()
        time.sleep(0.15)
        try:
            self.proc.terminate()
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown",
                                "sw_power_cap"), r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        # under load = the upper half of the samples
        sm_sorted = sorted(sm)
        return {"sm_mhz": float(np.median(sm_sorted[len(sm_sorted) // 2:])),
                "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(args, steps=None):
    """The reference's CPU implementation of the path timed on the host cores: a bounded number
    of cycles of the same workload (by default the same 256^3 mesh as the GPU arm at N = 1)."""
    from artemis_b200 import pgen
    from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
    from artemis_b200.mesh import UniformMesh
    from artemis_b200.params import FluidParams
    from oracle import oracle_py, ref_py
    # oracle/_ref (the reference's own sources, prebuilt where /root/reference is mounted)
    # when present, else the restatement
    use_ref = os.path.exists(ref_py._LIB)
    n = args.cpu_sample
    mesh = UniformMesh(nx=(n, n, n), xmin=(-1, -1, -1), xmax=(1, 1, 1),
                       block_nx=(min(64, n),) * 3, nghost=4, bcs=(BoundaryFlag.outflow,) * 6)
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod.ppm, RSolver.hllc,
                     cfl=0.3, nspecies=1, dfloor=1e-10, gamma=1.4, siefloor=1e-10)
    sim = (ref_py.RefSim if use_ref else oracle_py.OracleSim)(mesh, gas=gp, integrator="rk2")
    sim.gas.prim[:] = pgen.blast(mesh, gp.gamma, d0=1.0, p0=1e-5, internal_energy=1.0, radius=0.1,
                                 samples=0)
    sim.initialize()
    sim.step()  # warm-up (page faults, OpenMP pool)
    ncyc = steps or args.cpu_cycles
    times = []
    for _ in range(ncyc):
        t0 = time.perf_counter()
        sim.step()
        times.append(time.perf_counter() - t0)
    tot = sum(times)
    zc = mesh.interior_zones * ncyc / tot
    what = ("the reference's own flux/update/source/C2P/P2C sources (oracle/_ref, OpenMP mock "
            "Parthenon) + restated ghost exchange" if use_ref else "the OpenMP oracle port")
    return {"value": zc, "unit": "zone-cycles/s", "cores": int(oracle_py.lib().ao_num_threads()),
            "kind": "reference" if use_ref else "port",
            "sample": f"{ncyc} rk2 cycles of the {n}^3 blast ({mesh.nb} blocks of "
                      f"{mesh.block_nx[0]}^3, PPM+HLLC fp64) with {what}"}, times


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if "WORLD_SIZE" in os.environ:
        # torchrun exports OMP_NUM_THREADS=1 to every rank; the CPU arm uses all host cores
        import ctypes
        try:
            ctypes.CDLL("libgomp.so.1").omp_set_num_threads(os.cpu_count() or 1)
        except OSError:
            pass
    base, times = cpu_baseline(args, steps=max(1, args.steps))
    ms = 1e3 * float(np.mean(times))
    line = {"impl": "reference", "metric": METRIC, "value": base["value"],
            "unit": "zone-cycles/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"3D Sedov blast PPM+HLLC rk2 fp64, {args.cpu_sample}^3 zones in 64^3 "
                                   f"MeshBlocks, one rk2 cycle per step on the host cores (GPU arm: "
                                   f"{args.tile}^3 per GPU, 64^3 MeshBlocks"
                                   + (" -- the same mesh)" if args.cpu_sample == args.tile else ")"),
                       "note": "the full Kokkos/Parthenon build needs cmake and is not "
                               "buildable under this round's rules; this times the reference's "
                               "own hot-path sources compiled against a mock Parthenon "
                               "(oracle/_ref) with all host threads, or the oracle port when "
                               "that library is absent (see cpu_baseline.kind)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": "zone-cycles/s", "h2d_bytes_per_step": 0,
                    "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    _emit(line)


_REAL_STDOUT = None


def _claim_stdout():
    """stdout must carry exactly ONE JSON line, but NCCL (its version banner at any NCCL_DEBUG
    level >= VERSION), torch and the CUDA runtime write to file descriptor 1 behind Python's
    back.  Point fd 1 at stderr for the whole run and keep a private duplicate of the real
    stdout for the result line."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.dup(1)
        os.dup2(2, 1)


def _emit(line: dict):
    sys.stdout.flush()
    os.write(_REAL_STDOUT if _REAL_STDOUT is not None else 1, (json.dumps(line) + "\n").encode())


def main_config3(args):
    """BASELINE.json config 3: gas + dust multifluid (inputs/drag/simple_drag.in state extended to
    3-D: gas rho 10, v (1,0,0); dust rho 0.01 at rest; seeded sin-mode perturbation),
    PLM+HLLE both fluids, periodic, rk2, 64^3 MeshBlocks, STRONG scaling: --mesh^3 zones in total
    split block-spatially over the GPUs.  Drag itself stays on the reference path."""
    import torch
    import torch.distributed as dist

    from artemis_b200 import pgen
    from artemis_b200.driver import ArtemisDriver
    from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
    from artemis_b200.mesh import UniformMesh
    from artemis_b200.meshdata import MeshData
    from artemis_b200.params import FluidParams

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    lay = rank_layout(world)
    cfg4 = args.config == 4
    M, B, S = args.mesh, (32 if cfg4 else args.block), (1 if cfg4 else args.dust_species)
    if cfg4 and args.mesh == 512:
        M = 256
    nbt = tuple(M // B // lay[d] for d in range(3))
    rl = (rank % lay[0], (rank // lay[0]) % lay[1], rank // (lay[0] * lay[1]))
    if cfg4:   # inputs/disk/disk_sph.in geometry: r in [0.4, 2.5], theta around the midplane
        Cc = Coordinates.spherical3D
        mesh = UniformMesh(nx=(M, M, M), xmin=(0.4, 1.0707963267948966, 0.0),
                           xmax=(2.5, 2.0707963267948966, 6.283185307179586), block_nx=(B, B, B),
                           # the deck's `ic` user condition on r and theta (AB200_BC_FIXED: resident
                           # ghost zones, single-rank topologies); outflow when the mesh is split
                           nghost=4, bcs=((BoundaryFlag.fixed if world == 1 else BoundaryFlag.outflow),) * 4
                           + (BoundaryFlag.periodic,) * 2,
                           coords=Cc, lattice_lo=tuple(rl[d] * nbt[d] for d in range(3)), lattice_n=nbt)
        recon, per = ReconstructionMethod.ppm, (False, False, True)
    else:
        Cc = Coordinates.cartesian
        mesh = UniformMesh(nx=(M, M, M), xmin=(0, 0, 0), xmax=(1, 1, 1), block_nx=(B, B, B), nghost=4,
                           bcs=(BoundaryFlag.periodic,) * 6,
                           lattice_lo=tuple(rl[d] * nbt[d] for d in range(3)), lattice_n=nbt)
        recon, per = ReconstructionMethod.plm, (True, True, True)
    gp = FluidParams(Fluid.gas, Cc, recon, RSolver.hlle, cfl=0.3, nspecies=1,
                     dfloor=1e-10, gamma=1.4, siefloor=1e-10)
    dp = FluidParams(Fluid.dust, Cc, recon, RSolver.hlle, cfl=0.3, nspecies=S, dfloor=1e-10)
    bcs = [int(v) for v in mesh.bcs]
    for d in range(3):
        if lay[d] > 1:
            if per[d]:          # periodic + several ranks: both faces belong to other ranks
                bcs[2 * d] = bcs[2 * d + 1] = 3
            else:
                if rl[d] > 0:
                    bcs[2 * d] = 3
                if rl[d] < lay[d] - 1:
                    bcs[2 * d + 1] = 3
    md = MeshData(mesh, gas=gp, dust=dp, device=local, materialize_fluxes=False, bcs=bcs)
    md.set_stage_path(args.path)
    if cfg4:   # Keplerian power-law disk (disk.hpp profile shape) with a seeded azimuthal mode
        prim = np.zeros(mesh.shape(6))
        dprim = np.zeros(mesh.shape(4))
        for b in range(mesh.nb):
            x1v, x2v, x3v = pgen.cell_centers(mesh, b)
            r = x1v[None, None, :] + 0 * x2v[None, :, None] + 0 * x3v[:, None, None]
            th = x2v[None, :, None] + 0 * r
            ph = x3v[:, None, None] + 0 * r
            R = r * np.sin(th)
            rho = R ** -1.5 * np.exp(-((th - np.pi / 2) / 0.2) ** 2) * (1 + 1e-2 * np.sin(3 * ph))
            prim[b, 0] = np.maximum(rho, 1e-6)
            om = 0.0 if args.no_sources else 1.0     # <rotating_frame> omega = 1 (disk_sph.in)
            prim[b, 3] = R ** -0.5 - om * R
            prim[b, 1] = 1e-3 * np.cos(2 * ph)
            prim[b, 5] = 0.05 ** 2 / gp.gm1 / R
            dprim[b, 0] = 0.01 * prim[b, 0]
            dprim[b, 3] = R ** -0.5 - om * R
    else:
        prim, dprim = pgen.perturbed_constant(mesh, 6, S, amp=1e-3, seed=1234)
    prim[:, 4] = gp.gm1 * prim[:, 0] * prim[:, 5]
    md.gas.prim.set(prim)
    md.dust.prim.set(dprim)
    del prim, dprim
    comm = native = None
    if world > 1:
        from artemis_b200.comm import HaloComm, NativeComm
        comm = HaloComm(md, lay, rl, rank, world, periodic=per)
        native = NativeComm(md, lay, rank, world, periodic=per)
    drv = ArtemisDriver(md, "rk2", mode="fused", comm=comm)
    drv.Initialize()
    md.set_time_state(drv.dt)
    md.call("ab200_set_ghost_cons_lazy", 1)
    big = float(np.finfo(np.float64).max)
    if cfg4 and not args.no_sources:
        # inputs/disk/disk_sph.in: <gravity/point> mass = 1, <rotating_frame> omega = 1 -- every
        # stage is split (passes + mass-flux tap -> PointMassGravity -> RotatingFrameImpl ->
        # SetAux/C2P/P2C); the frame velocity also enters FluxSource (fluid_fluxes.hpp:433-437)
        import ctypes as C
        from artemis_b200 import capi
        md.call("ab200_set_rotating_frame", 1.0)
        sd = capi.SourcesDesc()
        sd.point_mass, sd.pm = 1, capi.PointMassDesc(1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)
        sd.rotating_frame, sd.rf_omega = 1, 1.0
        md.call("ab200_configure_sources", C.byref(sd))
        # <gas/viscosity> type = alpha, alpha = 1e-3 (nu = alpha c_s^2 / Omega_K, r0 = 1, gm = 1)
        dd = capi.DiffusionDesc()
        dd.visc_type, dd.alpha, dd.r0, dd.omega0 = 2, 1e-3, 1.0, 1.0
        md.call("ab200_configure_diffusion", C.byref(dd))
        drv.block_dt = drv.EstimateTimestep()    # the viscous limit enters the first dt
        drv.dt = float(np.finfo(np.float64).max)
        drv.SetGlobalTimeStep()
        md.set_time_state(drv.dt)
    if not args.no_drag and not cfg4:   # <drag/dust> type = constant, one stopping time per species
        import ctypes as C
        from artemis_b200 import capi
        sd = capi.SourcesDesc()
        sd.drag, sd.ntau = 1, S
        for n in range(S):
            sd.tau[n] = 10.0 ** (n - 3)
        md.call("ab200_configure_sources", C.byref(sd))

    def one_step():
        md.call("ab200_run_cycles_mr" if world > 1 else "ab200_run_cycles", 1, 1, big)

    def sync_all():
        md.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        one_step()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = md.launch_count()
    sync_all()
    md.call("ab200_timer_begin")
    for _ in range(args.steps):
        one_step()
    ms = __import__("ctypes").c_float()
    md.call("ab200_timer_end", __import__("ctypes").byref(ms))
    sync_all()
    launches = md.launch_count() - l0
    clocks = sampler.finish() if sampler else None
    t_ms = float(ms.value)
    if world > 1:
        tt = torch.tensor([t_ms, float(launches)], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt[0:1], op=dist.ReduceOp.MAX)
        dist.all_reduce(tt[1:2], op=dist.ReduceOp.SUM)
        t_ms, launches = float(tt[0].item()), int(tt[1].item())
    zones = M ** 3
    ms_step = t_ms / args.steps
    value = zones * args.steps / (t_ms * 1e-3)
    peak, peak_src = measured_peaks()
    alg = (240.0 + 160.0 * S) * 2          # B per zone-cycle, rk2
    if rank == 0:
        _emit({"metric": METRIC, "value": value, "unit": "zone-cycles/s", "n_gpus": world,
               "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_step,
               "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64",
               "data": "synthetic",
               "config": {"workload": (f"config 4: spherical 3-D Keplerian disk, gas + 1 dust species, "
                                       f"PPM+HLLE (WENO5 absent upstream), rk2, "
                                       f"{'ic user BCs (AB200_BC_FIXED)' if world == 1 else 'outflow'} in r/theta + periodic "
                                       f"phi, {M}^3 zones in TOTAL in {B}^3 MeshBlocks over {world} GPU(s); "
                                       "curvilinear fluxes, PLM_G-free PPM, geometric source terms; "
                                       + ("no source terms" if args.no_sources else
                                          "point-mass gravity (gm 1) + rotating frame (omega 1, mass-flux "
                                          "tap) + alpha viscosity (1e-3) every stage as in disk_sph.in, "
                                          "split stage")
                                       if cfg4 else
                                       f"config 3: gas + {S} dust species (inputs/drag state + seeded "
                                       f"perturbation), PLM+HLLE, rk2, periodic, {M}^3 zones in TOTAL in "
                                       f"{B}^3 MeshBlocks split over {world} GPU(s) (strong scaling); "
                                       + ("Drag::DragSource stays on the reference path" if args.no_drag
                                          else "implicit gas-dust drag (constant stopping times) every stage, "
                                               "split stage: passes -> drag -> SetAux/C2P/P2C")),
                          "zones_total": zones, "ranks": list(lay), "stage_path": md.stage_path(),
                          "l2": "state >> 126 MB L2, no flush needed"},
               "roofline": {"bound": "hbm", "achieved": value * alg / world / 1e9, "peak": peak,
                            "unit": "GB/s", "frac": value * alg / world / 1e9 / peak, "traffic": None,
                            "peak_source": peak_src,
                            "algorithmic_bytes_per_zone_cycle": alg,
                            "note": "whole cycle per GPU (stages + ghost fills + exchange)"},
               "cpu_baseline": None, "e2e": None, "gpu_launches": launches, "clocks": clocks})
    md.call("ab200_set_ghost_cons_lazy", 0)
    md.close()
    if world > 1:
        dist.destroy_process_group()


def main_config5(args):
    """config 5: 3-D shearing sheet with static mesh refinement on ONE GPU -- the root lattice
    of 4^3 MeshBlocks with its central 2^3 blocks refined (the core has the resolution of a
    uniform mesh twice as fine per direction), gas PLM+HLLC, rk2, shearing box (Omega 1, q 1.5)
    + point-mass gravity as in inputs/ssheet/ssheet.in.  Every stage runs the reference's task
    list: CalculateFluxes -> flux correction on the fine-coarse faces -> ApplyUpdate ->
    FluxSource -> sources -> SetAux / C2P -> multilevel ghost exchange (restrict, copies,
    coarse BCs, prolongate, fine BCs) -> P2C.  `--no-flux-correction` runs the fused stage
    kernels instead (the non-conservative variant)."""
    import ctypes as C

    import torch

    from artemis_b200.driver import ArtemisDriver
    from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
    from artemis_b200.meshdata import MeshData
    from artemis_b200.multilevel import MultilevelExchange, MultilevelMesh
    from artemis_b200.params import FluidParams

    if int(os.environ.get("WORLD_SIZE", "1")) != 1 or args.gpus != 1:
        raise SystemExit("bench.py --config 5: refined meshes are single-GPU (DESIGN.md 6)")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(0)
    B = args.block
    O, P = BoundaryFlag.outflow, BoundaryFlag.periodic
    mesh = MultilevelMesh(root_blocks=(4, 4, 4), block_nx=(B, B, B), xmin=(-1.0, -1.0, -1.0),
                          xmax=(1.0, 1.0, 1.0),
                          refine=tuple((i, j, k) for k in (1, 2) for j in (1, 2) for i in (1, 2)),
                          nghost=4,
                          bcs=((BoundaryFlag.extrap, BoundaryFlag.extrap, BoundaryFlag.inflow,
                                BoundaryFlag.inflow, BoundaryFlag.extrap, BoundaryFlag.extrap)
                               if args.bcs == "deck" else (O, O, P, P, O, O)))
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod.plm, RSolver.hllc,
                     cfl=0.3, nspecies=1, dfloor=1e-10, gamma=1.000001, siefloor=1e-10)
    fused = args.no_flux_correction
    # StratParams of the deck's `inflow` faces: <rotating_frame> qshear 1.5, omega 1
    md = MeshData(mesh, gas=gp, device=0, materialize_fluxes=not fused, shear_bc=(1.5, 1.0))
    rng = np.random.default_rng(1234)
    prim = np.zeros(mesh.shape(6))
    for b in range(mesh.nb):
        x = [mesh.blk_lo[b, d] + (np.arange((mesh.ni, mesh.nj, mesh.nk)[d]) - mesh.ngd[d] + 0.5)
             * mesh.blk_dx[b, d] for d in range(3)]
        X, Y, Z = x[0][None, None, :], x[1][None, :, None], x[2][:, None, None]
        prim[b, 0] = np.exp(-0.5 * (Z / 0.4) ** 2) * (1 + 0.05 * np.sin(3 * np.pi * X) * np.cos(2 * np.pi * Y))
        prim[b, 1] = 1e-3 * np.sin(2 * np.pi * Y) + 0 * X + 0 * Z
        prim[b, 2] = -1.5 * X + 0 * Y + 0 * Z          # background shear, q Omega x
        prim[b, 3] = 1e-3 * rng.standard_normal(1)[0] * np.cos(np.pi * X) + 0 * Y + 0 * Z
        prim[b, 5] = 0.05 ** 2 / gp.gm1                 # isothermal-like, c_s = h Omega = 0.05
    prim[:, 4] = gp.gm1 * prim[:, 0] * prim[:, 5]
    md.gas.prim.set(prim)
    del prim
    ex = MultilevelExchange(md)
    # <gravity/point> mass = 1e-5, soft = 0.03 at the origin; <rotating_frame> omega 1, qshear 1.5
    drv = ArtemisDriver(md, "rk2", mode="fused" if fused else "tasks", comm=ex,
                        sources=[("point_mass", 1.0e-5, 0.0, 0.0, 0.0, 0.03, 0.0, 0.0),
                                 ("shearing_box", 1.0, 1.5)],
                        flux_correction=not fused)
    drv.Initialize()

    def sync_all():
        md.synchronize()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        drv.Step()
    sync_all()
    sampler = ClockSampler(0)
    sampler.start()
    time.sleep(0.3)
    l0 = md.launch_count()
    # the cycle is host-driven (dt comes back every cycle, like the reference's driver): wall
    # clock between synchronisations is the honest number; device time is reported beside it
    md.call("ab200_timer_begin")
    t0 = time.perf_counter()
    for _ in range(args.steps):
        drv.Step()
    ms = C.c_float()
    md.call("ab200_timer_end", C.byref(ms))
    sync_all()
    wall_ms = (time.perf_counter() - t0) * 1e3
    launches = md.launch_count() - l0
    clocks = sampler.finish()
    # the multilevel exchange on its own (device time of 10 calls)
    md.call("ab200_timer_begin")
    for _ in range(10):
        ex.exchange()
    ems = C.c_float()
    md.call("ab200_timer_end", C.byref(ems))
    fcms = C.c_float(0.0)
    if not fused:
        md.call("ab200_timer_begin")
        for _ in range(10):
            ex.flux_correct()
        md.call("ab200_timer_end", C.byref(fcms))
    sync_all()
    zones = mesh.interior_zones
    t_ms = max(float(ms.value), wall_ms)
    value = zones * args.steps / (t_ms * 1e-3)
    peak, peak_src = measured_peaks()
    alg = 240.0 * 2
    _emit({"metric": METRIC, "value": value, "unit": "zone-cycles/s", "n_gpus": 1,
           "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / args.steps,
           "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
           "data": "synthetic",
           "config": {"workload": (f"config 5: 3-D shearing sheet with static mesh refinement, gas PLM+HLLC, "
                                   f"rk2, gamma 1.000001, shearing box (Omega 1, q 1.5) + point-mass "
                                   f"gravity every stage; root lattice 4^3 MeshBlocks of {B}^3 with the "
                                   f"central 2^3 refined: {mesh.nb} MeshBlocks, {zones} zones, the core "
                                   f"at the resolution of a uniform {8 * B}^3 mesh; " +
                                   ("the deck's own user BCs: extrap x1 / x3, inflow x2 (strat.hpp) on "
                                    "fine arrays and coarse buffers" if args.bcs == "deck" else
                                    "outflow x1/x3, periodic x2 (--bcs deck runs the deck's extrap / "
                                    "inflow user conditions)")),
                      "zones_total": zones, "blocks": mesh.nb,
                      "path": ("fused stage kernels, NO flux correction (non-conservative variant)"
                               if fused else
                               "reference task list per stage incl. flux correction on every "
                               "fine-coarse face (AddFluxCorrectionTasks)"),
                      "multilevel_exchange_ms": float(ems.value) / 10,
                      "flux_correction_ms": float(fcms.value) / 10,
                      "exchange_descriptors": ex.n,
                      "device_ms_per_step": float(ms.value) / args.steps,
                      "wall_ms_per_step": wall_ms / args.steps,
                      "l2": "state >> 126 MB L2, no flush needed"},
           "roofline": {"bound": "hbm", "achieved": value * alg / 1e9, "peak": peak, "unit": "GB/s",
                        "frac": value * alg / 1e9 / peak, "traffic": None, "peak_source": peak_src,
                        "algorithmic_bytes_per_zone_cycle": alg,
                        "note": "whole cycle (task kernels materialise fluxes: the algorithmic "
                                "bytes are those of the fused stage, SURVEY 8d)"},
           "cpu_baseline": None, "e2e": None, "gpu_launches": launches, "clocks": clocks})
    ex.close()
    md.close()


def main():
    args = parse()
    _claim_stdout()
    if args.impl == "reference":
        run_reference(args)
        return
    if args.config in (3, 4):
        main_config3(args)
        return
    if args.config == 5:
        main_config5(args)
        return
    import torch
    import torch.distributed as dist

    from artemis_b200 import pgen
    from artemis_b200.driver import ArtemisDriver
    from artemis_b200.meshdata import MeshData

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    assert world == args.gpus, f"--gpus {args.gpus} but WORLD_SIZE={world}"
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local)
    comm = None
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    mesh, gp, lay, rl = make_problem(args, rank, world)
    bcs = list(mesh.bcs)
    if world > 1:
        from artemis_b200.comm import HaloComm
        for d in range(3):
            if rl[d] > 0:
                bcs[2 * d] = 3            # AB200_BC_NONE: neighbour rank
            if rl[d] < lay[d] - 1:
                bcs[2 * d + 1] = 3
    md = MeshData(mesh, gas=gp, device=local, materialize_fluxes=False, bcs=bcs)
    md.set_stage_path(args.path)
    path = md.stage_path()
    native = None
    if world > 1:
        comm = HaloComm(md, lay, rl, rank, world)   # Initialize() + the exchange-only probe
        if args.transport == "native":
            from artemis_b200.comm import NativeComm
            native = NativeComm(md, lay, rank, world)
    if args.state == "blast":
        prim = pgen.blast(mesh, gp.gamma, d0=1.0, p0=1e-5, internal_energy=1.0, radius=0.1, samples=0)
    else:
        from tests.helpers import random_prim
        prim = random_prim(mesh, gp, seed=1234)
    md.gas.prim.set(prim)
    drv = ArtemisDriver(md, "rk2", mode="fused", comm=comm)
    drv.Initialize()
    zones = mesh.interior_zones * world
    integ = 1  # rk2

    def sync_all():
        md.synchronize()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    def one_step():
        if world == 1:
            md.call("ab200_run_cycles", integ, 1, float(np.finfo(np.float64).max))
        elif native is not None:   # the whole multi-rank cycle behind the C ABI
            md.call("ab200_run_cycles_mr", integ, 1, float(np.finfo(np.float64).max))
        else:
            drv.StepDevice()

    md.set_time_state(drv.dt)
    # conserved ghost zones are never read by the fused stages: the ghost fills write primitives
    # only and ab200_sync_ghost_cons converts them once, before anything observes the arrays
    md.call("ab200_set_ghost_cons_lazy", 1)
    for _ in range(args.warmup):
        one_step()
    sync_all()
    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
        time.sleep(0.3)
    l0 = md.launch_count()
    sync_all()
    md.call("ab200_timer_begin")
    t_cpu0 = time.perf_counter()
    for _ in range(args.steps):
        one_step()
    cpu_issue_ms = 1e3 * (time.perf_counter() - t_cpu0) / args.steps   # host time to ISSUE a step
    ms = __import__("ctypes").c_float()
    md.call("ab200_timer_end", __import__("ctypes").byref(ms))
    sync_all()
    launches = md.launch_count() - l0
    md.call("ab200_set_ghost_cons_lazy", 0)
    md.call("ab200_sync_ghost_cons")
    clocks = sampler.finish() if sampler else None
    t_ms = float(ms.value)
    if world > 1:
        tt = torch.tensor([t_ms], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_ms = float(tt.item())
        lt = torch.tensor([launches], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(lt, op=dist.ReduceOp.SUM)
        launches = int(lt.item())
    value = zones * args.steps / (t_ms * 1e-3)
    # N > 1: device time of the remote halo exchange alone (single round: one pack launch, one
    # NCCL group of sends/receives to all peers, one unpack launch), back to back with no compute in between -- explains the scaling loss
    comm_ms = None
    if world > 1:
        sync_all()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        for _ in range(3):
            comm.exchange_direct(md)
        sync_all()
        ev0.record()
        for _ in range(20):
            comm.exchange_direct(md)
        ev1.record()
        torch.cuda.synchronize()
        ct = torch.tensor([ev0.elapsed_time(ev1) / 20.0], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(ct, op=dist.ReduceOp.MAX)
        comm_ms = float(ct.item())
    ts = md.time_state() if world == 1 else None

    # ---- per-kernel timing of the dominant kernels (fused directional passes) -------------
    import ctypes as C
    peak, peak_src = measured_peaks()
    kms = []
    reps = 3
    for stage, (g0, g1, b) in enumerate(((0.0, 1.0, 1.0), (0.5, 0.5, 0.5))):
        acc = 0.0
        for _ in range(reps):
            md.synchronize()
            md.call("ab200_timer_begin")
            # dt=0: state kept; flag 4 = AB200_STAGE_PINGPONG (the stage kernel alone, no copy back)
            md.call("ab200_fused_stage", g0, g1, b, 0.0, 0, int(stage == 0), 4)
            k = C.c_float()
            md.call("ab200_timer_end", C.byref(k))
            acc += k.value
        kms.append(acc / reps)
    md.call("ab200_sync_prim")
    stage_ms = float(np.mean(kms))
    zones_local = mesh.interior_zones
    achieved = ALG_BYTES_PER_ZONE_STAGE * zones_local / (stage_ms * 1e-3) / 1e9
    traffic = None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    counters = None
    if os.path.exists(tp):
        with open(tp) as fh:
            tj = json.load(fh)
        traffic = tj.get({"single_pass": "single_pass_dram_bytes_per_launch",
                          "role_split": "role_split_dram_bytes_per_launch"}.get(
                              path, "fused_stage_dram_bytes_per_launch"))
        # ncu counters of the same kernels (committed evidence, not re-measured here): the fp64
        # pipe and issue-slot utilisation that explain why the HBM fraction is what it is
        cj = tj.get("counters") or {}
        counters = cj.get(path if path in ("single_pass", "role_split") else "three_pass")
        if counters is not None:
            counters = dict(counters, source=cj.get("source"), secondary_bound=cj.get("secondary_bound"))
    roofline = {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic,
                "kernel": (("k_trio_stage" if path == "role_split" else "k_sweep_stage") +
                           " (one launch = one full stage: x1+x2+x3 reconstruct/Riemann/"
                           "update + C2P, primitives and conserved state cross HBM once)"
                           if path in ("single_pass", "role_split") else
                           "k_xchunk_pass + k_march_pass<2> + k_march_pass<3> (one fused stage = the "
                           "three directional passes; 'achieved' = algorithmic bytes of the stage / "
                           "their summed duration)"),
                "stage_ms": kms, "peak_source": peak_src, "ncu_counters": counters,
                "algorithmic_bytes_per_launch": ALG_BYTES_PER_ZONE_STAGE * zones_local,
                "whole_cycle_frac": value / world * 2 * ALG_BYTES_PER_ZONE_STAGE / (peak * 1e9)}

    # ---- end to end through the host-buffer C-ABI entry point ------------------------------
    e2e = None
    if not args.no_e2e and world == 1:
        nv = gp.nvar
        shape = mesh.shape(nv)
        hp = torch.empty(shape, dtype=torch.float64).pin_memory()
        hc = torch.empty(shape, dtype=torch.float64).pin_memory()
        hp.numpy()[:] = md.gas.prim.get()
        dt_io = C.c_double(drv.dt)
        DP = C.POINTER(C.c_double)
        php, phc = C.cast(hp.data_ptr(), DP), C.cast(hc.data_ptr(), DP)
        # the step's result is the complete new primitive state; conserved arrays are a pure
        # function of it and are not requested (cons pointer NULL); the input's pressure entries
        # are never uploaded (recomputed by PrimToCons)
        del phc, hc
        xfer = {"full": 0, "interior": 1 | 2 | 4, "interior_zc": 1 | 2 | 4 | 8,
                "interior_dma": 1 | 2}[args.e2e_transfer]
        md.call("ab200_set_host_transfer", xfer)
        md.call("ab200_cycles_host", integ, 1, C.byref(dt_io), php, None, None, None)  # warm-up
        md.synchronize()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            md.call("ab200_cycles_host", integ, 1, C.byref(dt_io), php, None, None, None)
        md.synchronize()
        te = (time.perf_counter() - t0) / args.e2e_steps
        md.call("ab200_set_host_transfer", 0)
        # bytes that crossed PCIe: whole arrays (ghost zones included) or interior zones only
        nbytes = (int(np.prod(shape)) if xfer == 0 else mesh.interior_zones * nv) * 8
        e2e = {"value": mesh.interior_zones / te, "unit": "zone-cycles/s",
               "h2d_bytes_per_step": nbytes * 5 // 6, "d2h_bytes_per_step": nbytes,
               "ms_per_step": te * 1e3, "transfer": args.e2e_transfer,
               "api": "ab200_cycles_host (pinned host primitives in, pressure not uploaded; "
                      "primitives out, cons not requested)" +
                      ("" if xfer == 0 else "; ab200_set_host_transfer: interior zones only, ghost "
                       "zones rebuilt on the device" +
                       {7: " (in: copy kernel on the pinned array, out: strided DMA)",
                        15: " (copy kernels on the pinned arrays)", 3: " (strided DMA)"}[xfer])}
    elif not args.no_e2e:
        # N > 1: the same end-to-end step through the public entry points every rank calls --
        # pinned host primitives in, ab200_prim_to_cons, one device-resident cycle with the
        # remote halo exchange and the dt all-reduce, the new primitives back to pinned host
        # memory (as in the N = 1 leg the conserved arrays are not requested: they are a pure
        # function of the primitives); wall clock between barriers, max over ranks
        nv = gp.nvar
        shape = mesh.shape(nv)
        nbytes = int(np.prod(shape)) * 8
        hp = torch.empty(shape, dtype=torch.float64).pin_memory()
        hp.numpy()[:] = md.gas.prim.get()

        xfer = {"full": 0, "interior": 1 | 2 | 4, "interior_zc": 1 | 2 | 4 | 8,
                "interior_dma": 1 | 2}[args.e2e_transfer]
        use_host_entry = native is not None     # ab200_cycles_host is collective over the ranks
        php = C.cast(hp.data_ptr(), C.POINTER(C.c_double))
        dt_io = C.c_double(drv.dt)

        def e2e_step():
            if use_host_entry:
                md.call("ab200_cycles_host", integ, 1, C.byref(dt_io), php, None, None, None)
                return
            capi_check(md.L.ab200_memcpy_h2d(md.ctx, md.gas.prim.ptr, hp.data_ptr(), nbytes))
            md.call("ab200_prim_to_cons")
            drv.StepDevice()
            capi_check(md.L.ab200_memcpy_d2h(md.ctx, hp.data_ptr(), md.gas.prim.ptr, nbytes))

        def capi_check(rc):
            if rc != 0:
                raise SystemExit(f"bench.py: host<->device copy failed ({rc})")

        if use_host_entry:
            md.call("ab200_set_host_transfer", xfer)
        e2e_step()  # warm-up
        sync_all()
        t0 = time.perf_counter()
        for _ in range(args.e2e_steps):
            e2e_step()
        sync_all()
        te = (time.perf_counter() - t0) / args.e2e_steps
        if use_host_entry:
            md.call("ab200_set_host_transfer", 0)
            if xfer:
                nbytes = mesh.interior_zones * nv * 8
        tt = torch.tensor([te], device=f"cuda:{local}", dtype=torch.float64)
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        te = float(tt.item())
        e2e = {"value": zones / te, "unit": "zone-cycles/s",
               "h2d_bytes_per_step": (nbytes * 5 // 6 if use_host_entry else nbytes) * world,
               "d2h_bytes_per_step": nbytes * world, "ms_per_step": te * 1e3,
               "transfer": args.e2e_transfer if use_host_entry else "full",
               "api": ("per rank, collectively: ab200_cycles_host (pinned host primitives in, "
                       "pressure not uploaded; ghost zones rebuilt on the device incl. the remote "
                       "round; ab200_run_cycles_mr; primitives out)" if use_host_entry else
                       "per rank: ab200_memcpy_h2d (pinned prim) -> ab200_prim_to_cons -> "
                       "device-resident cycle (remote halo exchange, dt all-reduce) -> "
                       "ab200_memcpy_d2h (prim)")}
    else:
        e2e = None

    base = None
    if rank == 0 and world == 1 and not args.no_cpu:
        base, _ = cpu_baseline(args)

    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": "zone-cycles/s", "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": t_ms / args.steps,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
                "data": "synthetic",
                "config": {"workload": f"3D Sedov blast (inputs/blast + 3D overrides), PPM+HLLC, "
                                       f"rk2, {args.tile}^3 zones per GPU in {args.block}^3 "
                                       f"MeshBlocks, nghost=4, outflow",
                           "zones_total": zones, "ranks": list(lay),
                           "l2": "state 3.4 GB/GPU >> 126 MB L2, no flush needed",
                           "path": ("single-pass stage kernel" if path in ("single_pass", "role_split") else
                                    "three directional fused passes") +
                                   (" + fused ghost fill + device-resident dt" if world == 1 else
                                    " + single-round NCCL halo exchange + device-resident dt all-reduce"),
                           "state": args.state, "transport": args.transport if world > 1 else None,
                           "peer_write_ipc": (bool(md.L.ab200_comm_is_direct(md.ctx)) if native is not None else None),
                           "overlap": bool(os.environ.get("AB200_OVERLAP")) if world > 1 else None,
                           "stage_path": path, "halo_exchange_ms": comm_ms, "host_issue_ms_per_step": cpu_issue_ms,
                           "halo_bytes_per_exchange": (comm.bytes_per_direct_exchange if comm else 0)},
                "roofline": roofline, "cpu_baseline": base, "e2e": e2e,
                "gpu_launches": launches, "clocks": clocks}
        if ts is not None:
            line["config"]["sim_time"] = float(ts[2])
        _emit(line)
    if world > 1:
        dist.destroy_process_group()
    md.close()


if __name__ == "__main__":
    main()
