"""Stage time of the two ab200_fused_stage paths (three directional passes vs the single-pass
kernel) per fluid / reconstruction / Riemann solver, 256^3 zones in 64^3 MeshBlocks on one GPU.
Feeds the AB200_PATH_AUTO policy (csrc/sweep_host.cu) and DESIGN.md section 3.1.

  python scripts/stage_matrix.py [--tile 256] [--out gpurun_out/stage_matrix.json]
"""
import argparse
import ctypes as C
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver
from artemis_b200.mesh import UniformMesh
from artemis_b200.meshdata import MeshData
from artemis_b200.params import FluidParams
from artemis_b200 import pgen

ap = argparse.ArgumentParser()
ap.add_argument("--tile", type=int, default=256)
ap.add_argument("--block", type=int, default=64)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--out", default="gpurun_out/stage_matrix.json")
args = ap.parse_args()

Cc = Coordinates.cartesian
T, B = args.tile, args.block
mesh = UniformMesh(nx=(T, T, T), xmin=(-1, -1, -1), xmax=(1, 1, 1), block_nx=(B, B, B), nghost=4,
                   bcs=(BoundaryFlag.outflow,) * 6)
cases = [("gas", "ppm", "hllc"), ("gas", "ppm", "hlle"), ("gas", "plm", "hllc"),
         ("gas", "plm", "hlle"), ("gas", "pcm", "llf"), ("dust", "ppm", "hlle"),
         ("dust", "plm", "hlle"), ("dust", "plm", "llf")]
rows = []
for fluid, rc, rs in cases:
    if fluid == "gas":
        fp = FluidParams(Fluid.gas, Cc, ReconstructionMethod[rc], RSolver[rs], cfl=0.3, nspecies=1,
                         dfloor=1e-10, gamma=1.4, siefloor=1e-10)
        md = MeshData(mesh, gas=fp, materialize_fluxes=False)
        ff = md.gas
        prim = pgen.blast(mesh, fp.gamma, d0=1.0, p0=1e-5, internal_energy=1.0, radius=0.1, samples=0)
        bytes_zone = 240.0
    else:
        fp = FluidParams(Fluid.dust, Cc, ReconstructionMethod[rc], RSolver[rs], cfl=0.3, nspecies=1,
                         dfloor=1e-10)
        md = MeshData(mesh, dust=fp, materialize_fluxes=False)
        ff = md.dust
        rng = np.random.default_rng(7)
        prim = np.zeros(mesh.shape(fp.nvar))
        prim[:, 0] = 1.0 + 0.1 * rng.random(prim[:, 0].shape)
        prim[:, 1:4] = 0.2 * (rng.random(prim[:, 1:4].shape) - 0.5)
        bytes_zone = 160.0
    ff.prim.set(prim)
    md.call("ab200_prim_to_cons")
    md.call("ab200_fill_ghosts")
    res = {"fluid": fluid, "recon": rc, "riemann": rs}
    for name, path in (("three_pass", 1), ("single_pass", 2)):
        md.set_stage_path(path)
        ms = []
        for stage, (g0, g1, b) in enumerate(((0.0, 1.0, 1.0), (0.5, 0.5, 0.5))):
            acc = 0.0
            for r in range(args.reps + 1):
                md.synchronize()
                md.call("ab200_timer_begin")
                md.call("ab200_fused_stage", g0, g1, b, 0.0, 0, int(stage == 0), 4)  # dt = 0: state kept
                k = C.c_float()
                md.call("ab200_timer_end", C.byref(k))
                if r > 0:
                    acc += k.value
            ms.append(acc / args.reps)
        md.call("ab200_sync_prim")
        res[name + "_ms"] = ms
        res[name + "_GBs"] = bytes_zone * mesh.interior_zones / (np.mean(ms) * 1e-3) / 1e9
    res["faster"] = "single_pass" if np.mean(res["single_pass_ms"]) < np.mean(res["three_pass_ms"]) else "three_pass"
    rows.append(res)
    print("%-4s %-3s %-4s  three-pass %6.3f %6.3f ms   single-pass %6.3f %6.3f ms   -> %s" % (
        fluid, rc, rs, *res["three_pass_ms"], *res["single_pass_ms"], res["faster"]), flush=True)
    md.close()
os.makedirs(os.path.dirname(args.out) or ".", exist_ok=True)
with open(args.out, "w") as fh:
    json.dump({"tile": T, "block": B, "rows": rows}, fh, indent=1)
