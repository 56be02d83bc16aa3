"""CUDA-graph replay of ab200_run_cycles against the eager loop on launch-bound meshes: config 1
(2-D linear wave deck, 128 x 64 in 32^2 MeshBlocks, PLM+HLLC, nghost 2) and a small 3-D mesh."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from artemis_b200 import pgen  # noqa: E402
from artemis_b200.driver import ArtemisDriver  # noqa: E402
from artemis_b200.enums import Coordinates, Fluid, ReconstructionMethod, RSolver  # noqa: E402
from artemis_b200.mesh import UniformMesh  # noqa: E402
from artemis_b200.meshdata import MeshData  # noqa: E402
from artemis_b200.params import FluidParams  # noqa: E402

out = {}
for name, nx, bnx, ng, recon in [("config1_2d_128x64", (128, 64, 1), (32, 32, 1), 2, "plm"),
                                 ("3d_64cubed_in_32cubed_blocks", (64, 64, 64), (32, 32, 32), 4, "ppm")]:
    mesh = UniformMesh(nx=nx, xmin=(0, 0, 0), xmax=(3.0, 1.5, 1.5), block_nx=bnx, nghost=ng)
    gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod[recon], RSolver.hllc,
                     cfl=0.9, nspecies=1, dfloor=1e-20, gamma=1.66666666667)
    prim, lw = pgen.linear_wave(mesh, gp.gamma, 3, 1e-3, 1.0)
    res = {}
    for graph in (0, 1):
        md = MeshData(mesh, gas=gp, device=0, materialize_fluxes=False)
        md.gas.prim.set(prim)
        drv = ArtemisDriver(md, "rk2", mode="fused")
        drv.Initialize()
        md.set_time_state(drv.dt)
        md.call("ab200_set_graph_replay", graph)
        big = float(np.finfo(np.float64).max)
        md.call("ab200_run_cycles", 1, 50, big)
        md.synchronize()
        t0 = time.perf_counter()
        md.call("ab200_run_cycles", 1, 1000, big)
        md.synchronize()
        res["graph" if graph else "eager"] = (time.perf_counter() - t0) / 1000 * 1e6
        md.close()
    out[name] = {"us_per_cycle": res, "speedup": res["eager"] / res["graph"],
                 "zones": int(np.prod(nx))}
    print(name, out[name], flush=True)
print(json.dumps(out))
