"""A/B of the host-transfer modes of ab200_cycles_host on the bench mesh (256^3 in 64^3 blocks):
whole arrays by DMA, interior zones by strided DMA, interior zones by copy kernels on the
pinned arrays (8- and 16-byte accesses), and each direction on its own."""
import ctypes as C
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from artemis_b200 import pgen  # noqa: E402
from artemis_b200.driver import ArtemisDriver  # noqa: E402
from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid, ReconstructionMethod, RSolver  # noqa: E402
from artemis_b200.mesh import UniformMesh  # noqa: E402
from artemis_b200.meshdata import MeshData  # noqa: E402
from artemis_b200.params import FluidParams  # noqa: E402

n, blk = 256, 64
mesh = UniformMesh(nx=(n,) * 3, xmin=(-1,) * 3, xmax=(1,) * 3, block_nx=(blk,) * 3, nghost=4,
                   bcs=(BoundaryFlag.outflow,) * 6)
gp = FluidParams(Fluid.gas, Coordinates.cartesian, ReconstructionMethod.ppm, RSolver.hllc, cfl=0.3,
                 nspecies=1, dfloor=1e-10, gamma=1.4, siefloor=1e-10)
md = MeshData(mesh, gas=gp, device=0, materialize_fluxes=False)
md.gas.prim.set(pgen.blast(mesh, gp.gamma, d0=1.0, p0=1e-5, internal_energy=1.0, radius=0.1, samples=0))
drv = ArtemisDriver(md, "rk2", mode="fused")
drv.Initialize()
hp = torch.empty(mesh.shape(gp.nvar), dtype=torch.float64).pin_memory()
hp.numpy()[:] = md.gas.prim.get()
php = C.cast(hp.data_ptr(), C.POINTER(C.c_double))
out = {}
for name, flags, env in [("full_dma", 0, None), ("in_zc16", 1 | 4, None), ("out_zc16", 2 | 8, None),
                         ("both_zc16", 15, None), ("both_zc8", 15, "1"), ("in_dma3d", 1, None),
                         ("out_dma3d", 2, None), ("in_zc16_out_dma3d", 7, None)]:
    if env:
        os.environ["AB200_XFER_SCALAR"] = env
    else:
        os.environ.pop("AB200_XFER_SCALAR", None)
    md.call("ab200_set_host_transfer", flags)
    dt_io = C.c_double(drv.dt)
    md.call("ab200_cycles_host", 1, 1, C.byref(dt_io), php, None, None, None)
    md.synchronize()
    t0 = time.perf_counter()
    for _ in range(3):
        md.call("ab200_cycles_host", 1, 1, C.byref(dt_io), php, None, None, None)
    md.synchronize()
    out[name] = (time.perf_counter() - t0) / 3 * 1e3
    print(name, f"{out[name]:.2f} ms", flush=True)
print(json.dumps(out))
md.close()
