"""Where does the sweep path differ from the oracle under reflecting boundaries?"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from artemis_b200.driver import ArtemisDriver
from artemis_b200.enums import BoundaryFlag, Coordinates, Fluid
from artemis_b200.mesh import UniformMesh
from artemis_b200.meshdata import MeshData
from oracle.oracle_py import OracleSim
from tests.helpers import dust_params, gas_params, random_prim

C = Coordinates.cartesian
bc = sys.argv[1] if len(sys.argv) > 1 else "reflect"
with_dust = (sys.argv[2] == "dust") if len(sys.argv) > 2 else True
variant = sys.argv[3] if len(sys.argv) > 3 else "fast"
path = sys.argv[4] if len(sys.argv) > 4 else "auto"
integ = sys.argv[5] if len(sys.argv) > 5 else "rk2"
B = BoundaryFlag
bcs = {"periodic": (B.periodic,) * 6, "outflow": (B.outflow,) * 6, "reflect": (B.reflect,) * 6}[bc]
bnx = (16, 16, 16)
mesh = UniformMesh(nx=tuple(2 * b for b in bnx), xmin=(0, 0, 0), xmax=(1.0, 0.8, 0.6), block_nx=bnx,
                   nghost=4, bcs=bcs, coords=C)
gp = gas_params(C, "ppm", "hllc")
dp = dust_params(C, "plm", "hlle", S=2) if with_dust else None
osim = OracleSim(mesh, gas=gp, dust=dp, integrator=integ)
md = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
md.set_stage_path(path)
for which, fp in ((Fluid.gas, gp), (Fluid.dust, dp)):
    if fp is None:
        continue
    p = random_prim(mesh, fp, seed=5 + int(which), shocks=False)
    (osim.gas if which == Fluid.gas else osim.dust).prim[:] = p
    md.fluid(which).prim.set(p)
osim.nlim = 1
osim.initialize(); osim.run()
drv = ArtemisDriver(md, integ, mode="fused", nlim=1)
drv.Initialize(); drv.Execute()
print("bc", bc, variant, path, integ, md.stage_path(), "dust", with_dust, "NO_SWEEP", os.environ.get("AB200_NO_SWEEP"), "dt", drv.dt, osim.dt)
sl = mesh.interior()
for of, df in zip(osim.fluids, md.fluids):
    for name, a, b in (("u0", df.u0.get(), of.u0), ("prim", df.prim.get(), of.prim)):
        err = np.abs(a - b) / np.maximum(np.maximum(np.abs(a), np.abs(b)), 1e-3 * np.abs(b).max())
        ei = err[(slice(None), slice(None)) + sl]
        print(name, "max err all %.3e interior %.3e" % (err.max(), ei.max()))
        idx = np.unravel_index(np.argmax(err), err.shape)
        print("   worst at block,var,k,j,i =", idx, "got", a[idx], "want", b[idx])
        for v in range(a.shape[1]):
            e = err[:, v]
            n_bad = int((e > 0).sum())
            if n_bad:
                w = np.argwhere(e > 0)
                print("   var", v, "bad cells", n_bad, "k range", w[:, 1].min(), w[:, 1].max(), "j range",
                      w[:, 2].min(), w[:, 2].max(), "i range", w[:, 3].min(), w[:, 3].max(), "blocks", sorted(set(w[:, 0])))
md.close()
