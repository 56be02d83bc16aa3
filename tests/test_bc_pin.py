"""Physical boundary conditions: the oracle's outflow / reflecting
ghost fill (oracle/artemis_oracle.c, ao_exchange_ghosts_phase, phase 2) against
parthenon::BoundaryFunction::GenericBC ITSELF -- the function (and parthenon::IndexShape, which
gives the index ranges of the six boundary domains) is sliced out of
external/parthenon/src/bvals/boundary_conditions_generic.hpp:172-246 and mesh/domain.hpp at build
time and compiled against a mock of the few types it touches (oracle/ref_shim/bc).  The faces are
applied in Parthenon's order ix1, ox1, ix2, ox2, ix3, ox3 over the full transverse extent
(ApplyBoundaryConditionsOnCoarseOrFine, P:bvals/boundary_conditions.cpp), so corners inherit the
earlier faces' fills; every combination of outflow / reflect per face, 1-, 2- and 3-D blocks,
nghost 2 and 4, gas (FillGhost: density, velocity, sie) and dust (density, velocity)."""
import ctypes as C
import itertools
import os

import numpy as np
import pytest

from artemis_b200.enums import BoundaryFlag, Coordinates
from artemis_b200.mesh import UniformMesh
from oracle import oracle_py
from tests.helpers import dust_params, gas_params

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
BC_LIB = os.path.join(ROOT, "oracle", "_ref", "libbc_ref.so")
pytestmark = pytest.mark.skipif(not os.path.exists(BC_LIB), reason="oracle/_ref/libbc_ref.so not built")

O, R = BoundaryFlag.outflow, BoundaryFlag.reflect
_IP = C.POINTER(C.c_int)


def _reference_fill(L, mesh, a, ghost_vars, vec_dir, bcs):
    """GenericBC face by face on the FillGhost fields of the one block"""
    sub = np.ascontiguousarray(a[0, ghost_vars])
    vc = (C.c_int * len(vec_dir))(*vec_dir)
    nx = (C.c_int * 3)(*mesh.block_nx)
    for face in range(2 * mesh.ndim):
        L.ar_generic_bc(mesh.nghost, nx, len(ghost_vars), vc,
                        sub.ctypes.data_as(C.POINTER(C.c_double)), face // 2 + 1, face % 2,
                        0 if bcs[face] == O else 1)
    out = a.copy()
    out[0, ghost_vars] = sub
    return out


@pytest.mark.parametrize("fluid", ["gas", "dust"])
@pytest.mark.parametrize("ng", [2, 4])
@pytest.mark.parametrize("nx", [(12, 1, 1), (10, 8, 1), (8, 10, 6)])
def test_outflow_reflect_fill_equals_parthenons_generic_bc(nx, ng, fluid):
    L = C.CDLL(BC_LIB)
    L.ar_generic_bc.restype = None
    ndim = 1 + (nx[1] > 1) + (nx[2] > 1)
    Cc = Coordinates.cartesian
    fp = gas_params(Cc, "plm", "hlle", S=2) if fluid == "gas" else dust_params(Cc, "plm", "hlle", S=2)
    lib = oracle_py.lib()
    n = 0
    for combo in itertools.product((O, R), repeat=2 * ndim):
        bcs = tuple(combo) + (BoundaryFlag.periodic,) * (6 - 2 * ndim)
        mesh = UniformMesh(nx=nx, xmin=(0, 0, 0), xmax=(1, 1, 1), block_nx=nx, nghost=ng, bcs=bcs)
        fs = oracle_py.FluidState(mesh, fp, with_flux=False)
        rng = np.random.default_rng(100 + n)
        a = rng.standard_normal(mesh.shape(fp.nvar))
        want = _reference_fill(L, mesh, a, fs.ghost_vars, fs.vec_dir, bcs)
        got = a.copy()
        g = oracle_py.make_grid(mesh)
        vars_ = np.array(fs.ghost_vars, dtype=np.int32)
        vdir = np.array(fs.vec_dir, dtype=np.int32)
        bc = mesh.bc_ints()
        lib.ao_exchange_ghosts_phase(C.byref(g), 1, 1, 1, bc.ctypes.data_as(_IP), fp.nvar,
                                     oracle_py._p(got), len(vars_), vars_.ctypes.data_as(_IP),
                                     vdir.ctypes.data_as(_IP), 2)
        assert np.array_equal(got, want), (combo,)
        assert not np.array_equal(got, a)
        n += 1
    assert n == 4 ** ndim


@pytest.mark.parametrize("ndim,bcs", [
    (3, (BoundaryFlag.periodic,) * 6),
    (3, (R, O, BoundaryFlag.periodic, BoundaryFlag.periodic, O, R)),
    (2, (O, O, R, R, BoundaryFlag.periodic, BoundaryFlag.periodic)),
    (1, (R, O) + (BoundaryFlag.periodic,) * 4),
])
def test_uniform_mesh_exchange_equals_the_calcindices_plan(ndim, bcs):
    """The oracle's same-level ghost exchange on a uniform block lattice (its own send / receive
    ranges) must equal the executor of the multilevel plan on the same, unrefined lattice, whose
    index boxes are parthenon::CalcIndices (pinned to the reference's function in
    test_multilevel_plan.py): the same-level exchange is thereby pinned to it as well."""
    from artemis_b200.multilevel import MultilevelMesh, exchange_plan
    from oracle import multilevel_py
    nblk = tuple(3 if d < ndim else 1 for d in range(3))
    bnx = tuple((8, 10, 8)[d] if d < ndim else 1 for d in range(3))
    um = UniformMesh(nx=tuple(nblk[d] * bnx[d] for d in range(3)), xmin=(0, 0, 0), xmax=(1, 1, 1),
                     block_nx=bnx, nghost=2 if ndim < 3 else 4, bcs=bcs)
    mm = MultilevelMesh(root_blocks=nblk, block_nx=bnx, xmin=(0, 0, 0), xmax=(1, 1, 1), refine=(),
                        nghost=um.nghost, bcs=bcs)
    assert mm.nb == um.nb
    fp = gas_params(Coordinates.cartesian, "plm", "hlle", S=2)
    fs = oracle_py.FluidState(um, fp, with_flux=False)
    rng = np.random.default_rng(9)
    a = rng.standard_normal(um.shape(fp.nvar))
    # block b of the multilevel mesh sits at root location leaves[b][1]; find it in the lattice
    loc_u = {tuple(int(v) for v in l): b for b, l in enumerate(um.blk_loc)}
    perm = [loc_u[tuple(int(v) for v in mm.leaves[b][1])] for b in range(mm.nb)]
    fine = np.ascontiguousarray(a[perm])
    coarse = np.zeros(mm.coarse_shape(fp.nvar))
    kinds = {BoundaryFlag.periodic: "periodic", O: "outflow", R: "reflect"}
    multilevel_py.run_plan(mm, exchange_plan(mm), fine, coarse, fs.ghost_vars, fs.vec_dir,
                           [kinds[b] for b in bcs])
    got = a.copy()
    g = oracle_py.make_grid(um)
    vars_ = np.array(fs.ghost_vars, dtype=np.int32)
    vdir = np.array(fs.vec_dir, dtype=np.int32)
    bc = um.bc_ints()
    oracle_py.lib().ao_exchange_ghosts(C.byref(g), *[int(v) for v in um.lattice_n],
                                       bc.ctypes.data_as(_IP), fp.nvar, oracle_py._p(got),
                                       len(vars_), vars_.ctypes.data_as(_IP), vdir.ctypes.data_as(_IP))
    assert np.array_equal(got[perm], fine)
    assert not np.array_equal(got, a)
