"""Shearing-box user boundary conditions, host logic (no GPU): the two CPU executors that check
the GPU path -- ao_exchange_ghosts_user on a uniform block lattice and oracle/multilevel_py's
run_plan on the same lattice expressed as an unrefined MultilevelMesh -- must agree bit for bit
(face order x1 -> x2 -> x3, full transverse extent, fluid-wide conditions), every ghost zone of
every FillGhost entry must be filled on uniform and refined meshes, and the host mirror's
descriptor lists must name every physical face once."""
import numpy as np
import pytest

from artemis_b200.enums import BoundaryFlag as B, Coordinates
from artemis_b200.mesh import UniformMesh
from artemis_b200.multilevel import MultilevelMesh, exchange_plan
from oracle.oracle_py import OracleSim
from tests.helpers import dust_params, gas_params, random_prim

STRAT = (B.extrap, B.extrap, B.inflow, B.inflow, B.extrap, B.extrap)
CART = Coordinates.cartesian
XMIN, XMAX = (-0.5, -0.4, -0.3), (0.5, 0.4, 0.3)


def _fluids():
    return gas_params(CART, "plm", "hlle"), dust_params(CART, "plm", "hlle", S=2)


def _poisoned(mesh, of, seed):
    p = random_prim(mesh, of.fp, seed=seed)
    mask = np.ones(p.shape[2:], dtype=bool)
    mask[mesh.interior()] = False
    for v in of.ghost_vars:
        p[:, v][:, mask] = -777.0
    return p


@pytest.mark.parametrize("ndim", [2, 3])
def test_uniform_and_multilevel_executors_agree(ndim):
    bcs = STRAT if ndim == 3 else STRAT[:4] + (B.periodic,) * 2
    root = tuple(2 if d < ndim else 1 for d in range(3))
    bnx = tuple(8 if d < ndim else 1 for d in range(3))
    mm = MultilevelMesh(root_blocks=root, block_nx=bnx, xmin=XMIN, xmax=XMAX, refine=(), nghost=4,
                        bcs=bcs, coords=CART)
    mu = UniformMesh(nx=tuple(r * n for r, n in zip(root, bnx)), xmin=XMIN, xmax=XMAX, block_nx=bnx,
                     nghost=4, bcs=bcs, coords=CART)
    gp, dp = _fluids()
    o1, o2 = OracleSim(mu, gas=gp, dust=dp), OracleSim(mm, gas=gp, dust=dp)
    o1.shear_bc = o2.shear_bc = (1.5, 0.8)
    order = [int(np.argmin(np.abs(mu.blk_xmin - mm.blk_xmin[b]).sum(1))) for b in range(mm.nb)]
    assert sorted(order) == list(range(mu.nb))
    for f1, f2, seed in zip(o1.fluids, o2.fluids, (3, 4)):
        p = _poisoned(mm, f2, seed)
        f2.prim[:] = p
        f1.prim[order] = p
        o1.ExchangeGhosts(f1)
        o2.ExchangeGhosts(f2)
        assert not (f1.prim[:, f1.ghost_vars] == -777.0).any()
        assert np.array_equal(f1.prim[order], f2.prim)


@pytest.mark.parametrize("ndim,refine", [(3, ((1, 1, 1), (0, 0, 0))), (2, ((1, 1, 0), (3, 0, 0)))])
def test_refined_mesh_fills_every_ghost_zone(ndim, refine):
    bcs = STRAT if ndim == 3 else STRAT[:4] + (B.periodic,) * 2
    root = tuple(4 if d < ndim else 1 for d in range(3))
    bnx = tuple(8 if d < ndim else 1 for d in range(3))
    mm = MultilevelMesh(root_blocks=root, block_nx=bnx, xmin=XMIN, xmax=XMAX, refine=refine,
                        nghost=4, bcs=bcs, coords=CART)
    plan = exchange_plan(mm)
    assert plan.coarse_bcs and plan.fine_bcs
    gp, dp = _fluids()
    o = OracleSim(mm, gas=gp, dust=dp)
    o.shear_bc = (1.5, 0.8)
    for of, seed in zip(o.fluids, (5, 6)):
        of.prim[:] = _poisoned(mm, of, seed)
        o.ExchangeGhosts(of)
        assert not (of.prim[:, of.ghost_vars] == -777.0).any()
        assert np.isfinite(of.prim).all()


def test_two_cycles_with_the_decks_boundaries_stay_finite():
    mu = UniformMesh(nx=(16, 16, 16), xmin=XMIN, xmax=XMAX, block_nx=(8, 8, 8), nghost=4, bcs=STRAT,
                     coords=CART)
    gp, dp = gas_params(CART, "ppm", "hllc"), dust_params(CART, "plm", "hlle", S=1)
    o = OracleSim(mu, gas=gp, dust=dp)
    o.shear_bc = (1.5, 0.8)
    o.gas.prim[:] = random_prim(mu, gp, seed=51)
    o.dust.prim[:] = random_prim(mu, dp, seed=52)
    o.nlim = 2
    o.initialize()
    o.run()
    assert o.ncycle == 2 and np.isfinite(o.gas.u0).all() and np.isfinite(o.dust.u0).all()


def test_input_deck_names_map_to_the_user_flags():
    from artemis_b200.params import ParameterInput
    deck = "\n".join(["<parthenon/mesh>", "nx1 = 16", "nx2 = 16", "nx3 = 16", "nghost = 4",
                      "ix1_bc = extrap", "ox1_bc = extrap", "ix2_bc = inflow", "ox2_bc = inflow",
                      "ix3_bc = extrap", "ox3_bc = extrap", "<parthenon/meshblock>", "nx1 = 8",
                      "nx2 = 8", "nx3 = 8", "<artemis>", "coordinates = cartesian"])
    pin = ParameterInput(deck)
    m = UniformMesh.from_input(pin)
    assert tuple(m.bcs) == STRAT


def test_host_mirror_names_every_physical_face_once():
    """MeshData._physical_bc_list (no GPU needed: it only reads the mesh and the fluids): one
    descriptor per (fluid, boundary block, non-periodic face) for a user condition -- the whole
    fluid -- and one per FillGhost range for a generic condition; interior blocks get none"""
    import types

    from artemis_b200.meshdata import MeshData
    bcs = (B.outflow, B.extrap, B.inflow, B.reflect, B.periodic, B.periodic)
    mu = UniformMesh(nx=(24, 16, 8), xmin=XMIN, xmax=XMAX, block_nx=(8, 8, 8), nghost=4, bcs=bcs,
                     coords=CART)
    gp, dp = _fluids()
    fake = types.SimpleNamespace(mesh=mu, fluids=[types.SimpleNamespace(fp=gp),
                                                  types.SimpleNamespace(fp=dp)])
    lst = MeshData._physical_bc_list(fake, mu.bc_ints())
    nbx, nby, nbz = mu.lattice_n
    seen = {}
    for d in lst:
        lb = (d.block % nbx, (d.block // nbx) % nby, d.block // (nbx * nby))
        axis, outer = d.face // 2, d.face % 2
        assert lb[axis] == ((nbx, nby, nbz)[axis] - 1 if outer else 0)
        assert d.type == int(bcs[d.face]) and not d.coarse and not d.coarse_entries
        nvar = (gp if d.fluid == 0 else dp).nvar
        if d.type >= int(B.extrap):
            assert (d.var0, d.ncomp) == (0, nvar)
        seen.setdefault((d.fluid, d.block, d.face), []).append((d.var0, d.ncomp))
    # x3 is periodic: faces 4 and 5 never appear; every boundary (block, face) appears
    assert all(k[2] < 4 for k in seen)
    per_face = {0: nby * nbz, 1: nby * nbz, 2: nbx * nbz, 3: nbx * nbz}
    for fluid in (0, 1):
        for face, n in per_face.items():
            assert sum(1 for k in seen if k[0] == fluid and k[2] == face) == n
    # generic faces: gas = (density + velocity) and sie, dust = one range
    assert sorted(seen[(0, 0, 0)]) == [(0, 4), (5, 1)]
    assert seen[(1, 0, 0)] == [(0, 8)]


def test_user_conditions_on_a_split_mesh_are_refused_not_mis_exchanged():
    """AB200_BC_NONE means "another rank" to the transport and "user condition" to the host
    mirror; until the two are told apart the exchange task fails loudly (TaskStatus.fail with a
    reason) instead of exchanging across a physical boundary"""
    import types

    from artemis_b200.driver import AddBoundaryExchangeTasks, TaskStatus
    md = types.SimpleNamespace(user_bcs=True)
    assert AddBoundaryExchangeTasks(md, comm=object()) == TaskStatus.fail
    assert "split across ranks" in md.last_error
