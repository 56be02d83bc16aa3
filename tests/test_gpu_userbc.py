"""GPU execution of the shearing-box user boundary conditions (inputs/ssheet/
ssheet.in = config 5's deck: `extrap` on the x1 and x3 faces, `inflow` on the x2 faces;
strat::ExtrapInnerX1 ... ExtrapOuterX3, src/pgen/strat.hpp:154-666).  The checker is the oracle's
ao_strat_bc, which tests/test_strat_bc_pin.py pins bit for bit to the reference's own six
functions.  Covered: uniform meshes (every physical face through ab200_block_bcs in Parthenon's
x1 -> x2 -> x3 order, user and generic conditions mixed), refined meshes (the conditions on the
coarse buffers and on the fine arrays inside the multilevel exchange), whole rk2 cycles on the
task path and the fused path, and the error behaviour.

Strict build: bit-identical, except the densities beyond an x3 face, which pass through pow()
(CUDA's against glibc's: a few ulp, tolerance 1e-14 relative) -- as in test_gpu_diffusion.py."""
import numpy as np
import pytest

from artemis_b200 import capi
from artemis_b200.driver import AddBoundaryExchangeTasks, ArtemisDriver, TaskStatus
from artemis_b200.enums import BoundaryFlag, Coordinates
from artemis_b200.mesh import UniformMesh
from artemis_b200.meshdata import MeshData
from artemis_b200.multilevel import MultilevelExchange, MultilevelMesh
from oracle.oracle_py import OracleSim
from tests.helpers import dust_params, gas_params, random_prim, zone_rel_err

pytestmark = pytest.mark.gpu

B = BoundaryFlag
STRAT = (B.extrap, B.extrap, B.inflow, B.inflow, B.extrap, B.extrap)   # ssheet.in:42-55
Q, OM0 = 1.5, 0.8
CART = Coordinates.cartesian
# the box straddles x1 = 0: both branches of the shear inflow exist on every x2 face
XMIN, XMAX = (-0.5, -0.4, -0.3), (0.5, 0.4, 0.3)


def _mesh(ndim, bcs, nb=2, bn=8, ng=4):
    nx = tuple(nb * bn if d < ndim else 1 for d in range(3))
    bnx = tuple(bn if d < ndim else 1 for d in range(3))
    return UniformMesh(nx=nx, xmin=XMIN, xmax=XMAX, block_nx=bnx, nghost=ng, bcs=tuple(bcs),
                       coords=CART)


def _close(got, want, S, strict):
    """strict: exact, densities (entries < S) to 1e-14; fast: 1e-13 of the array scale"""
    if strict:
        assert np.array_equal(got[:, S:], want[:, S:])
        assert np.max(np.abs(got[:, :S] - want[:, :S]) / np.abs(want[:, :S])) <= 1e-14
    else:
        assert np.max(np.abs(got - want)) <= 1e-13 * np.max(np.abs(want))


UNIFORM_CASES = [
    (3, STRAT),
    (2, STRAT[:4] + (B.periodic,) * 2),
    (3, (B.extrap, B.extrap, B.periodic, B.periodic, B.outflow, B.reflect)),
    (3, (B.outflow, B.extrap, B.inflow, B.reflect, B.extrap, B.outflow)),
]


@pytest.mark.parametrize("ndim,bcs", UNIFORM_CASES)
@pytest.mark.parametrize("variant", ["strict", "fast"])
def test_uniform_mesh_user_bcs_equal_the_oracle(ndim, bcs, variant):
    m = _mesh(ndim, bcs)
    gp, dp = gas_params(CART, "plm", "hlle"), dust_params(CART, "plm", "hlle", S=2)
    osim = OracleSim(m, gas=gp, dust=dp)
    osim.shear_bc = (Q, OM0)
    md = MeshData(m, gas=gp, dust=dp, variant=variant, materialize_fluxes=False, shear_bc=(Q, OM0))
    assert md.user_bcs
    for ff, of, seed in zip(md.fluids, osim.fluids, (3, 4)):
        prim = random_prim(m, ff.fp, seed=seed)
        mask = np.ones(prim.shape[2:], dtype=bool)
        mask[m.interior()] = False
        gv = of.ghost_vars
        for v in gv:                       # poison the ghost zones of every FillGhost entry
            prim[:, v][:, mask] = -777.0
        ff.prim.set(prim)
        of.prim[:] = prim
        osim.ExchangeGhosts(of)
    assert AddBoundaryExchangeTasks(md) == TaskStatus.complete
    for ff, of in zip(md.fluids, osim.fluids):
        got = ff.prim.get()
        assert not (got[:, of.ghost_vars] == -777.0).any(), "a ghost zone was never filled"
        _close(got, of.prim, ff.fp.nspecies, variant == "strict")
    assert md.launch_count() > 0
    md.close()


@pytest.mark.parametrize("mode,variant,ncyc", [("tasks", "strict", 2), ("fused", "strict", 2),
                                               ("tasks", "fast", 1), ("fused", "fast", 1)])
def test_rk2_cycles_with_the_ssheet_decks_boundaries(mode, variant, ncyc):
    """north_star's bar: per-zone relative difference <= 1e-12 after one cycle (default build);
    the strict build differs from the oracle only through pow() in the x3 density extrapolation
    and holds the same bar after two"""
    tol = 1e-12
    m = _mesh(3, STRAT)
    gp, dp = gas_params(CART, "ppm", "hllc"), dust_params(CART, "plm", "hlle", S=1)
    prim, dprim = random_prim(m, gp, seed=51), random_prim(m, dp, seed=52)
    osim = OracleSim(m, gas=gp, dust=dp)
    osim.shear_bc = (Q, OM0)
    osim.gas.prim[:] = prim
    osim.dust.prim[:] = dprim
    osim.nlim = ncyc
    osim.initialize()
    osim.run()
    md = MeshData(m, gas=gp, dust=dp, variant=variant, materialize_fluxes=(mode == "tasks"),
                  shear_bc=(Q, OM0))
    md.gas.prim.set(prim)
    md.dust.prim.set(dprim)
    drv = ArtemisDriver(md, "rk2", mode=mode, nlim=ncyc)
    drv.Initialize()
    drv.Execute()
    assert drv.ncycle == osim.ncycle == ncyc
    assert abs(drv.time - osim.time) <= 1e-12 * osim.time
    for ff, of in zip(md.fluids, osim.fluids):
        vref = None if of is osim.gas else 1.0
        assert zone_rel_err(ff.u0.get(), of.u0, of.fp, "cons", vref=vref) <= tol
        assert zone_rel_err(ff.prim.get(), of.prim, of.fp, "prim", vref=vref) <= tol
    md.close()


@pytest.mark.parametrize("ndim,refine", [(3, [(0, 0, 0), (3, 3, 3), (1, 2, 0)]), (2, [(0, 3, 0), (2, 1, 0)])])
@pytest.mark.parametrize("variant,entry_tables", [("strict", False), ("fast", False), ("strict", True)])
def test_refined_mesh_user_bcs_on_fine_arrays_and_coarse_buffers(ndim, refine, variant, entry_tables):
    """entry_tables: the coarse arrays of the entries reach the user conditions as one device
    pointer per pack entry (ab200_block_bc_desc.coarse_entries), the way a Parthenon host holds
    them (one coarse buffer per Variable), instead of one slab per fluid -- same bits"""
    from artemis_b200.multilevel import exchange_plan
    root = tuple(4 if d < ndim else 1 for d in range(3))
    bnx = tuple(8 if d < ndim else 1 for d in range(3))
    bcs = STRAT if ndim == 3 else STRAT[:4] + (B.periodic,) * 2
    m = MultilevelMesh(root_blocks=root, block_nx=bnx, xmin=XMIN, xmax=XMAX,
                       refine=tuple(tuple(r) for r in refine), nghost=4, bcs=bcs, coords=CART)
    plan = exchange_plan(m)
    assert plan.coarse_bcs and plan.fine_bcs
    gp, dp = gas_params(CART, "plm", "hlle"), dust_params(CART, "plm", "hlle", S=2)
    osim = OracleSim(m, gas=gp, dust=dp)
    osim.shear_bc = (Q, OM0)
    md = MeshData(m, gas=gp, dust=dp, variant=variant, materialize_fluxes=False, shear_bc=(Q, OM0))
    ex = MultilevelExchange(md, plan, entry_tables=entry_tables)
    assert bool(getattr(ex, "_entry_tabs", [])) == entry_tables
    for ff, of, seed in zip(md.fluids, osim.fluids, (21, 22)):
        prim = random_prim(m, ff.fp, seed=seed)
        ff.prim.set(prim)
        of.prim[:] = prim
        # the oracle's coarse buffers start from zeros; so do the library's
        ex.coarse[int(ff.fp.fluid_type)].set(np.zeros(m.coarse_shape(ff.fp.nvar)))
        osim.ExchangeGhosts(of)
    ex.exchange()
    for ff, of in zip(md.fluids, osim.fluids):
        _close(ff.prim.get(), of.prim, ff.fp.nspecies, variant == "strict")
    ex.close()
    md.close()


def test_user_bc_error_behaviour():
    m = _mesh(3, STRAT)
    gp = gas_params(CART, "plm", "hlle")
    # `inflow` without StratParams: the task fails, loudly, and says why
    md = MeshData(m, gas=gp, materialize_fluxes=False)
    md.gas.prim.set(random_prim(m, gp, seed=1))
    assert AddBoundaryExchangeTasks(md) == TaskStatus.fail
    assert "ab200_set_shear_bc_params" in md.last_error
    # strat.hpp registers extrap on x1 / x3 and inflow on x2 only; a user condition takes the
    # whole fluid
    md.call("ab200_set_shear_bc_params", Q, OM0)
    for desc in (capi.BlockBcDesc(0, 0, 0, gp.nvar, 2, int(B.extrap), None),
                 capi.BlockBcDesc(0, 0, 0, gp.nvar, 0, int(B.inflow), None),
                 capi.BlockBcDesc(0, 0, 0, 4, 0, int(B.extrap), None)):
        with pytest.raises(capi.AB200Error):
            md.call("ab200_block_bcs", (capi.BlockBcDesc * 1)(desc), 1)
    # per-entry coarse tables belong to the user conditions
    with pytest.raises(capi.AB200Error):
        md.call("ab200_block_bcs", (capi.BlockBcDesc * 1)(
            capi.BlockBcDesc(0, 0, 0, 4, 0, int(B.outflow), None, md.gas.prim.ptr)), 1)
    # the device-resident cycle has no state-dependent conditions in its fused ghost fill
    drv = ArtemisDriver(md, "rk2", mode="fused", nlim=1)
    with pytest.raises(capi.AB200Error):
        drv.StepDevice()
    with pytest.raises(capi.AB200Error):
        md.call("ab200_fill_ghosts")
    md.close()
