"""Host-side mirror of the reference's deck selection logic, mesh bookkeeping and
integrator tables (no GPU)."""
import numpy as np
import pytest

from artemis_b200.enums import (INTEGRATORS, BoundaryFlag, Coordinates, CoordSelect, Fluid,
                                ReconstructionMethod, RSolver)
from artemis_b200.mesh import UniformMesh
from artemis_b200.params import ParameterInput, dust_params, gas_params

DECK = """
<artemis>
coordinates = spherical   # coordinate system
<parthenon/mesh>
nghost = 4
nx1 = 64
x1min = 0.4
x1max = 2.5
ix1_bc = outflow
ox1_bc = outflow
nx2 = 32
x2min = 1.0
x2max = 2.0
ix2_bc = reflect
ox2_bc = reflect
nx3 = 1
<parthenon/meshblock>
nx1 = 16
nx2 = 16
<gas>
cfl = 0.3
reconstruct = ppm
riemann = hlle
gamma = 1.4
<dust>
nspecies = 3
reconstruct = plm
riemann = llf
"""


def test_deck_selection_follows_gas_and_dust_initialize():
    pin = ParameterInput(DECK)
    gp = gas_params(pin, ndim=2)
    dp = dust_params(pin, ndim=2)
    assert gp.coords == Coordinates.spherical2D
    assert (gp.recon, gp.rsolver) == (ReconstructionMethod.ppm, RSolver.hlle)
    assert gp.cfl == 0.3 and abs(gp.gm1 - 0.4) < 1e-15 and gp.nvar == 6
    assert (dp.recon, dp.rsolver, dp.nspecies, dp.nvar) == (ReconstructionMethod.plm,
                                                            RSolver.llf, 3, 12)
    assert dp.fluid_type == Fluid.dust


def test_overrides_and_reference_error_messages():
    pin = ParameterInput(DECK, overrides=["gas/riemann=hllc", "parthenon/mesh/nghost=2"])
    with pytest.raises(ValueError, match="PPM requires at least 3 ghost cells"):
        gas_params(pin, 2)          # src/gas/gas.cpp:71
    pin = ParameterInput(DECK, overrides=["dust/riemann=hllc"])
    with pytest.raises(ValueError, match="Riemann solver \\(dust\\) not recognized"):
        dust_params(pin, 2)         # src/dust/dust.cpp:76-85: HLLC is gas-only
    with pytest.raises(ValueError, match="Coordinate type not recognized"):
        CoordSelect("toroidal", 3)
    assert CoordSelect("spherical", 1) == Coordinates.spherical1D
    assert CoordSelect("spherical", 3) == Coordinates.spherical3D


def test_mesh_from_input_matches_parthenon_index_shapes():
    pin = ParameterInput(DECK)
    m = UniformMesh.from_input(pin)
    assert m.ndim == 2 and m.nb == 4 * 2
    assert (m.ni, m.nj, m.nk) == (16 + 8, 16 + 8, 1)          # symmetry direction: no ghosts
    assert (m.is_, m.ie, m.js, m.je, m.ks, m.ke) == (4, 19, 4, 19, 0, 0)
    assert (m.fni, m.fnj, m.fnk) == (25, 25, 1)               # P:interface/metadata.cpp:378-387
    assert m.bcs[2] == BoundaryFlag.reflect
    # uniform_cartesian.hpp:30-36: xmin_ = block xmin - nghost*dx; faces tile the domain
    dx = (2.5 - 0.4) / 64
    assert np.allclose(m.blk_dx[:, 0], dx, rtol=1e-14)
    assert np.isclose(m.blk_xmin[0, 0], 0.4 - 4 * dx, rtol=1e-14)
    xf_last = m.face_positions(3, 0)
    assert np.isclose(xf_last[m.ie + 1], 2.5, rtol=1e-14)
    assert m.interior_zones == 64 * 32


def test_block_edges_are_bitwise_shared_between_neighbours():
    """Symmetrised logical->physical mapping (P:mesh/forest/logical_location.cpp:61-74): the
    upper face of block b is the same double as the lower face of block b+1."""
    m = UniformMesh(nx=(96, 1, 1), xmin=(-0.7, 0, 0), xmax=(1.9, 1, 1), block_nx=(8, 1, 1),
                    nghost=2)
    for b in range(m.nb - 1):
        hi = m.blk_lo[b, 0] + 8 * m.blk_dx[b, 0]
        assert abs(hi - m.blk_lo[b + 1, 0]) <= 2e-16 * max(1.0, abs(hi))


def test_integrator_tables():
    """P:time_integration/low_storage_integrator.cpp."""
    assert INTEGRATORS["rk2"] == ((0.0, 1.0, 1.0), (0.5, 0.5, 0.5))
    assert INTEGRATORS["vl2"] == ((0.0, 1.0, 0.5), (0.0, 1.0, 1.0))
    rk3 = INTEGRATORS["rk3"]
    assert rk3[1] == (0.25, 0.75, 0.25) and abs(rk3[2][0] - 2 / 3) < 1e-16
    for st in INTEGRATORS.values():
        for g0, g1, _ in st:
            assert abs(g0 + g1 - 1.0) < 1e-15     # consistency of the convex combination


def test_sub_lattice_partition_reproduces_global_block_geometry():
    g = UniformMesh(nx=(32, 32, 32), xmin=(-1, -1, -1), xmax=(1, 1, 1), block_nx=(8, 8, 8),
                    nghost=4)
    t = UniformMesh(nx=(32, 32, 32), xmin=(-1, -1, -1), xmax=(1, 1, 1), block_nx=(8, 8, 8),
                    nghost=4, lattice_lo=(2, 0, 2), lattice_n=(2, 2, 2))
    assert t.nb == 8
    for b in range(t.nb):
        l = t.blk_loc[b]
        gb = int(l[0] + 4 * (l[1] + 4 * l[2]))
        assert np.array_equal(t.blk_xmin[b], g.blk_xmin[gb])
        assert np.array_equal(t.blk_dx[b], g.blk_dx[gb])


def test_oracle_ic_boundary_keeps_the_profile_and_later_faces_copy_from_it():
    """AO_BC_IC (Disk::DiskBoundaryIC): every ghost zone of an ic face holds the generator's
    profile at its own position; a later outflow face copies the (fixed) zones of the corner"""
    import ctypes as C
    import numpy as np
    from artemis_b200.enums import BoundaryFlag as B, Coordinates
    from oracle import oracle_py
    from tests.helpers import gas_params, make_mesh, random_prim
    mesh = make_mesh(Coordinates.cartesian, 3, bcs=(B.fixed, B.fixed, B.outflow, B.outflow, B.periodic, B.periodic))
    gp = gas_params(Coordinates.cartesian, "plm", "hlle")
    sim = oracle_py.OracleSim(mesh, gas=gp)
    prim = random_prim(mesh, gp, seed=3)
    sim.gas.prim[:] = prim
    sim.nlim = 2
    sim.initialize()
    sim.run()
    ng = mesh.nghost
    inner = mesh.interior()
    lo = [b for b in range(mesh.nb) if b % mesh.lattice_n[0] == 0]
    vs = [0, 1, 2, 3, 5]   # FillGhost fields (pressure is derived)
    got, ic = sim.gas.prim[lo][:, vs], prim[lo][:, vs]
    assert np.array_equal(got[:, :, inner[0], inner[1], :ng], ic[:, :, inner[0], inner[1], :ng])
    # x2 outflow is applied after x1: the corner zones copy the first interior row of the slab
    lo2 = [b for b in lo if (b // mesh.lattice_n[0]) % mesh.lattice_n[1] == 0]
    g2 = sim.gas.prim[lo2][:, vs]
    for j in range(ng):
        assert np.array_equal(g2[:, :, inner[0], j, :ng], g2[:, :, inner[0], ng, :ng])
    assert not np.array_equal(sim.gas.prim[(slice(None), slice(None)) + inner],
                              prim[(slice(None), slice(None)) + inner])
