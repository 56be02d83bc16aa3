"""Pointwise source terms between FluxSource and SetAuxillaryFields:
the oracle's restatement of Gravity::UniformGravity (src/gravity/uniform.cpp:28-90) and
RotatingFrame::ShearingBoxImpl (src/rotating_frame/rotating_frame_impl.hpp:28-94) against the
reference's OWN code compiled in oracle/_ref -- bit for bit, all geometries, gas + dust, over a
whole rk2 cycle; and the reference's own checks on the drag restatement (no reference build of
drag.hpp here): total momentum conserved to 1e-13 and relaxation to the common velocity
(tst/scripts/drag/drag.py:57-137)."""
import numpy as np
import pytest

from artemis_b200.enums import BoundaryFlag, Coordinates
from oracle import ref_py
from oracle.oracle_py import OracleSim, make_diffusion, make_drag
from tests.helpers import dust_params, gas_params, make_mesh, random_prim

needs_ref = pytest.mark.skipif(not ref_py.available(), reason="oracle/_ref not built")


def _pair(coords, sources, integ="rk2", ncyc=1, ndim=3):
    bcs = (BoundaryFlag.outflow,) * 6 if coords != Coordinates.cartesian else None
    mesh = make_mesh(coords, ndim, bcs=bcs)
    gp, dp = gas_params(coords, "plm", "hlle"), dust_params(coords, "plm", "hlle", S=2)
    sims = []
    for cls in (OracleSim, ref_py.RefSim):
        sim = cls(mesh, gas=gp, dust=dp, integrator=integ)
        sim.gas.prim[:] = random_prim(mesh, gp, seed=61)
        sim.dust.prim[:] = random_prim(mesh, dp, seed=62)
        sim.sources = list(sources)
        sim.nlim = ncyc
        sim.initialize()
        sim.run()
        sims.append(sim)
    return sims


@needs_ref
@pytest.mark.parametrize("coords", [Coordinates.cartesian, Coordinates.cylindrical,
                                    Coordinates.spherical3D, Coordinates.axisymmetric])
def test_uniform_gravity_restatement_is_bit_identical_to_reference_code(coords):
    o, r = _pair(coords, [("gravity", 0.3, -0.2, 0.1)])
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim)
    # and the source really acted
    base, _ = _pair(coords, [])
    assert not np.array_equal(base.gas.u0, o.gas.u0)


@needs_ref
def test_shearing_box_restatement_is_bit_identical_to_reference_code():
    o, r = _pair(Coordinates.cartesian, [("shearing_box", 1.0, 1.5), ("gravity", 0.0, 0.0, -0.4)],
                 integ="vl2", ncyc=2)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim)


ALL_GEOMS = [(Coordinates.cartesian, 3), (Coordinates.cartesian, 2), (Coordinates.cylindrical, 3),
             (Coordinates.spherical1D, 1), (Coordinates.spherical2D, 2),
             (Coordinates.spherical3D, 3), (Coordinates.axisymmetric, 2)]


@needs_ref
@pytest.mark.parametrize("coords,ndim", ALL_GEOMS)
@pytest.mark.parametrize("sink", [False, True])
def test_point_mass_gravity_restatement_is_bit_identical_to_reference_code(coords, ndim, sink):
    """Gravity::PointMassGravity<GEOM> (src/gravity/point_mass.cpp:26-196): off-centre mass with
    softening; with `sink` the accretion radius covers part of the mesh."""
    pm = ("point_mass", 0.7, 0.11, -0.07, 0.05, 0.03) + ((40.0, 2.5) if sink else (0.0, 0.0))
    o, r = _pair(coords, [pm], ndim=ndim)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim)
    base, _ = _pair(coords, [], ndim=ndim)
    assert not np.array_equal(base.gas.u0, o.gas.u0)
    if sink:   # the sink removed mass somewhere
        nos, _ = _pair(coords, [pm[:6] + (0.0, 0.0)], ndim=ndim)
        assert np.any(o.gas.u0[:, 0] < nos.gas.u0[:, 0])


@needs_ref
@pytest.mark.parametrize("coords,ndim", [g for g in ALL_GEOMS if g[0] != Coordinates.cartesian])
def test_rotating_frame_restatement_is_bit_identical_to_reference_code(coords, ndim):
    """RotatingFrame::RotatingFrameImpl<GEOM> (rotating_frame_impl.hpp:96-199) on the density
    fluxes of the stage, together with the point mass (the disk decks' pair of sources)."""
    src = [("rotating_frame", 0.8), ("point_mass", 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0)]
    o, r = _pair(coords, src, integ="vl2", ncyc=2, ndim=ndim)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim)
    base, _ = _pair(coords, src[1:], integ="vl2", ncyc=2, ndim=ndim)
    assert not np.array_equal(base.gas.u0, o.gas.u0)


def _drag_pair(coords, ndim, drag_kw, diffusion=None, integ="rk2", ncyc=2):
    bcs = (BoundaryFlag.outflow,) * 6 if coords != Coordinates.cartesian else None
    mesh = make_mesh(coords, ndim, bcs=bcs)
    gp, dp = gas_params(coords, "plm", "hlle"), dust_params(coords, "plm", "hlle", S=3)
    sims = []
    for cls in (OracleSim, ref_py.RefSim):
        sim = cls(mesh, gas=gp, dust=dp, integrator=integ)
        sim.gas.prim[:] = random_prim(mesh, gp, seed=61)
        sim.dust.prim[:] = random_prim(mesh, dp, seed=62)
        if diffusion:
            sim.diffusion = make_diffusion(**diffusion)
        sim.sources = [("drag_model", make_drag(mesh, **drag_kw))]
        sim.nlim = ncyc
        sim.initialize()
        sim.run()
        sims.append(sim)
    return sims


def _zones(mesh):
    """damping zones covering the inner / outer ~third of every direction"""
    lo, hi = np.array(mesh.xmin, float), np.array(mesh.xmax, float)
    w = hi - lo
    return dict(inner=tuple(lo + 0.3 * w), outer=tuple(hi - 0.3 * w), inner_rate=(3.0, 2.0, 1.5),
                outer_rate=(2.5, 0.0, 4.0))


DRAG_CASES = {
    "constant_tau": lambda m: dict(tau=(1e-3, 0.2, 5.0), scale=0.7),
    "constant_tau_with_an_instantly_coupled_species": lambda m: dict(tau=(0.0, 0.2, 5.0)),
    "stokes": lambda m: dict(model="stokes", scale=0.4, grain_density=2.5, sizes=(1e-3, 3e-2, 0.4)),
    "constant_tau_with_damping_zones": lambda m: dict(tau=(1e-2, 0.2, 5.0), gas_damping=_zones(m),
                                                      dust_damping=_zones(m)),
    "stokes_with_damping_to_the_viscous_inflow": lambda m: dict(
        model="stokes", scale=0.4, grain_density=2.5, sizes=(1e-3, 3e-2, 0.4),
        gas_damping=_zones(m), dust_damping=_zones(m), damp_to_visc=True),
    "self_damping": lambda m: dict(coupling="self", gas_damping=_zones(m), dust_damping=_zones(m)),
    "self_damping_to_the_viscous_inflow": lambda m: dict(
        coupling="self", gas_damping=_zones(m), dust_damping=_zones(m), damp_to_visc=True),
}


@needs_ref
@pytest.mark.parametrize("coords,ndim", ALL_GEOMS)
@pytest.mark.parametrize("case", sorted(DRAG_CASES))
def test_drag_restatement_is_bit_identical_to_reference_code(coords, ndim, case):
    """Drag::Initialize + Drag::DragSource<GEOM> (src/drag/drag.cpp, drag.hpp) run from the
    reference's own sources on a parsed parameter set: simple_dust and self coupling, constant and
    Stokes stopping times, damping zones, damping towards the viscous inflow velocity."""
    mesh = make_mesh(coords, ndim)
    kw = DRAG_CASES[case](mesh)
    visc = dict(visc=("constant", 3e-3)) if kw.get("damp_to_visc") else None
    o, r = _drag_pair(coords, ndim, kw, diffusion=visc)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim)


@needs_ref
def test_legacy_constant_drag_entry_is_the_reference_drag():
    """("drag", [tau]) -- what the GPU's ab200_drag_simple is checked against -- now runs the
    reference's own DragSource on the reference side."""
    o, r = _pair(Coordinates.cylindrical, [("drag", [0.05, 2.0])], integ="rk3", ncyc=2)
    for fo, fr in zip(o.fluids, r.fluids):
        assert np.array_equal(fo.u0, fr.u0) and np.array_equal(fo.prim, fr.prim)


def test_drag_conserves_total_momentum_and_relaxes_to_the_common_velocity():
    """inputs/drag/simple_drag.in physics: constant stopping times, gas rho 10 v (1,0,0), four
    dust species rho 0.01 at rest, uniform state => no fluxes, only the implicit drag update."""
    mesh = make_mesh(Coordinates.cartesian, 3)
    gp = gas_params(Coordinates.cartesian, "plm", "hlle")
    dp = dust_params(Coordinates.cartesian, "plm", "hlle", S=4)
    sim = OracleSim(mesh, gas=gp, dust=dp)
    sim.gas.prim[:] = 0.0
    sim.gas.prim[:, 0] = 10.0
    sim.gas.prim[:, 1] = 1.0
    sim.gas.prim[:, 5] = 1.0
    sim.gas.prim[:, 4] = gp.gm1 * 10.0
    sim.dust.prim[:] = 0.0
    sim.dust.prim[:, :4] = 0.01
    tau = [1e-3, 1e-2, 1e-1, 1.0]
    sim.sources = [("drag", tau)]
    sim.initialize()
    sl = (slice(None),) + mesh.interior()
    p0 = sim.gas.u0[:, 1][sl] + sum(sim.dust.u0[:, 4 + 3 * n][sl] for n in range(4))
    sim.nlim = 40
    sim.run()
    p1 = sim.gas.u0[:, 1][sl] + sum(sim.dust.u0[:, 4 + 3 * n][sl] for n in range(4))
    assert np.max(np.abs(p1 - p0) / np.abs(p0)) <= 1e-13        # drag.py:135-137
    vd = [float(np.mean(sim.dust.prim[:, 4 + 3 * n][sl])) for n in range(4)]
    vg = float(np.mean(sim.gas.prim[:, 1][sl]))
    assert 0.0 < vd[3] < min(vd[:3])                             # the longest tau lags behind
    vcom = 10.0 / (10.0 + 0.04)
    assert abs(vd[0] - vg) < 1e-4 and abs(vg - vcom) < 1e-2
