"""GPU parity of the pointwise source terms and of the split stage that hosts
them: ab200_fused_stage(AB200_STAGE_DEFER_C2P) -> ab200_uniform_gravity / ab200_shearing_box /
ab200_drag_simple / ab200_point_mass_gravity / ab200_rotating_frame -> ab200_finish_stage,
against the oracle (gravity, point mass, shearing box and rotating frame pinned bit for bit to the
reference's own code, tests/test_sources_oracle.py)."""
import numpy as np
import pytest

from artemis_b200.driver import ArtemisDriver
from artemis_b200.enums import BoundaryFlag, Coordinates
from artemis_b200.meshdata import MeshData
from artemis_b200 import capi
from oracle.oracle_py import OracleSim, make_diffusion, make_drag
from tests.helpers import dust_params, gas_params, make_mesh, random_prim, zone_rel_err

pytestmark = pytest.mark.gpu

CASES = [
    (Coordinates.cartesian, [("gravity", 0.3, -0.2, 0.1)], "rk2"),
    (Coordinates.cartesian, [("shearing_box", 1.0, 1.5), ("gravity", 0.0, 0.0, -0.4)], "vl2"),
    (Coordinates.cartesian, [("drag", [1e-3, 0.5])], "rk2"),
    (Coordinates.cylindrical, [("gravity", -0.5, 0.0, 0.2), ("drag", [0.05, 2.0])], "rk3"),
    (Coordinates.spherical3D, [("gravity", -0.7, 0.0, 0.0)], "rk2"),
    # Gravity::PointMassGravity (off-centre, softened, with and without the mass sink) and the
    # curvilinear RotatingFrameImpl on the mass fluxes of the stage -- the disk decks' sources
    (Coordinates.cartesian, [("point_mass", 0.7, 0.11, -0.07, 0.05, 0.03, 40.0, 2.5)], "rk2"),
    (Coordinates.cylindrical, [("point_mass", 1.0, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0),
                               ("rotating_frame", 0.8)], "vl2"),
    (Coordinates.spherical3D, [("point_mass", 0.7, 0.11, -0.07, 0.05, 0.03, 40.0, 2.5),
                               ("rotating_frame", 0.6), ("drag", [0.05, 2.0])], "rk2"),
    (Coordinates.spherical2D, [("point_mass", 1.0, 0.0, 0.0, 0.0, 0.02, 0.0, 0.0),
                               ("rotating_frame", -0.7)], "rk2"),
    (Coordinates.spherical1D, [("point_mass", 1.0, 0.0, 0.0, 0.0, 0.0, 30.0, 1.2),
                               ("rotating_frame", 0.5)], "rk2"),
    (Coordinates.axisymmetric, [("point_mass", 0.9, 0.0, 0.0, 0.0, 0.03, 25.0, 1.0),
                                ("rotating_frame", 0.9)], "rk3"),
]
NDIM = {Coordinates.axisymmetric: 2}


def _zones(mesh):
    lo, hi = np.array(mesh.xmin, float), np.array(mesh.xmax, float)
    w = hi - lo
    return dict(inner=tuple(lo + 0.3 * w), outer=tuple(hi - 0.3 * w), inner_rate=(3.0, 2.0, 1.5),
                outer_rate=(2.5, 0.0, 4.0))


# Drag::DragSource in full: (coordinates, make_drag keywords as a function of the mesh, viscosity)
DRAG_CASES = [
    (Coordinates.cartesian, lambda m: dict(model="stokes", scale=0.4, grain_density=2.5,
                                           sizes=(1e-3, 0.4)), None),
    (Coordinates.cylindrical, lambda m: dict(tau=(1e-2, 5.0), scale=0.7, gas_damping=_zones(m),
                                             dust_damping=_zones(m)), None),
    (Coordinates.spherical3D, lambda m: dict(model="stokes", scale=0.4, grain_density=2.5,
                                             sizes=(1e-3, 0.4), gas_damping=_zones(m),
                                             dust_damping=_zones(m), damp_to_visc=True),
     dict(visc=("constant", 3e-3))),
    (Coordinates.axisymmetric, lambda m: dict(coupling="self", gas_damping=_zones(m),
                                              dust_damping=_zones(m)), None),
    (Coordinates.spherical2D, lambda m: dict(coupling="self", gas_damping=_zones(m),
                                             dust_damping=_zones(m), damp_to_visc=True),
     dict(visc=("constant", 2e-3))),
]


def _to_gpu_sources(sources):
    out = []
    for src in sources:
        if src[0] == "drag_model":   # oracle ao_drag and ab200_drag_desc share their layout
            out.append(("drag_model", capi.DragDesc.from_buffer_copy(bytes(src[1]))))
        else:
            out.append(src)
    return out


def _run(coords, sources, integ, mode, variant, ncyc=2, diffusion=None):
    bcs = (BoundaryFlag.outflow,) * 6 if coords != Coordinates.cartesian else None
    mesh = make_mesh(coords, NDIM.get(coords, 3), bcs=bcs)
    if callable(sources):
        sources = sources(mesh)
    gp, dp = gas_params(coords, "ppm", "hllc"), dust_params(coords, "plm", "hlle", S=2)
    prim, dprim = random_prim(mesh, gp, seed=71), random_prim(mesh, dp, seed=72)
    osim = OracleSim(mesh, gas=gp, dust=dp, integrator=integ)
    osim.gas.prim[:] = prim
    osim.dust.prim[:] = dprim
    osim.sources = list(sources)
    gdiff = None
    if diffusion:
        osim.diffusion = make_diffusion(**diffusion)
        gdiff = capi.DiffusionDesc.from_buffer_copy(bytes(osim.diffusion))
    osim.nlim = ncyc
    osim.initialize()
    osim.run()
    md = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=(mode == "tasks"))
    md.gas.prim.set(prim)
    md.dust.prim.set(dprim)
    drv = ArtemisDriver(md, integ, mode=mode, nlim=ncyc, sources=_to_gpu_sources(sources),
                        diffusion=gdiff)
    drv.Initialize()
    drv.Execute()
    out = [(f.u0.get(), f.prim.get(), of.u0, of.prim, of.fp) for f, of in zip(md.fluids, osim.fluids)]
    assert md.launch_count() > 0 and drv.ncycle == osim.ncycle == ncyc
    md.close()
    return out, drv, osim


@pytest.mark.parametrize("coords,sources,integ", CASES)
def test_tasks_with_sources_strict_bit_identical(coords, sources, integ):
    out, drv, osim = _run(coords, sources, integ, "tasks", "strict")
    assert drv.dt == osim.dt
    for u0, w, ou0, ow, _ in out:
        assert np.array_equal(u0, ou0) and np.array_equal(w, ow)


@pytest.mark.parametrize("coords,sources,integ", CASES)
def test_split_fused_stage_with_sources_within_1e12(coords, sources, integ):
    """fast build, fused passes + deferred C2P, two cycles"""
    out, drv, osim = _run(coords, sources, integ, "fused", "fast", ncyc=1)
    for u0, w, ou0, ow, fp in out:
        assert zone_rel_err(u0, ou0, fp, "cons") <= 1e-12
        assert zone_rel_err(w, ow, fp, "prim") <= 1e-12


@pytest.mark.parametrize("coords,drag_kw,visc", DRAG_CASES)
def test_full_drag_source_strict_bit_identical(coords, drag_kw, visc):
    """ab200_drag_source (Stokes / constant stopping times, damping zones, viscous target, self
    coupling) against the oracle, which is pinned to the reference's own drag.cpp"""
    out, drv, osim = _run(coords, lambda m: [("drag_model", make_drag(m, **drag_kw(m)))], "rk2",
                          "tasks", "strict", diffusion=visc)
    assert drv.dt == osim.dt
    for u0, w, ou0, ow, _ in out:
        assert np.array_equal(u0, ou0) and np.array_equal(w, ow)


@pytest.mark.parametrize("coords,drag_kw,visc", DRAG_CASES)
def test_full_drag_source_fused_within_1e12(coords, drag_kw, visc):
    out, drv, osim = _run(coords, lambda m: [("drag_model", make_drag(m, **drag_kw(m)))], "rk2",
                          "fused", "fast", ncyc=1, diffusion=visc)
    for u0, w, ou0, ow, fp in out:
        assert zone_rel_err(u0, ou0, fp, "cons") <= 1e-12
        assert zone_rel_err(w, ow, fp, "prim") <= 1e-12


def test_drag_on_the_gpu_conserves_total_momentum():
    """tst/scripts/drag/drag.py:135-137: total (gas + dust) momentum conserved to 1e-13."""
    mesh = make_mesh(Coordinates.cartesian, 3)
    gp = gas_params(Coordinates.cartesian, "plm", "hlle")
    dp = dust_params(Coordinates.cartesian, "plm", "hlle", S=4)
    prim = np.zeros(mesh.shape(6))
    prim[:, 0], prim[:, 1], prim[:, 5] = 10.0, 1.0, 1.0
    prim[:, 4] = gp.gm1 * 10.0
    dprim = np.zeros(mesh.shape(16))
    dprim[:, :4] = 0.01
    md = MeshData(mesh, gas=gp, dust=dp, materialize_fluxes=False)
    md.gas.prim.set(prim)
    md.dust.prim.set(dprim)
    drv = ArtemisDriver(md, "rk2", mode="fused", nlim=40, sources=[("drag", [1e-3, 1e-2, 1e-1, 1.0])])
    drv.Initialize()
    sl = (slice(None),) + mesh.interior()
    gu, du = md.gas.u0.get(), md.dust.u0.get()
    p0 = gu[:, 1][sl] + sum(du[:, 4 + 3 * n][sl] for n in range(4))
    drv.Execute()
    gu, du = md.gas.u0.get(), md.dust.u0.get()
    p1 = gu[:, 1][sl] + sum(du[:, 4 + 3 * n][sl] for n in range(4))
    assert np.max(np.abs(p1 - p0) / np.abs(p0)) <= 1e-13
    vd = md.dust.prim.get()
    assert 0.0 < np.mean(vd[:, 4 + 9][sl]) < np.mean(vd[:, 4][sl])   # tau = 1 lags tau = 1e-3
    md.close()


@pytest.mark.parametrize("variant", ["strict", "fast"])
def test_device_resident_cycles_with_configured_sources(variant):
    """ab200_configure_sources + ab200_run_cycles (split stages, beta*dt formed on the device)
    == the host-driven fused loop with the same sources; strict: bit for bit."""
    import ctypes as C
    from artemis_b200 import capi
    coords = Coordinates.cartesian
    mesh = make_mesh(coords, 3)
    gp, dp = gas_params(coords, "ppm", "hllc"), dust_params(coords, "plm", "hlle", S=2)
    prim, dprim = random_prim(mesh, gp, seed=81), random_prim(mesh, dp, seed=82)
    sources = [("gravity", 0.1, 0.0, -0.3), ("shearing_box", 0.8, 1.5), ("drag", [0.02, 0.7])]
    ncyc = 3
    md1 = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
    md1.gas.prim.set(prim)
    md1.dust.prim.set(dprim)
    d1 = ArtemisDriver(md1, "rk2", mode="fused", nlim=ncyc, sources=sources)
    d1.Initialize()
    d1.Execute()
    md2 = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
    md2.gas.prim.set(prim)
    md2.dust.prim.set(dprim)
    d2 = ArtemisDriver(md2, "rk2", mode="fused")
    d2.Initialize()
    sd = capi.SourcesDesc()
    sd.gravity, sd.g[0], sd.g[1], sd.g[2] = 1, 0.1, 0.0, -0.3
    sd.shearing_box, sd.omega, sd.qshear = 1, 0.8, 1.5
    sd.drag, sd.ntau, sd.tau[0], sd.tau[1] = 1, 2, 0.02, 0.7
    md2.call("ab200_configure_sources", C.byref(sd))
    md2.set_time_state(d2.dt)
    md2.call("ab200_run_cycles", 1, ncyc, float(np.finfo(np.float64).max))
    ts = md2.time_state()
    assert int(ts[3]) == ncyc
    for f1, f2 in zip(md1.fluids, md2.fluids):
        if variant == "strict":
            assert np.array_equal(f1.u0.get(), f2.u0.get())
            assert np.array_equal(f1.prim.get(), f2.prim.get())
        else:
            assert zone_rel_err(f2.u0.get(), f1.u0.get(), f1.fp, "cons") <= 1e-12
    if variant == "strict":
        assert ts[0] == d1.dt and ts[2] == d1.time
    md1.close()
    md2.close()


@pytest.mark.parametrize("variant", ["strict", "fast"])
def test_device_resident_cycles_with_point_mass_and_rotating_frame(variant):
    """The disk decks' sources through ab200_configure_sources + ab200_run_cycles on a spherical
    mesh: every stage taps the mass fluxes of its three passes (AB200_STAGE_TAP_DFLUX) for
    RotatingFrameImpl.  == the host-driven fused loop; strict: bit for bit."""
    import ctypes as C
    from artemis_b200 import capi
    coords = Coordinates.spherical3D
    mesh = make_mesh(coords, 3, bcs=(BoundaryFlag.outflow,) * 6)
    gp, dp = gas_params(coords, "ppm", "hlle"), dust_params(coords, "plm", "hlle", S=2)
    prim, dprim = random_prim(mesh, gp, seed=83), random_prim(mesh, dp, seed=84)
    pm = (0.7, 0.11, -0.07, 0.05, 0.03, 40.0, 2.5)
    sources = [("point_mass",) + pm, ("rotating_frame", 0.6)]
    ncyc = 3
    md1 = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
    md1.gas.prim.set(prim)
    md1.dust.prim.set(dprim)
    d1 = ArtemisDriver(md1, "rk2", mode="fused", nlim=ncyc, sources=sources)
    d1.Initialize()
    d1.Execute()
    md2 = MeshData(mesh, gas=gp, dust=dp, variant=variant, materialize_fluxes=False)
    md2.gas.prim.set(prim)
    md2.dust.prim.set(dprim)
    d2 = ArtemisDriver(md2, "rk2", mode="fused")
    d2.Initialize()
    sd = capi.SourcesDesc()
    sd.point_mass, sd.pm = 1, capi.PointMassDesc(*pm)
    sd.rotating_frame, sd.rf_omega = 1, 0.6
    md2.call("ab200_configure_sources", C.byref(sd))
    md2.set_time_state(d2.dt)
    md2.call("ab200_run_cycles", 1, ncyc, float(np.finfo(np.float64).max))
    ts = md2.time_state()
    assert int(ts[3]) == ncyc
    for f1, f2 in zip(md1.fluids, md2.fluids):
        if variant == "strict":
            assert np.array_equal(f1.u0.get(), f2.u0.get())
            assert np.array_equal(f1.prim.get(), f2.prim.get())
        else:
            assert zone_rel_err(f2.u0.get(), f1.u0.get(), f1.fp, "cons") <= 1e-12
    md1.close()
    md2.close()


def test_rotating_frame_without_mass_fluxes_is_an_error():
    """ab200_rotating_frame after a stage that did not keep its mass fluxes -> AB200_ESTATE."""
    from artemis_b200 import capi
    coords = Coordinates.cylindrical
    mesh = make_mesh(coords, 3, bcs=(BoundaryFlag.outflow,) * 6)
    gp = gas_params(coords, "plm", "hlle")
    md = MeshData(mesh, gas=gp, materialize_fluxes=False)
    md.gas.prim.set(random_prim(mesh, gp, seed=85))
    md.call("ab200_prim_to_cons")
    with pytest.raises(capi.AB200Error, match="no mass fluxes"):
        md.call("ab200_rotating_frame", 1e-3, 0.5)
    md.close()


@pytest.mark.parametrize("coords", [Coordinates.cartesian, Coordinates.spherical3D,
                                    Coordinates.cylindrical])
def test_history_volume_integrals(coords):
    """ab200_history_volume_integrals == sum(u * Volume) over interior zones (history.hpp:24-95)
    with the oracle's cell volumes, for every conserved pack entry of gas and dust."""
    import ctypes as C
    from oracle import oracle_py
    bcs = (BoundaryFlag.outflow,) * 6 if coords != Coordinates.cartesian else None
    mesh = make_mesh(coords, 3, bcs=bcs)
    gp, dp = gas_params(coords, "ppm", "hllc", S=2), dust_params(coords, "plm", "hlle", S=2)
    md = MeshData(mesh, gas=gp, dust=dp, materialize_fluxes=False)
    md.gas.prim.set(random_prim(mesh, gp, seed=91))
    md.dust.prim.set(random_prim(mesh, dp, seed=92))
    md.call("ab200_prim_to_cons")
    L = oracle_py.lib()
    vol = np.zeros((mesh.nb, mesh.nk, mesh.nj, mesh.ni))
    out = np.zeros(32)
    DP = C.POINTER(C.c_double)
    for b in range(mesh.nb):
        xm, dx = np.ascontiguousarray(mesh.blk_xmin[b]), np.ascontiguousarray(mesh.blk_dx[b])
        for k in range(mesh.ks, mesh.ke + 1):
            for j in range(mesh.js, mesh.je + 1):
                for i in range(mesh.is_, mesh.ie + 1):
                    L.ao_geom_cell(int(coords), xm.ctypes.data_as(DP), dx.ctypes.data_as(DP), k, j, i,
                                   out.ctypes.data_as(DP))
                    vol[b, k, j, i] = out[6]
    for ff in md.fluids:
        nv = ff.fp.nvar
        got = np.zeros(nv)
        md.call("ab200_history_volume_integrals", int(ff.fp.fluid_type), got.ctypes.data_as(DP), nv)
        u0 = ff.u0.get()
        want = np.array([np.sum(u0[:, v] * vol) for v in range(nv)])
        scale = np.array([np.sum(np.abs(u0[:, v]) * vol) for v in range(nv)])
        assert np.max(np.abs(got - want) / scale) <= 1e-13
    md.close()
