"""ctypes wrapper + stage driver for the CPU oracle.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (artemis_b200/) never imports it.

OracleSim restates, independently of the product's host mirror, the per-stage task order of
ArtemisDriver<GEOM>::StepTasks (src/artemis_driver.cpp:145-270) and Parthenon's
EvolutionDriver dt logic (P:driver/driver.cpp:210-269) on numpy arrays.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
import sys

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_build", "libartemis_oracle.so")
_DP = C.POINTER(C.c_double)
_IP = C.POINTER(C.c_int)


class _Grid(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("geom", "ndim", "ng", "nb", "ni", "nj", "nk", "is_", "ie", "js", "je", "ks",
                 "ke", "fni", "fnj", "fnk")] + [("xmin", _DP), ("dx", _DP)]


class _Fluid(C.Structure):
    _fields_ = [("fluid", C.c_int), ("nspecies", C.c_int), ("recon", C.c_int),
                ("riemann", C.c_int), ("gm1", C.c_double), ("dfloor", C.c_double),
                ("siefloor", C.c_double), ("de_switch", C.c_double), ("cfl", C.c_double)]


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("artemis_oracle.c", "artemis_oracle.h")]
    if (not force and os.path.exists(_LIB)
            and all(os.path.getmtime(_LIB) >= os.path.getmtime(s) for s in src)):
        return _LIB
    subprocess.check_call(["make", "-C", _HERE, "CC=gcc"], stdout=subprocess.DEVNULL)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.ao_estimate_dt.restype = C.c_double
        _lib.ao_num_threads.restype = C.c_int
    return _lib


def _p(a):
    if a is None:
        return None
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data_as(_DP)


def make_grid(mesh):
    g = _Grid(int(mesh.coords), mesh.ndim, mesh.nghost, mesh.nb, mesh.ni, mesh.nj, mesh.nk,
              mesh.is_, mesh.ie, mesh.js, mesh.je, mesh.ks, mesh.ke, mesh.fni, mesh.fnj,
              mesh.fnk, _p(mesh.blk_xmin), _p(mesh.blk_dx))
    g._keep = (mesh.blk_xmin, mesh.blk_dx)
    return g


def make_fluid(fp):
    return _Fluid(int(fp.fluid_type), fp.nspecies, int(fp.recon), int(fp.rsolver), fp.gm1,
                  fp.dfloor, fp.siefloor, fp.de_switch, fp.cfl)


class FluidState:
    """All arrays of one fluid in the MeshBlockPack layout (see artemis_oracle.h)."""

    def __init__(self, mesh, fp, with_flux=True):
        self.fp = fp
        S = fp.nspecies
        gas = int(fp.fluid_type) == 0
        nv = fp.nvar
        self.prim = np.zeros(mesh.shape(nv))
        self.u0 = np.zeros(mesh.shape(nv))
        self.u1 = np.zeros(mesh.shape(nv))
        self.flux = [np.zeros(mesh.shape(nv)) for _ in range(3)] if with_flux else None
        self.pflux = [np.zeros(mesh.shape(S)) for _ in range(3)] if gas and with_flux else [None] * 3
        self.vface = [np.zeros(mesh.face_shape(S)) for _ in range(3)] if gas and with_flux else [None] * 3
        # FillGhost variables: gas prim rho,v,sie (pressure is NOT exchanged:
        # src/gas/gas.cpp:243-270); dust prim rho,v (src/dust/dust.cpp:200-212)
        if gas:
            self.ghost_vars = list(range(0, 4 * S)) + list(range(5 * S, 6 * S))
        else:
            self.ghost_vars = list(range(0, 4 * S))
        self.vec_dir = [((v - S) % 3 + 1) if S <= v < 4 * S else 0 for v in self.ghost_vars]


class Drag(C.Structure):
    """ao_drag (artemis_oracle.h): <drag>, <dust/stopping_time>, <gas|dust/damping> parameters."""
    _fields_ = [("coupling", C.c_int), ("model", C.c_int), ("tau", C.c_double * 16),
                ("scale", C.c_double), ("grain_density", C.c_double), ("sizes", C.c_double * 16),
                ("g_ix", C.c_double * 3), ("g_ox", C.c_double * 3), ("g_irate", C.c_double * 3),
                ("g_orate", C.c_double * 3), ("g_damp_to_visc", C.c_int),
                ("d_ix", C.c_double * 3), ("d_ox", C.c_double * 3), ("d_irate", C.c_double * 3),
                ("d_orate", C.c_double * 3), ("xmin", C.c_double * 3), ("xmax", C.c_double * 3)]


def make_drag(mesh, coupling="simple_dust", model="constant", tau=(), scale=1.0,
              grain_density=1.0, sizes=(), gas_damping=None, dust_damping=None,
              damp_to_visc=False):
    """gas_damping / dust_damping: dict(inner=(x1,x2,x3), outer=(...), inner_rate=(...),
    outer_rate=(...)) with None entries = the deck defaults (-Big / +Big bounds, rate 0)."""
    big = float(np.finfo(np.float64).max)
    d = Drag()
    d.coupling = {"simple_dust": 0, "self": 1}[coupling]
    d.model = {"constant": 0, "stokes": 1}[model]
    d.scale, d.grain_density = scale, grain_density
    for n, v in enumerate(tau):
        d.tau[n] = float(v)
    for n, v in enumerate(sizes):
        d.sizes[n] = float(v)
    for pre, dm in (("g", gas_damping), ("d", dust_damping)):
        dm = dm or {}
        for k in range(3):
            getattr(d, pre + "_ix")[k] = float((dm.get("inner") or (-big,) * 3)[k])
            getattr(d, pre + "_ox")[k] = float((dm.get("outer") or (big,) * 3)[k])
            getattr(d, pre + "_irate")[k] = float((dm.get("inner_rate") or (0.0,) * 3)[k])
            getattr(d, pre + "_orate")[k] = float((dm.get("outer_rate") or (0.0,) * 3)[k])
    d.g_damp_to_visc = int(damp_to_visc)
    for k in range(3):
        d.xmin[k], d.xmax[k] = float(mesh.xmin[k]), float(mesh.xmax[k])
    return d


class Diffusion(C.Structure):
    """ao_diffusion (artemis_oracle.h): viscosity / conduction parameters of the gas."""
    _fields_ = [("visc_type", C.c_int), ("visc_avg", C.c_int), ("nu", C.c_double),
                ("eta", C.c_double), ("r0", C.c_double), ("r_exp", C.c_double),
                ("alpha", C.c_double), ("omega0", C.c_double), ("cond_type", C.c_int),
                ("cond_avg", C.c_int), ("cond", C.c_double), ("kappa", C.c_double),
                ("temp_exp", C.c_double), ("rho_exp", C.c_double), ("rho_ref", C.c_double),
                ("t_ref", C.c_double), ("cv", C.c_double)]


def make_diffusion(visc=None, cond=None, cv=1.0, r0=1.0, gm=1.0):
    """visc: ("constant"|"powerlaw", nu[, r_exp[, eta_bulk]]) | ("alpha", alpha[, eta_bulk]);
    cond: ("conductivity", cond[, temp_exp[, rho_exp]]) | ("diffusivity", kappa[, temp_exp[,
    rho_exp]]); optional trailing "harmonic" selects the face average (diffusion_coeff.hpp:84-140)."""
    d = Diffusion()
    d.cv, d.r0, d.rho_ref, d.t_ref = cv, r0, 1.0, 1.0
    if visc is not None:
        v = list(visc)
        if v[-1] in ("harmonic", "arithmetic"):
            d.visc_avg = int(v.pop() == "harmonic")
        if v[0] == "alpha":
            d.visc_type, d.alpha = 2, float(v[1])
            d.eta = float(v[2]) if len(v) > 2 else 0.0
            d.omega0 = float(np.sqrt(gm / (r0 * r0 * r0)))
        else:
            d.visc_type, d.nu = 1, float(v[1])
            d.r_exp = float(v[2]) if len(v) > 2 else 0.0
            d.eta = float(v[3]) if len(v) > 3 else 0.0
    if cond is not None:
        c = list(cond)
        if c[-1] in ("harmonic", "arithmetic"):
            d.cond_avg = int(c.pop() == "harmonic")
        d.cond_type = 1 if c[0] == "conductivity" else 2
        if d.cond_type == 1:
            d.cond = float(c[1])
        else:
            d.kappa = float(c[1])
        d.temp_exp = float(c[2]) if len(c) > 2 else 0.0
        d.rho_exp = float(c[3]) if len(c) > 3 else 0.0
    return d


class OracleSim:
    """Mini driver: same task order as the reference, one numpy-backed MeshData."""

    def __init__(self, mesh, gas=None, dust=None, integrator="rk2", omf=0.0):
        from artemis_b200.enums import INTEGRATORS  # plain data table, no product code path
        self.mesh = mesh
        self.g = make_grid(mesh)
        self.L = lib()
        self.fluids = []
        self.gas = FluidState(mesh, gas) if gas is not None else None
        self.dust = FluidState(mesh, dust) if dust is not None else None
        self.fluids = [f for f in (self.gas, self.dust) if f is not None]
        self.integrator = integrator
        self.stages = INTEGRATORS[integrator]
        self.omf = omf
        self.time = 0.0
        self.ncycle = 0
        self.dt = np.finfo(np.float64).max
        self.tlim = np.inf
        self.nlim = -1
        # pointwise source terms between FluxSource and SetAuxillaryFields
        # (src/artemis_driver.cpp:217-248): list of ("gravity", gx1, gx2, gx3) |
        # ("shearing_box", om0, qshear) | ("drag", [tau per dust species]) |
        # ("point_mass", gm, x, y, z, soft, sink_rate, sink) | ("rotating_frame", om0)
        self.sources = []
        # gas diffusion (ao_diffusion, make_diffusion()); None = physics/viscosity and
        # physics/conduction off
        self.diffusion = None
        self.dflx = None
        # StratParams q, Om0 of the shearing-box `inflow` user condition (strat.hpp:36-44)
        self.shear_bc = (0.0, 0.0)

    def _diff_lib(self):
        return self.L, "ao"

    def DiffusionFlux(self):
        """Gas::ZeroDiffusionFlux -> ViscousFlux -> ThermalFlux (artemis_driver.cpp:188-196)."""
        fs = self.gas
        if self.dflx is None:
            m = self.mesh
            self.dflx = [np.zeros((m.nb, 4 * fs.fp.nspecies, m.fnk, m.fnj, m.fni)) for _ in range(3)]
        L, pre = self._diff_lib()
        f = make_fluid(fs.fp)
        getattr(L, pre + "_diffusion_flux")(C.byref(self.g), C.byref(f), _p(fs.prim),
                                            C.byref(self.diffusion), *[_p(a) for a in self.dflx])

    def DiffusionUpdate(self, dt):
        """Gas::DiffusionUpdate (artemis_driver.cpp:217-221)."""
        fs = self.gas
        L, pre = self._diff_lib()
        f = make_fluid(fs.fp)
        getattr(L, pre + "_diffusion_update")(C.byref(self.g), C.byref(f), _p(fs.prim), _p(fs.u0),
                                              C.byref(self.diffusion), *[_p(a) for a in self.dflx],
                                              C.c_double(dt))

    def DiffusionTimestep(self):
        fs = self.gas
        L, pre = self._diff_lib()
        f = make_fluid(fs.fp)
        fn = getattr(L, pre + "_diffusion_dt")
        fn.restype = C.c_double
        return fs.fp.cfl * fn(C.byref(self.g), C.byref(f), _p(fs.prim), C.byref(self.diffusion))

    def _both(self):
        fg = make_fluid(self.gas.fp) if self.gas is not None else None
        fd = make_fluid(self.dust.fp) if self.dust is not None else None
        null = C.POINTER(C.c_double)()
        return (fg, fd,
                (C.byref(fg) if fg is not None else None, _p(self.gas.prim) if fg is not None else null,
                 _p(self.gas.u0) if fg is not None else null,
                 C.byref(fd) if fd is not None else None, _p(self.dust.prim) if fd is not None else null,
                 _p(self.dust.u0) if fd is not None else null))

    def _src_lib(self):
        return self.L, "ao"

    def ApplySources(self, dt):
        """ExternalGravity / RotatingFrameForce / DragSource in the reference's task order
        (src/artemis_driver.cpp:222-243: gravity, rotating frame, drag)."""
        if not self.sources:
            return
        L, pre = self._src_lib()
        fg, fd, args = self._both()
        order = {"gravity": 0, "point_mass": 0, "shearing_box": 1, "rotating_frame": 1, "drag": 2,
                 "drag_model": 2}
        for src in sorted(self.sources, key=lambda t: order[t[0]]):
            if src[0] == "point_mass":
                pm = np.ascontiguousarray(src[1:8], dtype=np.float64)
                getattr(L, pre + "_point_mass_gravity")(C.byref(self.g), *args, C.c_double(dt),
                                                        _p(pm))
                continue
            if src[0] == "rotating_frame":
                null = C.POINTER(C.c_double)()
                fa = []
                for fs in (self.gas, self.dust):
                    fa += ([args[0 if fs is self.gas else 3], _p(fs.u0)] + [_p(a) for a in fs.flux]
                           if fs is not None else [None, null, null, null, null])
                getattr(L, pre + "_rotating_frame")(C.byref(self.g), *fa, C.c_double(dt),
                                                    C.c_double(src[1]))
                continue
            if src[0] == "gravity":
                getattr(L, pre + "_uniform_gravity")(C.byref(self.g), *args, C.c_double(dt),
                                                    *[C.c_double(v) for v in src[1:4]])
            elif src[0] == "shearing_box":
                getattr(L, pre + "_shearing_box")(C.byref(self.g), *args, C.c_double(dt),
                                                 C.c_double(src[1]), C.c_double(src[2]))
            elif src[0] in ("drag", "drag_model"):
                # ("drag", [tau]) = constant stopping times, no damping; ("drag_model", Drag)
                dp = make_drag(self.mesh, tau=src[1]) if src[0] == "drag" else src[1]
                dd = C.byref(self.diffusion) if self.diffusion is not None else None
                getattr(L, pre + "_drag_source")(C.byref(self.g), args[0], args[2], args[3],
                                                 args[5], C.byref(dp), dd, C.c_double(dt))

    # ---- task functions (names follow the reference) ---------------------------------
    def CalculateFluxes(self, fs, pcm):
        f = make_fluid(fs.fp)
        self.L.ao_calculate_fluxes(C.byref(self.g), C.byref(f), int(pcm), _p(fs.prim),
                                   *[_p(a) for a in fs.flux], *[_p(a) for a in fs.pflux],
                                   *[_p(a) for a in fs.vface])

    def ApplyUpdate(self, fs, gam0, gam1, beta_dt):
        self.L.ao_apply_update(C.byref(self.g), fs.fp.nvar, _p(fs.u0), _p(fs.u1),
                               *[_p(a) for a in fs.flux], C.c_double(gam0), C.c_double(gam1),
                               C.c_double(beta_dt))

    def FluxSource(self, fs, dt):
        f = make_fluid(fs.fp)
        self.L.ao_flux_source(C.byref(self.g), C.byref(f), _p(fs.prim), _p(fs.u0),
                              *[_p(a) for a in fs.pflux], *[_p(a) for a in fs.vface],
                              C.c_double(self.omf), C.c_double(dt))

    def SetAuxillaryFields(self, fs):
        f = make_fluid(fs.fp)
        self.L.ao_set_aux(C.byref(self.g), C.byref(f), _p(fs.u0))

    def ConsToPrim(self, fs):
        f = make_fluid(fs.fp)
        self.L.ao_cons_to_prim(C.byref(self.g), C.byref(f), _p(fs.u0), _p(fs.prim))

    def PrimToCons(self, fs):
        f = make_fluid(fs.fp)
        self.L.ao_prim_to_cons(C.byref(self.g), C.byref(f), _p(fs.prim), _p(fs.u0))

    def ExchangeGhosts(self, fs):
        m = self.mesh
        if hasattr(m, "leaves"):     # multilevel mesh: Parthenon's multilevel exchange sequence
            from artemis_b200.multilevel import exchange_plan   # plain index bookkeeping
            from . import multilevel_py
            if getattr(self, "_ml_plan", None) is None:
                self._ml_plan = exchange_plan(m)
                self._ml_coarse = {}
            key = id(fs)
            if key not in self._ml_coarse:
                self._ml_coarse[key] = np.zeros(m.coarse_shape(fs.fp.nvar))
            names = {0: "periodic", 1: "outflow", 2: "reflect", 5: "extrap", 6: "inflow"}
            kinds = [names[int(b)] for b in m.bcs]
            user = dict(fluid=int(fs.fp.fluid_type), S=fs.fp.nspecies, q=self.shear_bc[0],
                        om0=self.shear_bc[1])
            multilevel_py.run_plan(m, self._ml_plan, fs.prim, self._ml_coarse[key], fs.ghost_vars,
                                   fs.vec_dir, kinds, user=user)
            return
        vars_ = np.array(fs.ghost_vars, dtype=np.int32)
        vdir = np.array(fs.vec_dir, dtype=np.int32)
        bc = m.bc_ints()
        if (bc == 4).any() and getattr(fs, "ic", None) is None:
            # user `ic` faces: the problem generator's profile = the state the run starts from
            # (ghost zones included), kept for Disk::DiskBoundaryIC
            fs.ic = fs.prim.copy()
        ic = getattr(fs, "ic", None)
        if (bc >= 5).any():   # shearing-box user conditions (strat.hpp), Parthenon's face order
            self.L.ao_exchange_ghosts_user(
                C.byref(self.g), *[int(v) for v in m.lattice_n], bc.ctypes.data_as(_IP),
                fs.fp.nvar, _p(fs.prim), len(vars_), vars_.ctypes.data_as(_IP),
                vdir.ctypes.data_as(_IP), _p(ic) if ic is not None else None,
                int(fs.fp.fluid_type), fs.fp.nspecies, C.c_double(self.shear_bc[0]),
                C.c_double(self.shear_bc[1]))
            return
        self.L.ao_exchange_ghosts_ic(C.byref(self.g), *[int(v) for v in m.lattice_n],
                                     bc.ctypes.data_as(_IP), fs.fp.nvar, _p(fs.prim), len(vars_),
                                     vars_.ctypes.data_as(_IP), vdir.ctypes.data_as(_IP), 3,
                                     _p(ic) if ic is not None else None)

    def EstimateTimestep(self):
        dts = []
        for fs in self.fluids:
            f = make_fluid(fs.fp)
            dts.append(self.L.ao_estimate_dt(C.byref(self.g), C.byref(f), _p(fs.prim)))
        if self.diffusion is not None:   # cfl * min(hydro, viscous, conductive), gas.cpp:437-467
            dts.append(self.DiffusionTimestep())
        return min(dts)

    # ---- Mesh::Initialize sequence after the pgen (P:mesh/mesh.cpp:783-814) -----------
    def initialize(self):
        for fs in self.fluids:
            self.PrimToCons(fs)      # PostInitialization
            self.ConsToPrim(fs)      # PreCommFillDerived
            self.ExchangeGhosts(fs)  # CommunicateBoundaries + physical BCs
            self.PrimToCons(fs)      # FillDerived
        self.block_dt = self.EstimateTimestep()
        self.SetGlobalTimeStep()

    def SetGlobalTimeStep(self):
        """P:driver/driver.cpp:210-269 (dt_factor 2, no user limits)."""
        big = np.finfo(np.float64).max
        if self.dt < 0.1 * big:
            self.dt *= 2.0
        self.dt = min(self.dt, self.block_dt)
        if self.time < self.tlim and (self.tlim - self.time) < self.dt:
            self.dt = self.tlim - self.time

    def stage(self, s):
        gam0, gam1, beta = self.stages[s]
        bdt = beta * self.dt
        pcm = (s == 0 and self.integrator == "vl2")
        for fs in self.fluids:
            self.CalculateFluxes(fs, pcm)
        if self.diffusion is not None:
            self.DiffusionFlux()
        if hasattr(self.mesh, "leaves") and getattr(self, "flux_correction", True):
            # AddFluxCorrectionTasks (artemis_driver.cpp:198-202): every Metadata::Flux field --
            # the conserved fluxes and the interface pressure (gas.prim.pressure WithFluxes)
            from artemis_b200.multilevel import flux_correction_plan
            from . import multilevel_py
            if getattr(self, "_fc_plan", None) is None:
                self._fc_plan = flux_correction_plan(self.mesh)
            for fs in self.fluids:
                multilevel_py.flux_correct(self.mesh, self._fc_plan, fs.flux)
                if fs.pflux[0] is not None:
                    multilevel_py.flux_correct(self.mesh, self._fc_plan, fs.pflux)
            if self.diffusion is not None and self.dflx is not None:
                # gas.diff.momentum / gas.diff.energy are Metadata::WithFluxes fields too
                # (src/gas/gas.cpp:277-285): the same task corrects them
                multilevel_py.flux_correct(self.mesh, self._fc_plan, self.dflx)
        for fs in self.fluids:
            self.ApplyUpdate(fs, gam0, gam1, bdt)
        for fs in self.fluids:
            self.FluxSource(fs, bdt)
        if self.diffusion is not None:
            self.DiffusionUpdate(bdt)
        self.ApplySources(bdt)
        for fs in self.fluids:
            self.SetAuxillaryFields(fs)
            self.ConsToPrim(fs)
            self.ExchangeGhosts(fs)
            self.PrimToCons(fs)

    def step(self):
        for fs in self.fluids:           # DeepCopyConservedData (artemis_driver.cpp:156-163)
            np.copyto(fs.u1, fs.u0)
        for s in range(len(self.stages)):
            self.stage(s)
        self.block_dt = self.EstimateTimestep()
        self.ncycle += 1
        self.time += self.dt
        self.SetGlobalTimeStep()

    def keep_going(self):
        return (self.time < self.tlim) and (self.nlim < 0 or self.ncycle < self.nlim)

    def run(self):
        while self.keep_going():
            self.step()


if __name__ == "__main__":
    build(force=True)
    print(_LIB, "threads:", lib().ao_num_threads(), file=sys.stderr)


# ---- multilevel operators ----------------------------------------------
class RefineGeom(C.Structure):
    """ao_refine_geom: one MeshBlock's fine array and coarse buffer (artemis_oracle.h)."""
    _fields_ = [(n, C.c_int) for n in
                ("geom", "ndim", "ng", "ni", "nj", "nk", "cni", "cnj", "cnk", "ib_s", "jb_s",
                 "kb_s", "cib_s", "cjb_s", "ckb_s")] + [("xmin", C.c_double * 3),
                                                        ("dx", C.c_double * 3)]


def refine_geom(mesh, b=0):
    """Coarse-buffer geometry of block b: nx/2 interior cells + ng ghosts in every active
    direction (P:mesh/meshblock.cpp:205-228)."""
    ng, nd = mesh.nghost, mesh.ndim
    cn = [mesh.block_nx[d] // 2 + 2 * ng if d < nd else 1 for d in range(3)]
    cs = [ng if d < nd else 0 for d in range(3)]
    r = RefineGeom(int(mesh.coords), nd, ng, mesh.ni, mesh.nj, mesh.nk, cn[0], cn[1], cn[2],
                   mesh.is_, mesh.js, mesh.ks, cs[0], cs[1], cs[2])
    for d in range(3):
        r.xmin[d] = float(mesh.blk_xmin[b, d])
        r.dx[d] = float(mesh.blk_dx[b, d])
    return r


def _box(box):
    return (C.c_int * 6)(*[int(v) for v in box])


def restrict_average(L, r, fine, coarse, box, prefix="ao"):
    """coarse[box] <- volume-weighted average of fine; fine [nvar][nk][nj][ni]."""
    getattr(L, prefix + "_restrict_average")(C.byref(r), fine.shape[0], _p(fine), _p(coarse),
                                             _box(box))


def restrict_average_face(L, r, fine, coarse, box, el, prefix="ao"):
    """the same operator on a flux field living on x1 / x2 / x3 faces (el = 1..3)"""
    getattr(L, prefix + "_restrict_average_face")(C.byref(r), fine.shape[0], _p(fine), _p(coarse),
                                                  _box(box), el)


def prolongate_minmod(L, r, coarse, fine, box, prefix="ao"):
    getattr(L, prefix + "_prolongate_minmod")(C.byref(r), coarse.shape[0], _p(coarse), _p(fine),
                                              _box(box))


def strat_bc(arr, geom, xmin, dx, s, e, fluid, S, face, kind, q=0.0, om0=0.0):
    """ao_strat_bc on one block array [nvar][nk][nj][ni] in place (strat.hpp:154-666);
    kind 5 = extrap (x1 / x3 faces), 6 = inflow (x2 faces)."""
    assert arr.flags["C_CONTIGUOUS"] and arr.dtype == np.float64
    nk, nj, ni = arr.shape[1:]
    D3, I3 = C.c_double * 3, C.c_int * 3
    rc = lib().ao_strat_bc(int(geom), D3(*[float(v) for v in xmin]), D3(*[float(v) for v in dx]),
                           ni, nj, nk, I3(*[int(v) for v in s]), I3(*[int(v) for v in e]),
                           int(fluid), int(S), _p(arr), int(face), int(kind), C.c_double(q),
                           C.c_double(om0))
    if rc != 0:
        raise ValueError(f"strat.hpp registers no condition of kind {kind} on face {face}")
