"""Host mirror of the reference's task functions and stage driver for the hot path.

Same names, argument meaning and error behaviour as the reference:
every task function takes the MeshData it acts on and returns a TaskStatus; failures of the
C ABI surface as ``TaskStatus.fail`` from the task (and raise from the driver, the analogue
of PARTHENON_FAIL).  All compute goes through libartemis_b200 -- there is no CPU path here.

  Gas.CalculateFluxes / Gas.FluxSource / Gas.EstimateTimestepMesh   src/gas/gas.cpp:391-519
  Dust.CalculateFluxes / Dust.FluxSource / Dust.EstimateTimestepMesh src/dust/dust.cpp:238-326
  ArtemisUtils.ApplyUpdate / DeepCopyConservedData   src/utils/integrators/artemis_integrator.hpp
  ArtemisDerived.SetAuxillaryFields / ConsToPrim / PrimToCons        src/derived/fill_derived.cpp
  ArtemisDriver.Step / StepTasks / PostStepTasks                     src/artemis_driver.cpp:101-297
"""
from __future__ import annotations

import ctypes as C
from enum import Enum

import numpy as np

from . import capi
from .enums import INTEGRATORS, Fluid
from .meshdata import MeshData

_DP = C.POINTER(C.c_double)
_BIG = float(np.finfo(np.float64).max)


class TaskStatus(Enum):  # P:basic_types.hpp:59
    complete = 0
    incomplete = 1
    iterate = 2
    fail = 3


def _task(md: MeshData, name, *args) -> TaskStatus:
    rc = getattr(md.L, name)(md.ctx, *args)
    if rc != 0:
        md.last_error = md.L.ab200_last_error().decode()
        return TaskStatus.fail
    return TaskStatus.complete


class LowStorageIntegrator:
    """gam0/gam1/beta tables of P:time_integration/low_storage_integrator.cpp."""

    def __init__(self, name: str):
        if name not in INTEGRATORS:
            raise ValueError(f"integrator {name!r} not supported on the hot path")
        self._name = name
        st = INTEGRATORS[name]
        self.nstages = len(st)
        self.gam0 = [s[0] for s in st]
        self.gam1 = [s[1] for s in st]
        self.beta = [s[2] for s in st]
        self.dt = 0.0

    def GetName(self):
        return self._name


class _FluidPkg:
    fluid = Fluid.gas

    @classmethod
    def CalculateFluxes(cls, md: MeshData, pcm: bool) -> TaskStatus:
        return _task(md, "ab200_calculate_fluxes", int(cls.fluid), int(bool(pcm)))

    @classmethod
    def FluxSource(cls, md: MeshData, dt: float) -> TaskStatus:
        return _task(md, "ab200_flux_source", int(cls.fluid), float(dt))

    @classmethod
    def EstimateTimestepMesh(cls, md: MeshData) -> float:
        out = C.c_double()
        capi.check(md.L, md.L.ab200_estimate_timestep(md.ctx, int(cls.fluid), C.byref(out)),
                   f"{cls.__name__}::EstimateTimestepMesh")
        return out.value


class Gas(_FluidPkg):
    fluid = Fluid.gas


class Dust(_FluidPkg):
    fluid = Fluid.dust


class ArtemisUtils:
    @staticmethod
    def DeepCopyConservedData(md: MeshData) -> TaskStatus:
        return _task(md, "ab200_deep_copy_conserved")

    @staticmethod
    def ApplyUpdate(md: MeshData, stage: int, integrator: LowStorageIntegrator) -> TaskStatus:
        gam0 = integrator.gam0[stage - 1]
        gam1 = integrator.gam1[stage - 1]
        beta_dt = integrator.beta[stage - 1] * integrator.dt
        return _task(md, "ab200_apply_update", gam0, gam1, beta_dt)


class ArtemisDerived:
    @staticmethod
    def SetAuxillaryFields(md: MeshData) -> TaskStatus:
        return _task(md, "ab200_set_auxillary_fields")

    @staticmethod
    def ConsToPrim(md: MeshData) -> TaskStatus:
        return _task(md, "ab200_cons_to_prim")

    @staticmethod
    def PrimToCons(md: MeshData) -> TaskStatus:
        return _task(md, "ab200_prim_to_cons")


def AddBoundaryExchangeTasks(md: MeshData, comm=None) -> TaskStatus:
    """parthenon::AddBoundaryExchangeTasks (P:bvals/comms/boundary_communication.cpp:433-444):
    same-GPU neighbours in one kernel, remote neighbours through `comm`, then physical BCs.
    On a multilevel mesh `comm` is a multilevel.MultilevelExchange, which runs the whole
    sequence (restrict / copy / restrict / coarse BCs / prolongate / fine BCs)."""
    if getattr(comm, "multilevel", False):
        try:
            comm.exchange()
        except capi.AB200Error as e:
            md.last_error = str(e)
            return TaskStatus.fail
        return TaskStatus.complete
    if comm is not None and getattr(md, "user_bcs", False):
        # a face flagged AB200_BC_NONE means "another rank" to the transport and "user condition"
        # to this mirror: the two are not told apart yet, so refuse instead of exchanging wrongly
        md.last_error = ("extrap / inflow user conditions on a mesh split across ranks are not "
                         "supported by the host mirror")
        return TaskStatus.fail
    st = _task(md, "ab200_exchange_ghosts")
    if st != TaskStatus.complete:
        return st
    if comm is not None:
        comm.exchange(md)
    if getattr(md, "user_bcs", False):
        # state-dependent user conditions (strat.hpp extrap / inflow): every physical face of
        # the mesh goes through the per-block list, in Parthenon's x1 -> x2 -> x3 order
        try:
            md.apply_physical_bcs()
        except capi.AB200Error as e:
            md.last_error = str(e)
            return TaskStatus.fail
        return TaskStatus.complete
    return _task(md, "ab200_apply_physical_bcs")


class ArtemisDriver:
    """Stage driver for the hot path (uniform mesh; out-of-scope source terms absent).

    mode = "tasks": one C-ABI call per reference task, in the reference's order.
    mode = "fused": ab200_fused_stage + ghost-zone PrimToCons (the fast path).
    """

    def __init__(self, md: MeshData, integrator: str = "rk2", mode: str = "tasks",
                 tlim: float = np.inf, nlim: int = -1, comm=None, sources=(), diffusion=None,
                 flux_correction: bool = True):
        self.md = md
        # gas diffusion (physics/viscosity, physics/conduction): a capi.DiffusionDesc; the
        # operators run every stage as in src/artemis_driver.cpp:188-196, 217-221
        self.diffusion = diffusion
        if diffusion is not None:
            md.call("ab200_configure_diffusion", C.byref(diffusion))
        # pointwise source terms between FluxSource and SetAuxillaryFields
        # (src/artemis_driver.cpp:217-248): ("gravity", gx1, gx2, gx3) |
        # ("shearing_box", omega, qshear) | ("drag", [tau per dust species])
        self.sources = list(sources)
        self.integrator = LowStorageIntegrator(integrator)
        self.mode = mode
        self.comm = comm
        self.time = 0.0
        self.ncycle = 0
        self.dt = _BIG
        self.tlim = tlim
        self.nlim = nlim
        self.do_gas = md.gas is not None
        self.do_dust = md.dust is not None
        # AddFluxCorrectionTasks on refined meshes (src/artemis_driver.cpp:198-202).  The fused
        # path never stores fluxes, so it cannot correct them: on a refined mesh it is only
        # available as the explicitly non-conservative variant (flux_correction=False).
        self.flux_correction = flux_correction
        if getattr(comm, "multilevel", False) and flux_correction and mode != "tasks":
            raise ValueError("refined mesh: the fused stage path stores no fluxes and cannot apply "
                             "Parthenon's flux correction -- use mode='tasks', or pass "
                             "flux_correction=False for the non-conservative variant")

    @staticmethod
    def _require(st: TaskStatus, md, what):
        if st != TaskStatus.complete:
            raise capi.AB200Error(f"{what}: {getattr(md, 'last_error', '')}")

    # Mesh::Initialize after the pgen (P:mesh/mesh.cpp:783-814)
    def Initialize(self):
        md = self.md
        self._require(ArtemisDerived.PrimToCons(md), md, "PostInitialization")
        self._require(ArtemisDerived.ConsToPrim(md), md, "PreCommFillDerived")
        self._require(AddBoundaryExchangeTasks(md, self.comm), md, "CommunicateBoundaries")
        self._require(ArtemisDerived.PrimToCons(md), md, "FillDerived")
        self.block_dt = self.EstimateTimestep()
        self.SetGlobalTimeStep()

    def EstimateTimestep(self) -> float:
        dts = []
        if self.do_gas:
            dts.append(Gas.EstimateTimestepMesh(self.md))
        if self.do_dust:
            dts.append(Dust.EstimateTimestepMesh(self.md))
        dt = min(dts)
        if self.comm is not None:
            dt = self.comm.allreduce_min(dt)
        return dt

    def SetGlobalTimeStep(self):
        """EvolutionDriver::SetGlobalTimeStep, P:driver/driver.cpp:210-269."""
        if self.dt < 0.1 * _BIG:
            self.dt *= 2.0
        self.dt = min(self.dt, self.block_dt)
        if self.time < self.tlim and (self.tlim - self.time) < self.dt:
            self.dt = self.tlim - self.time

    def ApplySources(self, bdt: float):
        """ExternalGravity -> RotatingFrameForce -> DragSource, the reference's task order
        (src/artemis_driver.cpp:222-243)."""
        md, req = self.md, self._require
        order = {"gravity": 0, "point_mass": 0, "shearing_box": 1, "rotating_frame": 1, "drag": 2,
                 "drag_model": 2}
        for src in sorted(self.sources, key=lambda t: order[t[0]]):
            if src[0] == "point_mass":
                pm = capi.PointMassDesc(*[float(v) for v in src[1:8]])
                req(_task(md, "ab200_point_mass_gravity", float(bdt), C.byref(pm)),
                    md, "Gravity::PointMassGravity")
            elif src[0] == "rotating_frame":
                req(_task(md, "ab200_rotating_frame", float(bdt), float(src[1])),
                    md, "RotatingFrame::RotatingFrameImpl")
            elif src[0] == "gravity":
                req(_task(md, "ab200_uniform_gravity", float(bdt), *[float(v) for v in src[1:4]]),
                    md, "Gravity::UniformGravity")
            elif src[0] == "shearing_box":
                req(_task(md, "ab200_shearing_box", float(bdt), float(src[1]), float(src[2])),
                    md, "RotatingFrame::ShearingBoxImpl")
            elif src[0] == "drag_model":    # ("drag_model", capi.DragDesc)
                req(_task(md, "ab200_drag_source", float(bdt), C.byref(src[1])), md,
                    "Drag::DragSource")
            elif src[0] == "drag":
                tau = np.ascontiguousarray(src[1], dtype=np.float64)
                req(_task(md, "ab200_drag_simple", float(bdt), len(tau), tau.ctypes.data_as(_DP)),
                    md, "Drag::DragSource")
            else:
                raise ValueError(f"unknown source term {src[0]!r}")

    def StepTasks(self):
        md, integ = self.md, self.integrator
        req = self._require
        if self.mode == "tasks":
            req(ArtemisUtils.DeepCopyConservedData(md), md, "DeepCopyConservedData")
        for stage in range(1, integ.nstages + 1):
            bdt = integ.beta[stage - 1] * integ.dt
            do_pcm = (stage == 1) and (integ.GetName() == "vl2")
            if self.mode == "tasks":
                if self.do_gas:
                    req(Gas.CalculateFluxes(md, do_pcm), md, "Gas::CalculateFluxes")
                if self.do_dust:
                    req(Dust.CalculateFluxes(md, do_pcm), md, "Dust::CalculateFluxes")
                if self.diffusion is not None:
                    req(_task(md, "ab200_diffusion_flux"), md, "Gas::ViscousFlux/ThermalFlux")
                if getattr(self.comm, "multilevel", False) and self.flux_correction:
                    self.comm.flux_correct()     # AddFluxCorrectionTasks, artemis_driver.cpp:198-202
                req(ArtemisUtils.ApplyUpdate(md, stage, integ), md, "ApplyUpdate")
                if self.do_gas:
                    req(Gas.FluxSource(md, bdt), md, "Gas::FluxSource")
                if self.do_dust:
                    req(Dust.FluxSource(md, bdt), md, "Dust::FluxSource")
                if self.diffusion is not None:
                    req(_task(md, "ab200_diffusion_update", float(bdt)), md, "Gas::DiffusionUpdate")
                self.ApplySources(bdt)
                req(ArtemisDerived.SetAuxillaryFields(md), md, "SetAuxillaryFields")
                req(ArtemisDerived.ConsToPrim(md), md, "ConsToPrim")
                req(AddBoundaryExchangeTasks(md, self.comm), md, "AddBoundaryExchangeTasks")
                req(ArtemisDerived.PrimToCons(md), md, "PrimToCons")
            else:
                # with source terms the conserved state must exist between the update and C2P:
                # the fused passes stop after FluxSource (AB200_STAGE_DEFER_C2P = 8)
                defer = 8 if (self.sources or self.diffusion is not None) else 0
                if any(src[0] == "rotating_frame" for src in self.sources):
                    defer |= 64   # AB200_STAGE_TAP_DFLUX: the passes keep their mass fluxes
                req(_task(md, "ab200_fused_stage", integ.gam0[stage - 1], integ.gam1[stage - 1],
                          integ.beta[stage - 1], integ.dt, int(do_pcm), int(stage == 1), defer),
                    md, "ab200_fused_stage")
                if defer:
                    if self.diffusion is not None:   # stage-start primitives are still in place
                        req(_task(md, "ab200_diffusion_flux"), md, "Gas::ViscousFlux/ThermalFlux")
                        req(_task(md, "ab200_diffusion_update", float(bdt)), md,
                            "Gas::DiffusionUpdate")
                    self.ApplySources(bdt)
                    req(_task(md, "ab200_finish_stage", 0), md, "ab200_finish_stage")
                req(AddBoundaryExchangeTasks(md, self.comm), md, "AddBoundaryExchangeTasks")
                req(_task(md, "ab200_prim_to_cons_ghosts"), md, "PrimToCons(ghosts)")

    def Step(self):
        self.integrator.dt = self.dt        # PreStepTasks, artemis_driver.cpp:128-130
        self.StepTasks()
        self.block_dt = self.EstimateTimestep()   # PostStepTasks, :277-297
        self.ncycle += 1
        self.time += self.dt
        self.SetGlobalTimeStep()

    # ---- device-resident cycle (fused path; dt, time and ncycle live on the device) ----------
    def BeginDeviceResident(self):
        self.md.set_time_state(self.dt, 0.0, self.time, self.ncycle)
        # conserved ghost zones are dead between fused stages: convert them once at the end
        self.md.call("ab200_set_ghost_cons_lazy", 1)

    def StepDevice(self, tlim: float = _BIG):
        """One cycle with no host round trip: per stage ab200_fused_stage (dt read from the
        device scalar, CFL reduction folded into the last stage) -> ghost fill; then the dt
        all-reduce on the device scalar and ab200_set_global_timestep_device.  Multi-rank:
        single-round remote exchange (pack / NCCL / unpack on a second stream) concurrent with
        the local ghost fill -> finish."""
        md, integ = self.md, self.integrator
        if getattr(md, "user_bcs", False):
            raise capi.AB200Error("the device-resident cycle fills ghost zones with the fused "
                                  "kernel, which has no state-dependent user conditions "
                                  "(extrap / inflow): use Step()")
        for stage in range(1, integ.nstages + 1):
            do_pcm = (stage == 1) and (integ.GetName() == "vl2")
            # DEVICE_DT | PINGPONG (primitives may stay in the alternate set) | REDUCE_DT
            flags = 1 | 4 | (2 if stage == integ.nstages else 0)
            md.call("ab200_fused_stage", integ.gam0[stage - 1], integ.gam1[stage - 1],
                    integ.beta[stage - 1], 0.0, int(do_pcm), int(stage == 1), flags)
            if self.comm is None:
                md.call("ab200_fill_ghosts")
            else:
                # single remote round (faces, rank edges, corners at once) on its own stream,
                # concurrent with the same-GPU ghost fill; both join before the finish pass
                self.comm.begin_direct(md)
                md.call("ab200_fill_ghosts_local")
                self.comm.end_direct(md)
                md.call("ab200_finish_remote_ghosts")
        if self.comm is not None:
            self.comm.allreduce_min_device()
        md.call("ab200_set_global_timestep_device", float(tlim), 1)
        md.call("ab200_sync_prim")  # no-op after an even number of single-pass stages

    def EndDeviceResident(self):
        self.md.call("ab200_set_ghost_cons_lazy", 0)
        self.md.call("ab200_sync_ghost_cons")
        ts = self.md.time_state()
        self.dt, self.time, self.ncycle = float(ts[0]), float(ts[2]), int(ts[3])

    def KeepGoing(self):
        return (self.time < self.tlim) and (self.nlim < 0 or self.ncycle < self.nlim)

    def Execute(self):
        while self.KeepGoing():
            self.Step()
