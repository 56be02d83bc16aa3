"""Device-resident MeshData partition bound to libartemis_b200 through the C ABI.

Plays the role of Parthenon's ``MeshData<Real>`` + ``SparsePack`` for the hot path
(P:interface/mesh_data.hpp, P:interface/sparse_pack.hpp): it owns (or borrows) the device
arrays in the MeshBlockPack layout, builds the (block, pack-index) pointer tables the glue TU
would build from ``pack_h_`` and binds them with ``ab200_bind_pack``.  u0 and u1 share the
OneCopy fields (prim, flux, face velocity) and duplicate only the conserved fields
(src/artemis_driver.cpp:133-138).
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .enums import Fluid
from .enums import BoundaryFlag
from .mesh import UniformMesh
from .params import FluidParams

_DP = C.POINTER(C.c_double)


class DeviceArray:
    """[nb][nvar][nk][nj][ni] slab in device memory (library- or torch-allocated)."""

    def __init__(self, md, shape, torch_tensor=None):
        self.md = md
        self.shape = tuple(shape)
        self.nbytes = int(np.prod(shape)) * 8
        self.tensor = torch_tensor
        if torch_tensor is not None:
            self.ptr = int(torch_tensor.data_ptr())
        else:
            p = C.c_void_p()
            capi.check(md.L, md.L.ab200_malloc(md.ctx, C.byref(p), self.nbytes), "ab200_malloc")
            self.ptr = p.value
            self.zero()

    def zero(self):
        self.set(np.zeros(self.shape))

    def set(self, a):
        a = np.ascontiguousarray(a, dtype=np.float64)
        assert a.shape == self.shape, (a.shape, self.shape)
        capi.check(self.md.L, self.md.L.ab200_memcpy_h2d(self.md.ctx, self.ptr, a.ctypes.data,
                                                          self.nbytes), "h2d")

    def get(self):
        out = np.empty(self.shape)
        capi.check(self.md.L, self.md.L.ab200_memcpy_d2h(self.md.ctx, out.ctypes.data, self.ptr,
                                                          self.nbytes), "d2h")
        return out

    def table(self):
        """Pointer table [nb*nvar] of 3-D sub-array base addresses."""
        nb, nv = self.shape[0], self.shape[1]
        stride = int(np.prod(self.shape[2:])) * 8
        arr = (C.c_void_p * (nb * nv))()
        for e in range(nb * nv):
            arr[e] = self.ptr + e * stride
        return arr

    def free(self):
        if self.tensor is None and self.ptr:
            self.md.L.ab200_free(self.md.ctx, self.ptr)
        self.ptr = 0


class FluidFields:
    def __init__(self, md, fp: FluidParams, materialize_fluxes, alloc):
        m = md.mesh
        self.fp = fp
        S, nv = fp.nspecies, fp.nvar
        gas = fp.fluid_type == Fluid.gas
        self.prim = alloc(m.shape(nv))
        self.u0 = alloc(m.shape(nv))
        self.u1 = alloc(m.shape(nv))
        self.flux = [None] * 3
        self.pflux = [None] * 3
        self.vface = [None] * 3
        if materialize_fluxes:
            for d in range(m.ndim):
                self.flux[d] = alloc(m.shape(nv))
                if gas:
                    self.pflux[d] = alloc(m.shape(S))
                    self.vface[d] = alloc(m.face_shape(S))

    def arrays(self):
        out = [self.prim, self.u0, self.u1]
        for lst in (self.flux, self.pflux, self.vface):
            out += [a for a in lst if a is not None]
        return out


class MeshData:
    def __init__(self, mesh: UniformMesh, gas: FluidParams | None = None,
                 dust: FluidParams | None = None, device: int = 0, variant: str | None = None,
                 stream: int | None = None, materialize_fluxes: bool = True,
                 use_torch: bool = False, bcs=None, shear_bc=None):
        self.mesh = mesh
        self.L = capi.load(variant)
        self.ctx = C.c_void_p()
        capi.check(self.L, self.L.ab200_create(C.byref(self.ctx), device, stream), "ab200_create")
        self.device = device
        self.stream = int(stream) if stream else 0   # 0 = the legacy default stream
        self._torch = None
        if use_torch:
            import torch
            self._torch = torch
        g = capi.GridDesc(int(mesh.coords), mesh.ndim, mesh.nghost, mesh.nb, mesh.ni, mesh.nj,
                          mesh.nk, mesh.is_, mesh.ie, mesh.js, mesh.je, mesh.ks, mesh.ke, mesh.fni,
                          mesh.fnj, mesh.fnk, mesh.blk_xmin.ctypes.data_as(_DP),
                          mesh.blk_dx.ctypes.data_as(_DP))
        capi.check(self.L, self.L.ab200_set_grid(self.ctx, C.byref(g)), "ab200_set_grid")
        self.gas = self._bind(gas, materialize_fluxes) if gas is not None else None
        self.dust = self._bind(dust, materialize_fluxes) if dust is not None else None
        self.fluids = [f for f in (self.gas, self.dust) if f is not None]
        bc = np.array([int(v) for v in (bcs if bcs is not None else mesh.bcs)], dtype=np.int32)
        self.bc = bc
        # State-dependent user conditions (strat.hpp `extrap` / `inflow`): the library's fused
        # ghost fill knows position-independent faces only, so every physical face of such a
        # mesh is the caller's (AB200_BC_NONE in the topology) and AddBoundaryExchangeTasks
        # applies them per block through ab200_block_bcs, in Parthenon's x1 -> x2 -> x3 order.
        self.user_bcs = bool((bc >= int(BoundaryFlag.extrap)).any())
        self._phys_bcs = None
        topo = bc
        if self.user_bcs:
            if (bc == int(BoundaryFlag.fixed)).any():
                raise ValueError("`ic` faces cannot be combined with extrap / inflow faces")
            topo = np.where(bc == int(BoundaryFlag.periodic), 0, 3).astype(np.int32)
        if shear_bc is not None:     # StratParams q, Om0 of the `inflow` condition
            self.call("ab200_set_shear_bc_params", float(shear_bc[0]), float(shear_bc[1]))
        capi.check(self.L, self.L.ab200_set_topology(
            self.ctx, *[int(v) for v in mesh.lattice_n], topo.ctypes.data_as(C.POINTER(C.c_int))),
            "ab200_set_topology")
        if self.user_bcs and not hasattr(mesh, "leaves"):
            self._phys_bcs = self._physical_bc_list(bc)

    def _physical_bc_list(self, bc):
        """ab200_block_bc_desc of every (fluid, block on the lattice boundary, non-periodic face)
        of a uniform mesh -- what ApplyBoundaryConditionsOnCoarseOrFineMD walks"""
        from .multilevel import fill_ghost_ranges
        nbd = [int(v) for v in self.mesh.lattice_n]
        out = []
        for ff in self.fluids:
            fl = int(ff.fp.fluid_type)
            for b in range(self.mesh.nb):
                lb = (b % nbd[0], (b // nbd[0]) % nbd[1], b // (nbd[0] * nbd[1]))
                for face in range(2 * self.mesh.ndim):
                    d, outer = face // 2, face % 2
                    t = int(bc[face])
                    if t == int(BoundaryFlag.periodic) or lb[d] != (nbd[d] - 1 if outer else 0):
                        continue
                    if t >= int(BoundaryFlag.extrap):
                        out.append(capi.BlockBcDesc(fl, b, 0, ff.fp.nvar, face, t, None))
                    else:
                        out += [capi.BlockBcDesc(fl, b, v0, nc, face, t, None)
                                for v0, nc in fill_ghost_ranges(ff.fp)]
        return (capi.BlockBcDesc * len(out))(*out) if out else None

    def apply_physical_bcs(self):
        """the per-block boundary list of a mesh with user conditions (ab200_block_bcs)"""
        if self._phys_bcs is not None:
            self.call("ab200_block_bcs", self._phys_bcs, len(self._phys_bcs))

    # ---- allocation / binding ----------------------------------------------------------
    def _alloc(self, shape):
        if self._torch is not None:
            t = self._torch.zeros(shape, dtype=self._torch.float64,
                                  device=f"cuda:{self.device}")
            return DeviceArray(self, shape, torch_tensor=t)
        return DeviceArray(self, shape)

    def _bind(self, fp: FluidParams, materialize_fluxes):
        ff = FluidFields(self, fp, materialize_fluxes, self._alloc)
        fd = capi.FluidDesc(int(fp.fluid_type), fp.nspecies, int(fp.recon), int(fp.rsolver),
                            fp.gm1, fp.dfloor, fp.siefloor, fp.de_switch, fp.cfl)
        pk = capi.PackDesc()
        keep = []

        def tab(a):
            if a is None:
                return None
            t = a.table()
            keep.append(t)
            return C.cast(t, C.POINTER(C.c_void_p))

        pk.prim, pk.cons0, pk.cons1 = tab(ff.prim), tab(ff.u0), tab(ff.u1)
        for d in range(3):
            pk.flux[d] = tab(ff.flux[d])
            pk.pflux[d] = tab(ff.pflux[d])
            pk.vface[d] = tab(ff.vface[d])
        capi.check(self.L, self.L.ab200_bind_pack(self.ctx, C.byref(fd), C.byref(pk)),
                   "ab200_bind_pack")
        return ff

    def fluid(self, which):
        return self.gas if int(which) == int(Fluid.gas) else self.dust

    def call(self, name, *args):
        capi.check(self.L, getattr(self.L, name)(self.ctx, *args), name)

    def set_stage_path(self, path: int | str):
        """ab200_set_stage_path: 'auto' (0), 'three_pass' (1), 'single_pass' (2) or
        'role_split' (3)."""
        code = {"auto": 0, "three_pass": 1, "single_pass": 2, "role_split": 3}.get(path, path)
        self.call("ab200_set_stage_path", int(code))

    def stage_path(self, which=Fluid.gas) -> str:
        """What ab200_fused_stage runs for a fluid: 'three_pass', 'single_pass', 'role_split'."""
        out = C.c_int(0)
        self.call("ab200_get_stage_path", int(which), C.byref(out))
        return {1: "three_pass", 2: "single_pass", 3: "role_split"}[out.value]

    def synchronize(self):
        self.call("ab200_synchronize")

    def launch_count(self):
        return int(self.L.ab200_launch_count(self.ctx))

    def time_state(self):
        out = np.zeros(4)
        self.call("ab200_read_time_state", out.ctypes.data_as(_DP))
        return out

    def set_time_state(self, dt, new_dt=0.0, time=0.0, ncycle=0):
        a = np.array([dt, new_dt, time, float(ncycle)])
        self.call("ab200_write_time_state", a.ctypes.data_as(_DP))

    def close(self):
        if self.ctx:
            for ff in self.fluids:
                for a in ff.arrays():
                    a.free()
            self.L.ab200_destroy(self.ctx)
            self.ctx = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
