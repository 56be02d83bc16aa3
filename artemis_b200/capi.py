"""ctypes binding of include/ab200.h (the C ABI of libartemis_b200).

This is the only way the Python host mirror reaches the CUDA path.  There is no fallback: if
the shared library cannot be loaded, or no CUDA device is present, calls raise.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
_DP = C.POINTER(C.c_double)
_PP = C.POINTER(C.c_void_p)


class GridDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in
                ("geom", "ndim", "nghost", "nblocks", "ni", "nj", "nk", "is_", "ie", "js", "je",
                 "ks", "ke", "fni", "fnj", "fnk")] + [("xmin", _DP), ("dx", _DP)]


class FluidDesc(C.Structure):
    _fields_ = [("fluid", C.c_int), ("nspecies", C.c_int), ("recon", C.c_int),
                ("riemann", C.c_int), ("gm1", C.c_double), ("dfloor", C.c_double),
                ("siefloor", C.c_double), ("de_switch", C.c_double), ("cfl", C.c_double)]


class PackDesc(C.Structure):
    _fields_ = [("prim", _PP), ("cons0", _PP), ("cons1", _PP), ("flux", _PP * 3),
                ("pflux", _PP * 3), ("vface", _PP * 3)]


class BndDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("fluid", "block", "var0", "ncomp", "si", "ei", "sj", "ej",
                                       "sk", "ek")] + [("buf", C.c_void_p)]


class RefineDesc(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("fluid", "block", "var0", "nvar", "kind", "cis", "cie",
                                       "cjs", "cje", "cks", "cke")] + [("coarse", C.c_void_p)]


class PointMassDesc(C.Structure):
    """ab200_point_mass_desc"""
    _fields_ = [(n, C.c_double) for n in ("gm", "x", "y", "z", "soft", "sink_rate", "sink")]


class DiffusionDesc(C.Structure):
    """ab200_diffusion_desc"""
    _fields_ = [("visc_type", C.c_int), ("visc_avg", C.c_int), ("nu", C.c_double),
                ("eta_bulk", C.c_double), ("r0", C.c_double), ("r_exp", C.c_double),
                ("alpha", C.c_double), ("omega0", C.c_double), ("cond_type", C.c_int),
                ("cond_avg", C.c_int), ("cond", C.c_double), ("kappa", C.c_double),
                ("temp_exp", C.c_double), ("rho_exp", C.c_double), ("rho_ref", C.c_double),
                ("t_ref", C.c_double), ("cv", C.c_double)]


class BoxDesc(C.Structure):
    """ab200_box_desc"""
    _fields_ = [("fluid", C.c_int), ("ncomp", C.c_int), ("src_block", C.c_int),
                ("src_var0", C.c_int), ("src_coarse", C.c_void_p), ("dst_block", C.c_int),
                ("dst_var0", C.c_int), ("dst_coarse", C.c_void_p)] + [
                    (n, C.c_int) for n in ("ssi", "ssj", "ssk", "dsi", "dsj", "dsk", "ni", "nj", "nk")]


class FluxCorDesc(C.Structure):
    """ab200_fluxcor_desc"""
    _fields_ = [(n, C.c_int) for n in ("fluid", "fine_block", "coarse_block", "dir", "cis", "cie",
                                       "cjs", "cje", "cks", "cke", "dsi", "dsj", "dsk")]


class BlockBcDesc(C.Structure):
    """ab200_block_bc_desc"""
    _fields_ = [(n, C.c_int) for n in ("fluid", "block", "var0", "ncomp", "face", "type")] + [
        ("coarse", C.c_void_p), ("coarse_entries", C.c_void_p)]


class DragDesc(C.Structure):
    """ab200_drag_desc"""
    _fields_ = [("coupling", C.c_int), ("model", C.c_int), ("tau", C.c_double * 16),
                ("scale", C.c_double), ("grain_density", C.c_double), ("sizes", C.c_double * 16),
                ("g_ix", C.c_double * 3), ("g_ox", C.c_double * 3), ("g_irate", C.c_double * 3),
                ("g_orate", C.c_double * 3), ("g_damp_to_visc", C.c_int),
                ("d_ix", C.c_double * 3), ("d_ox", C.c_double * 3), ("d_irate", C.c_double * 3),
                ("d_orate", C.c_double * 3), ("xmin", C.c_double * 3), ("xmax", C.c_double * 3)]


class SourcesDesc(C.Structure):
    """ab200_sources_desc"""
    _fields_ = [("gravity", C.c_int), ("g", C.c_double * 3), ("shearing_box", C.c_int),
                ("omega", C.c_double), ("qshear", C.c_double), ("drag", C.c_int),
                ("ntau", C.c_int), ("tau", C.c_double * 16), ("point_mass", C.c_int),
                ("pm", PointMassDesc), ("rotating_frame", C.c_int), ("rf_omega", C.c_double),
                ("drag_model", C.c_int), ("drag_desc", DragDesc)]


class AB200Error(RuntimeError):
    pass


# every symbol include/ab200.h declares (checked by tests/test_abi.py)
SYMBOLS = [
    "ab200_abi_version", "ab200_last_error", "ab200_device_count", "ab200_create",
    "ab200_destroy", "ab200_synchronize", "ab200_set_grid", "ab200_bind_pack", "ab200_unbind",
    "ab200_set_rotating_frame", "ab200_set_stage_path", "ab200_get_stage_path", "ab200_calculate_fluxes", "ab200_apply_update",
    "ab200_flux_source", "ab200_set_auxillary_fields", "ab200_cons_to_prim",
    "ab200_prim_to_cons", "ab200_deep_copy_conserved", "ab200_estimate_timestep",
    "ab200_fused_stage", "ab200_sync_prim", "ab200_prim_to_cons_ghosts", "ab200_set_ghost_cons_lazy",
    "ab200_sync_ghost_cons", "ab200_estimate_timestep_device",
    "ab200_set_global_timestep_device", "ab200_dt_device", "ab200_read_time_state",
    "ab200_write_time_state", "ab200_halo_pack", "ab200_halo_unpack", "ab200_set_halo_stream", "ab200_coarse_shape", "ab200_restrict", "ab200_prolongate",
    "ab200_set_topology",
    "ab200_exchange_ghosts", "ab200_apply_physical_bcs", "ab200_fill_ghosts",
    "ab200_fill_ghosts_local", "ab200_finish_remote_ghosts", "ab200_cycles_host",
    "ab200_run_cycles", "ab200_malloc", "ab200_free", "ab200_memcpy_h2d", "ab200_memcpy_d2h",
    "ab200_launch_count", "ab200_timer_begin", "ab200_timer_end",
    "ab200_history_volume_integrals", "ab200_configure_sources", "ab200_finish_stage", "ab200_uniform_gravity", "ab200_shearing_box", "ab200_drag_simple",
    "ab200_point_mass_gravity", "ab200_rotating_frame", "ab200_drag_source",
    "ab200_box_copy", "ab200_block_bcs", "ab200_set_shear_bc_params", "ab200_flux_correct", "ab200_set_host_transfer", "ab200_set_graph_replay", "ab200_graph_replay_count",
    "ab200_configure_diffusion", "ab200_diffusion_flux", "ab200_diffusion_update",
    "ab200_diffusion_timestep", "ab200_diffusion_flux_array",
    "ab200_comm_unique_id", "ab200_comm_init", "ab200_comm_destroy", "ab200_comm_set_layout",
    "ab200_comm_bytes_per_exchange", "ab200_comm_is_direct", "ab200_comm_exchange_begin", "ab200_comm_exchange_end",
    "ab200_allreduce_min", "ab200_run_cycles_mr", "ab200_comm_plan_direct", "ab200_comm_plan_free",
]

_libs: dict[str, C.CDLL] = {}


def lib_path(variant: str = "fast") -> str:
    name = "libartemis_b200.so" if variant == "fast" else f"libartemis_b200_{variant}.so"
    return os.path.join(_HERE, "lib", name)


def load(variant: str | None = None) -> C.CDLL:
    variant = variant or os.environ.get("AB200_VARIANT", "fast")
    if variant in _libs:
        return _libs[variant]
    path = lib_path(variant)
    if not os.path.exists(path):
        raise AB200Error(
            f"{path} not found: build it with `python -m artemis_b200.build` "
            "(libartemis_b200 is the only compute path; there is no CPU fallback)")
    L = C.CDLL(path)
    L.ab200_last_error.restype = C.c_char_p
    L.ab200_dt_device.restype = C.c_void_p
    L.ab200_dt_device.argtypes = [C.c_void_p]
    L.ab200_launch_count.restype = C.c_longlong
    L.ab200_launch_count.argtypes = [C.c_void_p]
    L.ab200_comm_bytes_per_exchange.restype = C.c_longlong
    L.ab200_comm_bytes_per_exchange.argtypes = [C.c_void_p]
    L.ab200_comm_plan_free.restype = None
    L.ab200_comm_plan_free.argtypes = [C.POINTER(C.c_longlong)]
    vp, i, d = C.c_void_p, C.c_int, C.c_double
    sig = {
        "ab200_create": [C.POINTER(vp), i, vp], "ab200_destroy": [vp], "ab200_synchronize": [vp],
        "ab200_set_grid": [vp, C.POINTER(GridDesc)],
        "ab200_bind_pack": [vp, C.POINTER(FluidDesc), C.POINTER(PackDesc)],
        "ab200_unbind": [vp, i], "ab200_set_rotating_frame": [vp, d],
        "ab200_set_stage_path": [vp, i], "ab200_get_stage_path": [vp, i, C.POINTER(C.c_int)],
        "ab200_calculate_fluxes": [vp, i, i], "ab200_apply_update": [vp, d, d, d],
        "ab200_flux_source": [vp, i, d], "ab200_set_auxillary_fields": [vp],
        "ab200_cons_to_prim": [vp], "ab200_prim_to_cons": [vp],
        "ab200_deep_copy_conserved": [vp], "ab200_estimate_timestep": [vp, i, _DP],
        "ab200_fused_stage": [vp, d, d, d, d, i, i, i], "ab200_prim_to_cons_ghosts": [vp], "ab200_sync_prim": [vp],
        "ab200_set_ghost_cons_lazy": [vp, i], "ab200_sync_ghost_cons": [vp],
        "ab200_estimate_timestep_device": [vp],
        "ab200_set_global_timestep_device": [vp, d, i],
        "ab200_read_time_state": [vp, _DP], "ab200_write_time_state": [vp, _DP],
        "ab200_halo_pack": [vp, C.POINTER(BndDesc), i],
        "ab200_halo_unpack": [vp, C.POINTER(BndDesc), i],
        "ab200_set_halo_stream": [vp, vp],
        "ab200_coarse_shape": [vp, C.POINTER(C.c_int)],
        "ab200_restrict": [vp, C.POINTER(RefineDesc), i],
        "ab200_prolongate": [vp, C.POINTER(RefineDesc), i],
        "ab200_set_topology": [vp, i, i, i, C.POINTER(C.c_int)],
        "ab200_exchange_ghosts": [vp], "ab200_apply_physical_bcs": [vp],
        "ab200_fill_ghosts": [vp], "ab200_fill_ghosts_local": [vp],
        "ab200_finish_remote_ghosts": [vp],
        "ab200_cycles_host": [vp, i, i, _DP, _DP, _DP, _DP, _DP],
        "ab200_run_cycles": [vp, i, i, d],
        "ab200_malloc": [vp, C.POINTER(vp), C.c_size_t], "ab200_free": [vp, vp],
        "ab200_memcpy_h2d": [vp, vp, vp, C.c_size_t], "ab200_memcpy_d2h": [vp, vp, vp, C.c_size_t],
        "ab200_timer_begin": [vp], "ab200_timer_end": [vp, C.POINTER(C.c_float)],
        "ab200_history_volume_integrals": [vp, i, _DP, i], "ab200_configure_sources": [vp, vp], "ab200_finish_stage": [vp, i], "ab200_uniform_gravity": [vp, d, d, d, d],
        "ab200_shearing_box": [vp, d, d, d], "ab200_drag_simple": [vp, d, i, _DP],
        "ab200_point_mass_gravity": [vp, d, C.POINTER(PointMassDesc)],
        "ab200_rotating_frame": [vp, d, d], "ab200_drag_source": [vp, d, C.POINTER(DragDesc)],
        "ab200_flux_correct": [vp, C.POINTER(FluxCorDesc), i],
        "ab200_set_host_transfer": [vp, i], "ab200_set_graph_replay": [vp, i],
        "ab200_graph_replay_count": [vp, C.POINTER(C.c_longlong)],
        "ab200_box_copy": [vp, C.POINTER(BoxDesc), i], "ab200_block_bcs": [vp, C.POINTER(BlockBcDesc), i],
        "ab200_set_shear_bc_params": [vp, d, d],
        "ab200_configure_diffusion": [vp, C.POINTER(DiffusionDesc)], "ab200_diffusion_flux": [vp],
        "ab200_diffusion_update": [vp, d], "ab200_diffusion_timestep": [vp, _DP],
        "ab200_diffusion_flux_array": [vp, i, C.POINTER(_DP), C.POINTER(C.c_size_t)],
        "ab200_comm_unique_id": [C.c_char_p], "ab200_comm_init": [vp, i, i, C.c_char_p],
        "ab200_comm_destroy": [vp], "ab200_comm_set_layout": [vp, i, i, i, C.POINTER(C.c_int)],
        "ab200_comm_is_direct": [vp], "ab200_comm_exchange_begin": [vp], "ab200_comm_exchange_end": [vp],
        "ab200_allreduce_min": [vp, vp], "ab200_run_cycles_mr": [vp, i, i, d],
        "ab200_comm_plan_direct": [C.POINTER(C.c_int)] * 5 + [i] + [C.POINTER(C.c_int)] * 5 +
                                  [C.POINTER(C.POINTER(C.c_longlong)), C.POINTER(C.c_int)],
    }
    for name, args in sig.items():
        fn = getattr(L, name)
        fn.argtypes = args
        fn.restype = C.c_int
    _libs[variant] = L
    return L


def check(L, rc, what=""):
    if rc != 0:
        msg = L.ab200_last_error()
        raise AB200Error(f"{what} failed (code {rc}): {msg.decode() if msg else ''}")
