"""ctypes wrapper of oracle/_ref/libartemis_ref.so -- the reference's OWN hot-path sources
compiled against the mock Parthenon (oracle/ref_shim).  TEST INFRASTRUCTURE ONLY.

RefSim is OracleSim with the per-task kernels swapped for the reference's code: K1-K7, K12,
K15 run reference code; the same-level ghost exchange / outflow-reflect BCs stay the oracle's
restatement (Parthenon's boundary communication cannot be compiled here)."""
from __future__ import annotations

import ctypes as C
import os

from . import oracle_py
from .oracle_py import OracleSim, _p, make_fluid

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "_ref", "libartemis_ref.so")
_lib = None


def build():
    import importlib.util
    spec = importlib.util.spec_from_file_location(
        "build_ref", os.path.join(_HERE, "ref_shim", "build_ref.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.build()


def available():
    return os.path.exists(_LIB) or os.path.isdir("/root/reference/src")


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB):
            build()
        _lib = C.CDLL(_LIB)
        _lib.ar_estimate_dt.restype = C.c_double
        _lib.ar_num_threads.restype = C.c_int
    return _lib


class RefSim(OracleSim):
    def __init__(self, *a, **kw):
        super().__init__(*a, **kw)
        self.R = lib()

    def _src_lib(self):   # gravity + shearing box run the reference's own kernels
        return self.R, "ar"

    def _diff_lib(self):  # the reference's own diffusion operators
        return self.R, "ar"

    def CalculateFluxes(self, fs, pcm):
        f = make_fluid(fs.fp)
        self.R.ar_calculate_fluxes(C.byref(self.g), C.byref(f), int(pcm), _p(fs.prim),
                                   *[_p(a) for a in fs.flux], *[_p(a) for a in fs.pflux],
                                   *[_p(a) for a in fs.vface])

    def ApplyUpdate(self, fs, gam0, gam1, beta_dt):
        self.R.ar_apply_update(C.byref(self.g), fs.fp.nvar, _p(fs.u0), _p(fs.u1),
                               *[_p(a) for a in fs.flux], C.c_double(gam0), C.c_double(gam1),
                               C.c_double(beta_dt))

    def FluxSource(self, fs, dt):
        f = make_fluid(fs.fp)
        self.R.ar_flux_source(C.byref(self.g), C.byref(f), _p(fs.prim), _p(fs.u0),
                              *[_p(a) for a in fs.pflux], *[_p(a) for a in fs.vface],
                              C.c_double(self.omf), C.c_double(dt))

    def SetAuxillaryFields(self, fs):
        f = make_fluid(fs.fp)
        self.R.ar_set_aux(C.byref(self.g), C.byref(f), _p(fs.u0))

    def ConsToPrim(self, fs):
        f = make_fluid(fs.fp)
        self.R.ar_cons_to_prim(C.byref(self.g), C.byref(f), _p(fs.u0), _p(fs.prim))

    def PrimToCons(self, fs):
        f = make_fluid(fs.fp)
        self.R.ar_prim_to_cons(C.byref(self.g), C.byref(f), _p(fs.prim), _p(fs.u0))

    def EstimateTimestep(self):
        dts = []
        for fs in self.fluids:
            f = make_fluid(fs.fp)
            dts.append(self.R.ar_estimate_dt(C.byref(self.g), C.byref(f), _p(fs.prim)))
        return min(dts)
