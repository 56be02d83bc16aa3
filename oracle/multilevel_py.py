"""CPU executor of a multilevel ghost-exchange plan (artemis_b200.multilevel.exchange_plan):
TEST INFRASTRUCTURE ONLY.  Box copies and the generic outflow / reflect boundary conditions
(P:bvals/boundary_conditions_generic.hpp:178-256, applied face by face over the full
transverse extent, in face order) are numpy slicing; restriction and prolongation are the
oracle's ao_restrict_average / ao_prolongate_minmod, which are pinned bit for bit to the
reference's own operators (src/utils/refinement/*.hpp, tests/test_refine_oracle.py)."""
import ctypes as C

import numpy as np

from . import oracle_py
from .oracle_py import RefineGeom, prolongate_minmod, restrict_average


def _sl(box):
    return tuple(slice(box[d][0], box[d][1] + 1) for d in (2, 1, 0))


def block_geom(mesh, b):
    ng, nd = mesh.nghost, mesh.ndim
    r = RefineGeom(int(mesh.coords), nd, ng, mesh.ni, mesh.nj, mesh.nk, mesh.cn[0], mesh.cn[1],
                   mesh.cn[2], mesh.is_, mesh.js, mesh.ks, mesh.cs[0], mesh.cs[1], mesh.cs[2])
    for d in range(3):
        r.xmin[d] = float(mesh.blk_xmin[b, d])
        r.dx[d] = float(mesh.blk_dx[b, d])
    return r


def apply_bc(arr, face, kind, n_int, ng_d, vdir):
    """GenericBC on one [nvar][nk][nj][ni] array: ghost layer of `face` (0..5) from the interior.
    n_int = interior zones along the face direction, ng_d = its ghost depth, vdir[v] = 1..3 for
    the vector components that flip under reflection in that direction (0: scalar)."""
    d = face // 2
    ax = 3 - d                      # axis of direction d in [v][k][j][i]
    lo = ng_d                       # first interior index
    hi = ng_d + n_int - 1
    for g in range(ng_d):
        dst = [slice(None)] * 4
        src = [slice(None)] * 4
        if face % 2 == 0:
            dst[ax] = lo - 1 - g
            src[ax] = lo if kind == "outflow" else lo + g
        else:
            dst[ax] = hi + 1 + g
            src[ax] = hi if kind == "outflow" else hi - g
        vals = arr[tuple(src)].copy()
        if kind == "reflect":
            for v in range(arr.shape[0]):
                if vdir[v] == d + 1:
                    vals[v] = -vals[v]
        arr[tuple(dst)] = vals


def coarse_coords(mesh, b):
    """UniformCartesian(src, coarsen = 2) of block b (P:coordinates/uniform_cartesian.hpp:41-55;
    the same arithmetic as the oracle's coarse_coords)"""
    act = (True, mesh.ndim > 1, mesh.ndim > 2)
    cx, cd = [], []
    for d in range(3):
        istart = mesh.nghost if act[d] else 0
        dx = float(mesh.blk_dx[b, d])
        xm = float(mesh.blk_xmin[b, d])
        xm += istart * dx * (1 - 2)
        dx *= 2 if (d == 0 or istart > 0) else 1
        cx.append(xm)
        cd.append(dx)
    return cx, cd


def apply_user_bc(mesh, b, arr, face, kind, is_coarse, user):
    """strat.hpp user condition on the FULL pack array of one block (fine or coarse), in place"""
    code = {"extrap": 5, "inflow": 6}[kind]
    if is_coarse:
        xmin, dx = coarse_coords(mesh, b)
        s, e = mesh.cs, mesh.ce
    else:
        xmin, dx = mesh.blk_xmin[b], mesh.blk_dx[b]
        s = (mesh.is_, mesh.js, mesh.ks)
        e = (mesh.ie, mesh.je, mesh.ke)
    oracle_py.strat_bc(arr, int(mesh.coords), xmin, dx, s, e, user["fluid"], user["S"], face, code,
                       user["q"], user["om0"])


def run_plan(mesh, plan, fine, coarse, vars_, vdir, bc_kinds, user=None):
    """fine [nb][nvar][nk][nj][ni], coarse [nb][nvar][cnk][cnj][cni] (both updated in place);
    vars_: pack entries that are exchanged (FillGhost), vdir: per ENTRY OF vars_ the vector
    direction 1..3 or 0; bc_kinds: 6 names ('outflow' | 'reflect' | 'periodic' | 'extrap' | 'inflow');
    user: fluid / S / q / om0 of the strat.hpp user conditions."""
    L = oracle_py.lib()
    vars_ = list(vars_)
    geoms = [block_geom(mesh, b) for b in range(mesh.nb)]

    def restrict(b, box):
        f = np.ascontiguousarray(fine[b, vars_])
        c = np.ascontiguousarray(coarse[b, vars_])
        flat = [box[0][0], box[0][1], box[1][0], box[1][1], box[2][0], box[2][1]]
        restrict_average(L, geoms[b], f, c, flat)
        coarse[b, vars_] = c

    def prolong(b, box):
        f = np.ascontiguousarray(fine[b, vars_])
        c = np.ascontiguousarray(coarse[b, vars_])
        flat = [box[0][0], box[0][1], box[1][0], box[1][1], box[2][0], box[2][1]]
        prolongate_minmod(L, geoms[b], c, f, flat)
        fine[b, vars_] = f

    for b, box in plan.restrict_send:
        restrict(b, box)
    staged = []
    for sb, sc, sbox, db, dc, dbox in plan.copies:      # all reads before any write
        src = coarse if sc else fine
        staged.append(src[sb][(vars_,) + _sl(sbox)].copy())
    for (sb, sc, sbox, db, dc, dbox), data in zip(plan.copies, staged):
        dst = coarse if dc else fine
        dst[db][(vars_,) + _sl(dbox)] = data
    for b, box in plan.restrict_set:
        restrict(b, box)
    nx = mesh.block_nx
    for face_dir in range(3):                            # x1 faces, then x2, then x3
        for b, face in plan.coarse_bcs:
            if face // 2 == face_dir:
                if bc_kinds[face] in ("extrap", "inflow"):
                    apply_user_bc(mesh, b, coarse[b], face, bc_kinds[face], True, user)
                    continue
                a = coarse[b, vars_]
                apply_bc(a, face, bc_kinds[face], nx[face_dir] // 2, mesh.nghost, vdir)
                coarse[b, vars_] = a
    for b, box in plan.prolongate:
        prolong(b, box)
    for face_dir in range(3):
        for b, face in plan.fine_bcs:
            if face // 2 == face_dir:
                if bc_kinds[face] in ("extrap", "inflow"):
                    apply_user_bc(mesh, b, fine[b], face, bc_kinds[face], False, user)
                    continue
                a = fine[b, vars_]
                apply_bc(a, face, bc_kinds[face], nx[face_dir], mesh.nghost, vdir)
                fine[b, vars_] = a


def flux_correct(mesh, fc_plan, flux):
    """SetFluxCorrections on the CPU: flux = [F1, F2, F3] arrays [nb][nvar][nk][nj][ni] (flux
    through the LOWER face of every zone); every coarse face shared with finer blocks receives
    the area-weighted average of the fine fluxes (RestrictAverage on face elements)."""
    L = oracle_py.lib()
    geoms = {}
    for fb, cb, d, rbox, dbox in fc_plan:
        if fb not in geoms:
            geoms[fb] = block_geom(mesh, fb)
        # face-shaped arrays (the diffusion fluxes, [nvar][fnk][fnj][fni]) address the same
        # (k, j, i) = lower face of zone (k, j, i): restrict their cell-shaped corner
        f = np.ascontiguousarray(flux[d][fb][:, :mesh.nk, :mesh.nj, :mesh.ni])
        nvar = f.shape[0]
        c = np.zeros((nvar, mesh.cn[2], mesh.cn[1], mesh.cn[0]))
        flat = (C.c_int * 6)(rbox[0][0], rbox[0][1], rbox[1][0], rbox[1][1], rbox[2][0], rbox[2][1])
        L.ao_restrict_average_face(C.byref(geoms[fb]), nvar, oracle_py._p(f), oracle_py._p(c), flat,
                                   d + 1)
        flux[d][cb][(slice(None),) + _sl(dbox)] = c[(slice(None),) + _sl(rbox)]
