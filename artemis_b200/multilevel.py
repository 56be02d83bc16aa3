"""Multilevel (static refinement) MeshBlock sets and the plan of their ghost exchange.

Host-side only (numpy).  What Parthenon hands the hot path on a multilevel mesh (SURVEY 8a row
a14, config 5) is, per stage, a list of boundary buffers with index ranges
(`BndInfo::idxer` = `CalcIndices`, P:bvals/comms/bnd_info.cpp:105-252) plus the
restriction / prolongation ranges of `ProResInfo` (:336-447).  This module restates that
bookkeeping so the C ABI's executors (ab200_restrict, ab200_box_copy, ab200_block_bcs,
ab200_prolongate) can be driven and tested without Parthenon:

  * `MultilevelMesh`   leaf blocks of a binary tree over a root lattice (P:mesh/forest/tree.cpp),
                       block domains by LogicalLocation::IndexToSymmetrizedCoordinate, the
                       neighbour rule of Tree::FindNeighborsImpl (:161-226): same-level leaf,
                       the touching daughters of a refined neighbour, or the coarser leaf -- the
                       latter only at the ONE offset `GetSameLevelOffsets` assigns to it;
  * `calc_indices`     CalcIndices for cell-centred fields (pinned against the reference's own
                       function, tests/test_multilevel_plan.py);
  * `exchange_plan`    the order of AddBoundaryExchangeTasks (P:bvals/comms/
                       boundary_communication.cpp:406-445): SendBoundBufs (restrict the region a
                       coarser neighbour needs, pack) -> SetBounds (unpack; restrict freshly set
                       ghost regions into the coarse buffer of blocks that have a coarser
                       neighbour) -> physical BCs on the coarse buffers -> ProlongateBounds ->
                       physical BCs on the fine arrays.
"""
from __future__ import annotations

from dataclasses import dataclass, field

import numpy as np

from .enums import BoundaryFlag, Coordinates
from .mesh import _symm_coord

# IndexRangeType (P:bvals/comms/bnd_info.hpp)
BOUNDARY_INTERIOR_SEND, BOUNDARY_EXTERIOR_RECV, INTERIOR_SEND, INTERIOR_RECV = range(4)


def calc_indices(ng, nx, my_level, my_l, nb_level, nb_l, offsets, ir_type, prores, flux_el=0):
    """CalcIndices (bnd_info.cpp:105-252) for a cell-centred field (flux_el = 0) or for the flux
    of one on the x1 / x2 / x3 faces (flux_el = 1..3: a Metadata::Flux field, element F1..F3).
    Returns ((si, ei), (sj, ej), (sk, ek)) in the fine index space of the block or -- for
    prolongation / restriction ranges and whenever the neighbour is coarser -- in the index space
    of its coarse buffer."""
    flux = flux_el != 0
    top = [int(flux_el == d + 1) for d in range(3)]

    def interior(n, d):
        g = ng if n > 1 else 0
        return g, g + n - 1 + top[d]

    use_coarse = prores or nb_level < my_level
    bounds = [interior((nx[d] // 2 if nx[d] > 1 else 1) if use_coarse else nx[d], d) for d in range(3)]
    coarse_fac = 2 if nb_level > my_level else 1
    nbounds = [interior(nx[d] // coarse_fac, d) for d in range(3)]
    not_symmetry = [nx[d] > 1 for d in range(3)]
    interior_offset = ng if ir_type == BOUNDARY_INTERIOR_SEND else 0
    exterior_offset = ng if ir_type == BOUNDARY_EXTERIOR_RECV else 0
    if prores:
        exterior_offset //= 2
    out = []
    for d in range(3):
        if offsets[d] == 0:
            s, e = bounds[d]
            if my_level < nb_level and not_symmetry[d]:
                extra = (bounds[d][1] - bounds[d][0] + 1) - (nbounds[d][1] - nbounds[d][0] + 1)
                if nb_l[d] % 2 == 1:
                    s += extra - interior_offset
                else:
                    e -= extra - interior_offset
                if ir_type == INTERIOR_SEND and not prores:
                    s -= ng
                    e += ng
            if my_level > nb_level and not_symmetry[d]:
                if my_l[d] % 2 == 1:
                    s -= exterior_offset
                else:
                    e += exterior_offset
                if ir_type == INTERIOR_RECV and not prores:
                    s -= ng
                    e += ng
            if prores and not_symmetry[d] and ir_type == INTERIOR_RECV:
                s -= ng // 2
                e += ng // 2
        elif offsets[d] > 0:      # fluxes are only communicated on shared elements
            s = bounds[d][1] + (0 if flux else -interior_offset + 1 - top[d])
            e = bounds[d][1] + (0 if flux else exterior_offset)
        else:
            s = bounds[d][0] + (0 if flux else -exterior_offset)
            e = bounds[d][0] + (0 if flux else interior_offset - 1 + top[d])
        out.append((s, e))
    return tuple(out)


@dataclass
class Neighbor:
    gid: int
    level: int
    loc: tuple          # LogicalLocation of the neighbour (wrapped into the domain)
    origin_loc: tuple   # the same location seen from this block (not wrapped)
    offsets: tuple      # GetSameLevelOffsets(origin_loc)


@dataclass
class MultilevelMesh:
    """Root lattice of `root_blocks` MeshBlocks, the root blocks listed in `refine` replaced by
    their 2^ndim daughters (one extra level; `levels` > 1 refines the daughters that touch the
    centre of the refined region again, keeping 2:1 nesting the caller's responsibility)."""
    root_blocks: tuple
    block_nx: tuple
    xmin: tuple
    xmax: tuple
    refine: tuple = ()                  # root-lattice locations (l1, l2, l3) to refine
    nghost: int = 4
    bcs: tuple = (BoundaryFlag.periodic,) * 6
    coords: Coordinates = Coordinates.cartesian
    ndim: int = field(init=False)

    def __post_init__(self):
        self.block_nx = tuple(int(v) for v in self.block_nx)
        self.root_blocks = tuple(int(v) for v in self.root_blocks)
        self.ndim = 1 + (self.block_nx[1] > 1) + (self.block_nx[2] > 1)
        ng = self.nghost
        if ng % 2:
            raise ValueError("Parthenon requires an even nghost with mesh refinement "
                             "(P:mesh/mesh_refinement.cpp:61)")
        for d in range(self.ndim):
            if self.block_nx[d] % 2 or self.block_nx[d] < 2 * ng:
                raise ValueError("multilevel MeshBlocks need an even nx >= 2 nghost per active "
                                 "direction (the coarse buffer must hold nghost interior zones)")
        self.ngd = tuple(ng if self.block_nx[d] > 1 else 0 for d in range(3))
        self.ni, self.nj, self.nk = (self.block_nx[d] + 2 * self.ngd[d] for d in range(3))
        self.is_, self.js, self.ks = self.ngd
        self.ie = self.is_ + self.block_nx[0] - 1
        self.je = self.js + self.block_nx[1] - 1
        self.ke = self.ks + self.block_nx[2] - 1
        self.fni = self.ni + (1 if self.ni > 1 else 0)
        self.fnj = self.nj + (1 if self.nj > 1 else 0)
        self.fnk = self.nk + (1 if self.nk > 1 else 0)
        # coarse buffer: nx/2 interior zones + nghost ghosts per active direction
        # (P:mesh/meshblock.cpp:205-228)
        self.cn = tuple(self.block_nx[d] // 2 + 2 * ng if d < self.ndim else 1 for d in range(3))
        self.cs = tuple(ng if d < self.ndim else 0 for d in range(3))
        self.ce = tuple(self.cs[d] + (self.block_nx[d] // 2 if d < self.ndim else 1) - 1
                        for d in range(3))
        # ---- the tree: leaves in Morton-free, deterministic order (level, l3, l2, l1) --------
        refined = {tuple(r) for r in self.refine}
        leaves = []
        for l3 in range(self.root_blocks[2]):
            for l2 in range(self.root_blocks[1]):
                for l1 in range(self.root_blocks[0]):
                    if (l1, l2, l3) in refined:
                        for c3 in range(2 if self.ndim > 2 else 1):
                            for c2 in range(2 if self.ndim > 1 else 1):
                                for c1 in range(2):
                                    leaves.append((1, (2 * l1 + c1, 2 * l2 + c2 if self.ndim > 1 else 0,
                                                       2 * l3 + c3 if self.ndim > 2 else 0)))
                    else:
                        leaves.append((0, (l1, l2, l3)))
        self.leaves = leaves
        self.internal = {(0, r) for r in refined}
        self.gid_of = {lf: g for g, lf in enumerate(leaves)}
        self.nb = len(leaves)
        self.level = np.array([lf[0] for lf in leaves], dtype=np.int32)
        self.blk_loc = np.array([lf[1] for lf in leaves], dtype=np.int64)
        self.blk_xmin = np.zeros((self.nb, 3))
        self.blk_dx = np.zeros((self.nb, 3))
        self.blk_lo = np.zeros((self.nb, 3))
        for b, (lev, l) in enumerate(leaves):
            for d in range(3):
                if self.block_nx[d] > 1:
                    nrange = self.root_blocks[d] << lev
                    ul, ur = _symm_coord(l[d], 0, nrange), _symm_coord(l[d], 2, nrange)
                    mid = 0.5 * (self.xmin[d] + self.xmax[d])
                    lo = mid + (ul * self.xmax[d] - ul * self.xmin[d])
                    hi = mid + (ur * self.xmax[d] - ur * self.xmin[d])
                else:
                    lo, hi = self.xmin[d], self.xmax[d]
                dx = (hi - lo) / self.block_nx[d]
                self.blk_lo[b, d] = lo
                self.blk_dx[b, d] = dx
                self.blk_xmin[b, d] = lo - self.ngd[d] * dx
        self.neighbors = [self._find_neighbors(b) for b in range(self.nb)]

    # ---- duck-typing of UniformMesh (what MeshData reads) ------------------------------------
    @property
    def interior_zones(self):
        return self.nb * self.block_nx[0] * self.block_nx[1] * self.block_nx[2]

    def shape(self, nvar):
        return (self.nb, nvar, self.nk, self.nj, self.ni)

    def coarse_shape(self, nvar):
        return (self.nb, nvar, self.cn[2], self.cn[1], self.cn[0])

    def face_shape(self, ns):
        return (self.nb, ns, self.fnk, self.fnj, self.fni)

    @property
    def lattice_n(self):
        # the uniform-lattice exchange of the library is never used on a multilevel mesh
        return (self.nb, 1, 1)

    def interior(self):
        return (slice(self.ks, self.ke + 1), slice(self.js, self.je + 1),
                slice(self.is_, self.ie + 1))

    def bc_ints(self):
        return np.array([int(v) for v in self.bcs], dtype=np.int32)

    # ---- Tree::FindNeighborsImpl -------------------------------------------------------------
    def _nlevel(self, lev, d):
        return (self.root_blocks[d] << lev) if d < self.ndim else 1

    def _wrap(self, lev, l):
        """location inside the domain, or None across a non-periodic boundary"""
        out = []
        for d in range(3):
            n = self._nlevel(lev, d)
            v = l[d]
            if v < 0 or v >= n:
                if self.bcs[2 * d] != BoundaryFlag.periodic:
                    return None
                v %= n
            out.append(v)
        return tuple(out)

    def _find_neighbors(self, b):
        lev, l = self.leaves[b]
        act = [(-1, 0, 1) if d < self.ndim else (0,) for d in range(3)]
        out = []
        for o3 in act[2]:
            for o2 in act[1]:
                for o1 in act[0]:
                    o = (o1, o2, o3)
                    if o == (0, 0, 0):
                        continue
                    neigh = tuple(l[d] + o[d] for d in range(3))
                    w = self._wrap(lev, neigh)
                    if w is None:
                        continue
                    if (lev, w) in self.gid_of:                      # same-level leaf
                        out.append(Neighbor(self.gid_of[(lev, w)], lev, w, neigh, o))
                    elif (lev, w) in self.internal:                  # refined: touching daughters
                        for c3 in range(2 if self.ndim > 2 else 1):
                            for c2 in range(2 if self.ndim > 1 else 1):
                                for c1 in range(2):
                                    c = (c1, c2, c3)
                                    dl = tuple(2 * neigh[d] + c[d] if d < self.ndim else 0
                                               for d in range(3))
                                    if not self._is_neighbor(lev, l, lev + 1, dl):
                                        continue
                                    dw = tuple(2 * w[d] + c[d] if d < self.ndim else 0 for d in range(3))
                                    off = tuple((dl[d] >> 1) - l[d] for d in range(3))
                                    out.append(Neighbor(self.gid_of[(lev + 1, dw)], lev + 1, dw, dl, off))
                    elif lev > 0:                                    # coarser leaf, one offset only
                        pw = tuple(v >> 1 for v in w)
                        if (lev - 1, pw) in self.gid_of:
                            pn = tuple(v >> 1 for v in neigh)        # floor division keeps -1 -> -1
                            sl = tuple(pn[d] - (l[d] >> 1) for d in range(3))
                            if sl == o:
                                out.append(Neighbor(self.gid_of[(lev - 1, pw)], lev - 1, pw, pn, o))
        return out

    @staticmethod
    def _is_neighbor(lev_a, la, lev_b, lb):
        """LogicalLocation::IsNeighbor (logical_location.cpp:110-129)"""
        ml = max(lev_a, lev_b)
        sa, sb = 1 << (ml - lev_a), 1 << (ml - lev_b)
        for d in range(3):
            low = la[d] * sa - 1
            hi = low + sa + 1
            low_in = lb[d] * sb
            hi_in = low_in + sb - 1
            if hi < low_in or low > hi_in:
                return False
        return True

    def physical_faces(self, b):
        """faces (0..5 = ix1, ox1, ix2, ox2, ix3, ox3) of block b on a non-periodic boundary"""
        lev, l = self.leaves[b]
        faces = []
        for d in range(self.ndim):
            if self.bcs[2 * d] == BoundaryFlag.periodic:
                continue
            if l[d] == 0:
                faces.append(2 * d)
            if l[d] == self._nlevel(lev, d) - 1:
                faces.append(2 * d + 1)
        return faces


def flux_correction_plan(mesh) -> list:
    """AddFluxCorrectionTasks (boundary_communication.cpp:454-461): every fine block sends the
    fluxes through each face it shares with a COARSER block (loop_utils.hpp:146-159: face
    neighbours one level down), restricted to the coarse face (ProResInfo::GetSend), and the
    coarse block overwrites its own flux there (SetBounds<flxcor_recv>).  Entries:
    (fine block, coarse block, dir 0..2, restrict box in the fine block's coarse index space,
    destination box in the coarse block's index space)."""
    ng, nx = mesh.nghost, mesh.block_nx
    out = []
    for b in range(mesh.nb):
        lev, l = mesh.leaves[b]
        for nb in mesh.neighbors[b]:
            if nb.level != lev - 1 or sum(abs(o) for o in nb.offsets) != 1:
                continue
            d = [abs(o) for o in nb.offsets].index(1)
            rbox = calc_indices(ng, nx, lev, l, nb.level, nb.origin_loc, nb.offsets,
                                BOUNDARY_INTERIOR_SEND, True, flux_el=d + 1)
            back = [m for m in mesh.neighbors[nb.gid]
                    if m.gid == b and m.offsets == tuple(-o for o in nb.offsets)]
            m = back[0]
            clev, cl = mesh.leaves[nb.gid]
            dbox = calc_indices(ng, nx, clev, cl, m.level, m.origin_loc, m.offsets,
                                BOUNDARY_EXTERIOR_RECV, False, flux_el=d + 1)
            if any(rbox[q][1] - rbox[q][0] != dbox[q][1] - dbox[q][0] for q in range(3)):
                raise RuntimeError(f"flux-correction shapes differ: {b} -> {nb.gid}: {rbox} {dbox}")
            out.append((b, nb.gid, d, rbox, dbox))
    return out


@dataclass
class ExchangePlan:
    """Index boxes of one multilevel ghost exchange; boxes are ((si, ei), (sj, ej), (sk, ek))."""
    restrict_send: list = field(default_factory=list)   # (block, coarse box)
    copies: list = field(default_factory=list)          # (src block, src coarse?, src box,
    #                                                       dst block, dst coarse?, dst box)
    restrict_set: list = field(default_factory=list)    # (block, coarse box)
    coarse_bcs: list = field(default_factory=list)      # (block, face)
    prolongate: list = field(default_factory=list)      # (block, coarse box)
    fine_bcs: list = field(default_factory=list)        # (block, face)


def exchange_plan(mesh: MultilevelMesh) -> ExchangePlan:
    ng, nx = mesh.nghost, mesh.block_nx
    plan = ExchangePlan()

    def ci(b, nb, ir, prores):
        lev, l = mesh.leaves[b]
        return calc_indices(ng, nx, lev, l, nb.level, nb.origin_loc, nb.offsets, ir, prores)

    for b in range(mesh.nb):
        lev = mesh.leaves[b][0]
        has_coarser = any(nb.level < lev for nb in mesh.neighbors[b])
        for nb in mesh.neighbors[b]:
            # ---- send side: b -> nb (BndInfo::GetSendBndInfo, ProResInfo::GetSend) -------------
            src_coarse = nb.level < lev
            if src_coarse:
                plan.restrict_send.append((b, ci(b, nb, BOUNDARY_INTERIOR_SEND, True)))
            sbox = ci(b, nb, BOUNDARY_INTERIOR_SEND, False)
            # ---- the matching receive entry of nb (its neighbour b at the opposite offset) ---
            back = [m for m in mesh.neighbors[nb.gid]
                    if m.gid == b and m.offsets == tuple(-o for o in nb.offsets)]
            if len(back) != 1:
                raise RuntimeError(f"no unique return neighbour for block {b} -> {nb.gid}")
            m = back[0]
            rlev, rl = mesh.leaves[nb.gid]
            dbox = calc_indices(ng, nx, rlev, rl, m.level, m.origin_loc, m.offsets,
                                BOUNDARY_EXTERIOR_RECV, False)
            dst_coarse = m.level < rlev
            if any(sbox[d][1] - sbox[d][0] != dbox[d][1] - dbox[d][0] for d in range(3)):
                raise RuntimeError(f"buffer shapes differ: {b} -> {nb.gid}: {sbox} vs {dbox}")
            plan.copies.append((b, src_coarse, sbox, nb.gid, dst_coarse, dbox))
            # ---- set side of b for this neighbour (ProResInfo::GetSet) --------------------------
            if nb.level < lev:
                plan.prolongate.append((b, ci(b, nb, BOUNDARY_EXTERIOR_RECV, True)))
            elif has_coarser:
                plan.restrict_set.append((b, ci(b, nb, BOUNDARY_EXTERIOR_RECV, True)))
        for face in mesh.physical_faces(b):
            plan.fine_bcs.append((b, face))
            if has_coarser:
                plan.coarse_bcs.append((b, face))
    return plan


def fill_ghost_ranges(fp):
    """contiguous (var0, ncomp) ranges of the FillGhost primitive pack entries: gas density +
    velocity and sie (the pressure entry is Derived, src/gas/gas.cpp:236-270), dust density +
    velocity"""
    from .enums import Fluid
    S = fp.nspecies
    if fp.fluid_type == Fluid.gas:
        return [(0, 4 * S), (5 * S, S)]
    return [(0, 4 * S)]


class MultilevelExchange:
    """AddBoundaryExchangeTasks on a MultilevelMesh whose blocks live on one GPU: descriptor
    lists built once from `exchange_plan`, executed through the C ABI in Parthenon's order."""
    multilevel = True

    @staticmethod
    def allreduce_min(x):     # one rank: the driver's dt all-reduce is the identity
        return x

    def __init__(self, md, plan: ExchangePlan | None = None, entry_tables: bool = False):
        """entry_tables: hand the user conditions their coarse arrays the way a Parthenon host
        has them -- one device pointer per pack entry (ab200_block_bc_desc.coarse_entries; one
        coarse buffer per Variable) -- instead of one slab per fluid"""
        import ctypes as C

        from . import capi
        self.md, self.C, self.capi = md, C, capi
        m = md.mesh
        self.plan = plan or exchange_plan(m)
        shape6 = (C.c_int * 6)()
        md.call("ab200_coarse_shape", shape6)
        assert tuple(shape6[:3]) == tuple(m.cn), (tuple(shape6), m.cn)
        ccells = m.cn[0] * m.cn[1] * m.cn[2]
        kinds = {int(BoundaryFlag.outflow): 1, int(BoundaryFlag.reflect): 2}
        # strat.hpp user conditions take the whole fluid: one descriptor per (block, face)
        user = (int(BoundaryFlag.extrap), int(BoundaryFlag.inflow))
        self.coarse = {}
        rs, cp, rt, cb, pr, fb = [], [], [], [], [], []
        for ff in md.fluids:
            fl, nv = int(ff.fp.fluid_type), ff.fp.nvar
            buf = md._alloc(m.coarse_shape(nv))
            self.coarse[fl] = buf

            def cptr(b, var0, buf=buf, nv=nv):
                return buf.ptr + ((b * nv + var0) * ccells) * 8

            for var0, nc in fill_ghost_ranges(ff.fp):
                def refine(lst, b, box):
                    lst.append(capi.RefineDesc(fl, b, var0, nc, 0, box[0][0], box[0][1], box[1][0],
                                               box[1][1], box[2][0], box[2][1], cptr(b, var0)))
                for b, box in self.plan.restrict_send:
                    refine(rs, b, box)
                for b, box in self.plan.restrict_set:
                    refine(rt, b, box)
                for b, box in self.plan.prolongate:
                    refine(pr, b, box)
                for sb, sc, sbox, db, dc, dbox in self.plan.copies:
                    cp.append(capi.BoxDesc(fl, nc, sb, var0, cptr(sb, var0) if sc else None, db, var0,
                                           cptr(db, var0) if dc else None, sbox[0][0], sbox[1][0],
                                           sbox[2][0], dbox[0][0], dbox[1][0], dbox[2][0],
                                           sbox[0][1] - sbox[0][0] + 1, sbox[1][1] - sbox[1][0] + 1,
                                           sbox[2][1] - sbox[2][0] + 1))
                for b, face in self.plan.coarse_bcs:
                    if int(m.bcs[face]) not in user:
                        cb.append(capi.BlockBcDesc(fl, b, var0, nc, face, kinds[int(m.bcs[face])],
                                                   cptr(b, var0)))
                for b, face in self.plan.fine_bcs:
                    if int(m.bcs[face]) not in user:
                        fb.append(capi.BlockBcDesc(fl, b, var0, nc, face, kinds[int(m.bcs[face])], None))
            if entry_tables and any(int(m.bcs[f]) in user for _, f in self.plan.coarse_bcs):
                # [block][entry] -> that entry's coarse array, as 64-bit words on the device
                ptrs = np.array([cptr(b, v) for b in range(m.nb) for v in range(nv)], dtype=np.uint64)
                tab = md._alloc((m.nb * nv,))
                tab.set(ptrs.view(np.float64))
                self._entry_tabs = getattr(self, "_entry_tabs", []) + [tab]
            for b, face in self.plan.coarse_bcs:
                if int(m.bcs[face]) in user:
                    if entry_tables:
                        cb.append(capi.BlockBcDesc(fl, b, 0, nv, face, int(m.bcs[face]), None,
                                                   tab.ptr + b * nv * 8))
                    else:
                        cb.append(capi.BlockBcDesc(fl, b, 0, nv, face, int(m.bcs[face]), cptr(b, 0)))
            for b, face in self.plan.fine_bcs:
                if int(m.bcs[face]) in user:
                    fb.append(capi.BlockBcDesc(fl, b, 0, nv, face, int(m.bcs[face]), None))

        def arr(T, lst):
            return (T * len(lst))(*lst) if lst else None

        self._rs, self._rt, self._pr = arr(capi.RefineDesc, rs), arr(capi.RefineDesc, rt), arr(capi.RefineDesc, pr)
        self._cp = arr(capi.BoxDesc, cp)
        self._cb, self._fb = arr(capi.BlockBcDesc, cb), arr(capi.BlockBcDesc, fb)
        fcs = []
        for ff in md.fluids:
            for fbk, cbk, d, rbox, dbox in flux_correction_plan(m):
                fcs.append(capi.FluxCorDesc(int(ff.fp.fluid_type), fbk, cbk, d, rbox[0][0], rbox[0][1],
                                            rbox[1][0], rbox[1][1], rbox[2][0], rbox[2][1],
                                            dbox[0][0], dbox[1][0], dbox[2][0]))
        self._fc = arr(capi.FluxCorDesc, fcs)
        self.n = dict(rs=len(rs), rt=len(rt), pr=len(pr), cp=len(cp), cb=len(cb), fb=len(fb),
                      fc=len(fcs))

    def exchange(self):
        """SendBoundBufs -> SetBounds -> coarse BCs -> ProlongateBounds -> fine BCs"""
        md, n = self.md, self.n
        if n["rs"]:
            md.call("ab200_restrict", self._rs, n["rs"])
        if n["cp"]:
            md.call("ab200_box_copy", self._cp, n["cp"])
        if n["rt"]:
            md.call("ab200_restrict", self._rt, n["rt"])
        if n["cb"]:
            md.call("ab200_block_bcs", self._cb, n["cb"])
        if n["pr"]:
            md.call("ab200_prolongate", self._pr, n["pr"])
        if n["fb"]:
            md.call("ab200_block_bcs", self._fb, n["fb"])

    def flux_correct(self):
        """AddFluxCorrectionTasks: between CalculateFluxes and ApplyUpdate (task path)"""
        if self.n["fc"]:
            self.md.call("ab200_flux_correct", self._fc, self.n["fc"])

    def close(self):
        for buf in list(self.coarse.values()) + getattr(self, "_entry_tabs", []):
            buf.free()
        self.coarse = {}
        self._entry_tabs = []
