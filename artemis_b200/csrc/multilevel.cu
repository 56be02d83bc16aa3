// multilevel.cu -- the data movement of the multilevel ghost exchange that the same-level
// kernels (halo.cu) and the operators of refine.cu do not cover:
//
//   ab200_box_copy    SendBoundBufs + SetBounds for blocks on the same GPU
//                     (P:bvals/comms/boundary_communication.cpp:48-140, 263-352) as ONE launch
//                     over a descriptor list: every descriptor copies a box of `ncomp` pack
//                     entries from a block's fine arrays or its coarse buffer (Parthenon's
//                     `coarse_s`) into another block's fine arrays or coarse buffer -- the
//                     three cases of BndInfo (bnd_info.cpp:273-304): same level (fine -> fine),
//                     to a coarser block (restricted coarse buffer -> fine ghosts) and to a
//                     finer block (fine interior -> the receiver's coarse-buffer ghosts).
//                     Index boxes are CalcIndices' (bnd_info.cpp:105-252), supplied by the host
//                     from Parthenon's own boundary cache.  Every box read is interior data,
//                     every box written is a ghost region, so one launch is race-free.
//   ab200_block_bcs   ApplyBoundaryConditionsOnCoarseOrFineMD: GenericBC outflow / reflect
//                     (P:bvals/boundary_conditions_generic.hpp:178-256) on named faces of named
//                     blocks, on the fine arrays or on the coarse buffers, over the full
//                     transverse extent, x1 faces first, then x2, then x3 (one launch each);
//                     plus the state-dependent user conditions of the shearing-box problem
//                     generators (`extrap`, `inflow`: src/pgen/strat.hpp:154-666) in the same
//                     list and the same order (k_block_user_bcs).
//
//   ab200_flux_correct   AddFluxCorrectionTasks (boundary_communication.cpp:454-461): the fluxes
//                     through every coarse face shared with finer blocks are replaced by the
//                     area-weighted average of the fine fluxes -- RestrictAverage<GEOM> on face
//                     elements (src/utils/refinement/restriction.hpp:41-114) fused with the
//                     send / set copies into ONE launch over all fine-coarse faces, all
//                     Metadata::Flux fields of a fluid (conserved fluxes + interface pressure).
//
// All are thin-slab streaming kernels: HBM-bound by nature, launch-latency bound in practice.
#include <cstring>

#include "ab200_ctx.cuh"
#include "tasks.cuh"

namespace ab200 {

struct BoxDev {
  int fluid, ncomp;
  int src_block, src_var0, dst_block, dst_var0;
  const double *src_coarse;
  double *dst_coarse;
  int ssi, ssj, ssk, dsi, dsj, dsk, ni, nj, nk;
};

// one thread per (descriptor = blockIdx.y, component, cell of the box), i fastest
__global__ void __launch_bounds__(kThreads)
k_box_copy(GridDev g, GridDev gc, FluidDev f0, FluidDev f1, const BoxDev *__restrict__ bx) {
  const BoxDev d = bx[blockIdx.y];
  const FluidDev &f = d.fluid == AB200_GAS ? f0 : f1;
  const long long cells = (long long)d.ni * d.nj * d.nk;
  const long long total = cells * d.ncomp;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const int comp = (int)(t / cells);
    long long r = t - comp * cells;
    const int i = (int)(r % d.ni); r /= d.ni;
    const int j = (int)(r % d.nj);
    const int k = (int)(r / d.nj);
    double v;
    if (d.src_coarse)
      v = d.src_coarse[(((size_t)comp * gc.nk + (d.ssk + k)) * gc.nj + (d.ssj + j)) * gc.ni + (d.ssi + i)];
    else
      v = f.prim[(size_t)d.src_block * f.nvar + d.src_var0 + comp]
                [((size_t)(d.ssk + k) * g.nj + (d.ssj + j)) * g.ni + (d.ssi + i)];
    if (d.dst_coarse)
      d.dst_coarse[(((size_t)comp * gc.nk + (d.dsk + k)) * gc.nj + (d.dsj + j)) * gc.ni + (d.dsi + i)] = v;
    else
      f.prim[(size_t)d.dst_block * f.nvar + d.dst_var0 + comp]
            [((size_t)(d.dsk + k) * g.nj + (d.dsj + j)) * g.ni + (d.dsi + i)] = v;
  }
}

struct BcDev {
  int fluid, block, var0, ncomp, face, type;
  double *coarse;
  double *const *ctab;  // user conditions: per-entry coarse arrays (one per Parthenon Variable)
};

// GenericBC of one direction: thread per (descriptor, component, transverse cell, ghost layer)
__global__ void __launch_bounds__(kThreads)
k_block_bcs(GridDev g, GridDev gc, FluidDev f0, FluidDev f1, const BcDev *__restrict__ bc, int dir) {
  const BcDev d = bc[blockIdx.y];
  if (d.face / 2 != dir) return;
  if (d.type != AB200_BC_OUTFLOW && d.type != AB200_BC_REFLECT) return;  // k_block_user_bcs
  const FluidDev &f = d.fluid == AB200_GAS ? f0 : f1;
  const GridDev &a = d.coarse ? gc : g;  // the index space the face lives in
  const int n[3] = {a.ni, a.nj, a.nk};
  const int lo[3] = {a.is, a.js, a.ks}, hi[3] = {a.ie, a.je, a.ke};
  const int ng = lo[dir];                 // ghost depth of this direction
  const int t1 = (dir + 1) % 3, t2 = (dir + 2) % 3;
  const long long plane = (long long)n[t1] * n[t2];
  const long long total = plane * ng * d.ncomp;
  const bool inner = (d.face % 2) == 0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    int idx[3];
    // enumerate so that consecutive threads walk i (direction 0) whenever it is transverse
    const int a1 = dir == 0 ? 1 : 0, a2 = dir == 2 ? 1 : 2;  // the two transverse axes, low first
    idx[a1] = (int)(r % n[a1]); r /= n[a1];
    idx[a2] = (int)(r % n[a2]); r /= n[a2];
    const int gl = (int)(r % ng);
    const int comp = (int)(r / ng);
    (void)t1; (void)t2;
    int src;
    if (inner) {
      idx[dir] = lo[dir] - 1 - gl;
      src = d.type == AB200_BC_REFLECT ? lo[dir] + gl : lo[dir];
    } else {
      idx[dir] = hi[dir] + 1 + gl;
      src = d.type == AB200_BC_REFLECT ? hi[dir] - gl : hi[dir];
    }
    int sidx[3] = {idx[0], idx[1], idx[2]};
    sidx[dir] = src;
    const int var = d.var0 + comp;
    // the velocity component normal to the face flips under reflection
    // (boundary_conditions_generic.hpp:229-246): pack entries S + 3 n + dir
    double sgn = 1.0;
    if (d.type == AB200_BC_REFLECT && var >= f.S && var < 4 * f.S && (var - f.S) % 3 == dir) sgn = -1.0;
    double *base = d.coarse ? d.coarse + (size_t)comp * n[0] * n[1] * n[2]
                            : f.prim[(size_t)d.block * f.nvar + var];
    const double v = base[((size_t)sidx[2] * n[1] + sidx[1]) * n[0] + sidx[0]];
    base[((size_t)idx[2] * n[1] + idx[1]) * n[0] + idx[0]] = sgn * v;
  }
}

// The shearing-box user conditions of the strat / ssheet problem generators
// (src/pgen/strat.hpp:154-666; inputs/ssheet/ssheet.in: `extrap` on x1 and x3, `inflow` on x2).
// They are functions of the STATE next to the face, so unlike AB200_BC_FIXED they run every
// exchange.  One thread per (transverse cell, ghost layer, species); the gas condition touches
// species 0 only (the reference writes gas::prim::*(0)), the dust condition every species.
//   AB200_BC_EXTRAP, x1 faces  (ExtrapInnerX1 / ExtrapOuterX1): density, v3, sie copied from the
//       last interior cell, v1 copied with inflow clipped to zero, v2 extrapolated linearly in x1
//   AB200_BC_INFLOW, x2 faces  (ShearInnerX2 / ShearOuterX2): copy; v2 = -q Om0 x on the half
//       of the face where the background shear flows INTO the box, outflow-only on the other
//   AB200_BC_EXTRAP, x3 faces  (ExtrapInnerX3 / ExtrapOuterX3): copy, v3 with inflow clipped,
//       density extrapolated as a power law in x3 (hydrostatic stratification)
// Positions come from the metric tables of the index space the face lives in (fine arrays or
// coarse buffers), i.e. geometry::Coords<GEOM>::x1v / x3v / bnds.x1[0] of the reference.
__global__ void __launch_bounds__(kThreads)
k_block_user_bcs(GridDev g, GridDev gc, FluidDev f0, FluidDev f1, const BcDev *__restrict__ bc,
                 int dir, double shear_q, double shear_om0) {
  const BcDev d = bc[blockIdx.y];
  if (d.face / 2 != dir || (d.type != AB200_BC_EXTRAP && d.type != AB200_BC_INFLOW)) return;
  const FluidDev &f = d.fluid == AB200_GAS ? f0 : f1;
  const bool gas = d.fluid == AB200_GAS;
  const GridDev &a = (d.coarse || d.ctab) ? gc : g;
  const int n[3] = {a.ni, a.nj, a.nk};
  const int lo[3] = {a.is, a.js, a.ks}, hi[3] = {a.ie, a.je, a.ke};
  const int ng = lo[dir];
  const int a1 = dir == 0 ? 1 : 0, a2 = dir == 2 ? 1 : 2;
  const int nsp = gas ? 1 : f.S;
  const long long total = (long long)n[a1] * n[a2] * ng * nsp;
  const bool inner = (d.face % 2) == 0;
  const int ref = inner ? lo[dir] : hi[dir];     // last interior cell (is / ie of the reference)
  const int ref1 = inner ? ref + 1 : ref - 1;    // its interior neighbour (is + 1 / ie - 1)
  const size_t cells = (size_t)n[0] * n[1] * n[2];
  const double *x1v = a.t.x1v + (size_t)d.block * n[0];
  const double *x3v = a.t.x3v + (size_t)d.block * n[2];
  const double *x1f = a.t.x1f + (size_t)d.block * (n[0] + 1);
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    int idx[3];
    idx[a1] = (int)(r % n[a1]); r /= n[a1];
    idx[a2] = (int)(r % n[a2]); r /= n[a2];
    const int gl = (int)(r % ng);
    const int sp = (int)(r / ng);
    idx[dir] = inner ? lo[dir] - 1 - gl : hi[dir] + 1 + gl;
    int s0[3] = {idx[0], idx[1], idx[2]}, s1[3] = {idx[0], idx[1], idx[2]};
    s0[dir] = ref; s1[dir] = ref1;
    const size_t o = ((size_t)idx[2] * n[1] + idx[1]) * n[0] + idx[0];
    const size_t o0 = ((size_t)s0[2] * n[1] + s0[1]) * n[0] + s0[0];
    const size_t o1 = ((size_t)s1[2] * n[1] + s1[1]) * n[0] + s1[0];
    // pack entries: density n, velocity S + 3 n + dir, sie 5 S + n (hllc.hpp:66-73; the
    // pressure at 4 S + n is not a FillGhost field, PrimToCons recomputes it)
    const int vd = sp, vv = f.S + 3 * sp, ve = 5 * f.S + sp;
    auto ptr = [&](int var) -> double * {
      if (d.ctab) return d.ctab[var];
      return d.coarse ? d.coarse + (size_t)var * cells : f.prim[(size_t)d.block * f.nvar + var];
    };
    double *pd = ptr(vd), *p1 = ptr(vv), *p2 = ptr(vv + 1), *p3 = ptr(vv + 2);
    const double dens = pd[o0], v1 = p1[o0], v2 = p2[o0], v3 = p3[o0];
    double od = dens, o_v1 = v1, o_v2 = v2, o_v3 = v3;
    if (d.type == AB200_BC_EXTRAP && dir == 0) {
      const double x0 = x1v[ref], xn = x1v[ref1];
      const double dx = inner ? xn - x0 : x0 - xn;
      const double x = x1v[idx[0]];
      const double v2n = p2[o1];
      o_v1 = inner ? ((v1 > 0.0) ? 0.0 : v1) : ((v1 < 0.0) ? 0.0 : v1);
      o_v2 = inner ? v2 + (v2n - v2) * (x - x0) / dx : v2 + (v2 - v2n) * (x - x0) / dx;
    } else if (d.type == AB200_BC_EXTRAP) {  // dir == 2
      const double z0 = x3v[ref], zn = x3v[ref1];
      const double dz = inner ? zn - z0 : z0 - zn;
      const double z = x3v[idx[2]];
      const double dn = pd[o1];
      o_v3 = inner ? ((v3 > 0.0) ? 0.0 : v3) : ((v3 < 0.0) ? 0.0 : v3);
      const double drho = inner ? dn / dens : dens / dn;
      od = dens * pow(drho, (z - z0) / dz);
    } else {  // AB200_BC_INFLOW, dir == 1
      const double x = x1v[idx[0]], xf = x1f[idx[0]];
      const double vy0 = -shear_q * shear_om0 * x;
      if (inner) o_v2 = (xf >= 0) ? ((v2 > 0.0) ? 0.0 : v2) : vy0;
      else o_v2 = (xf < 0) ? ((v2 < 0.0) ? 0.0 : v2) : vy0;
    }
    p1[o] = o_v1; p2[o] = o_v2; p3[o] = o_v3; pd[o] = od;
    if (gas) { double *pe = ptr(ve); pe[o] = pe[o0]; }
  }
}

constexpr int kMaxGridY = 65535;

struct FluxCorDev {
  int fluid, fine_block, coarse_block, dir;
  int cis, cie, cjs, cje, cks, cke;  // coarse-index box of the fine block's face
  int dsi, dsj, dsk;                 // origin of the same cells in the coarse block
};

// RestrictAverage<GEOM> on the face element el (1..3) of one coarse face cell whose lower-corner
// fine cell is (k, j, i) of `block`: the area-weighted mean of the 2^(ndim-1) fine faces
// (src/utils/refinement/restriction.hpp:41-114).  `fine` is one [..][snj][sni] array.
template <int GEOM>
AB_D double restrict_face(const GridDev &g, int block, int el, int k, int j, int i,
                          const double *__restrict__ fine, int snj, int sni) {
  const bool inc1 = el != 1, inc2 = g.ndim > 1 && el != 2, inc3 = g.ndim > 2 && el != 3;
  double vol[2][2][2], terms[2][2][2];
#pragma unroll
  for (int ok = 0; ok < 2; ++ok)
#pragma unroll
    for (int oj = 0; oj < 2; ++oj)
#pragma unroll
      for (int oi = 0; oi < 2; ++oi) {
        vol[ok][oj][oi] = 0.0;
        terms[ok][oj][oi] = 0.0;
        if ((ok == 0 || inc3) && (oj == 0 || inc2) && (oi == 0 || inc1)) {
          Coords<GEOM> cc(g, block, k + ok, j + oj, i + oi);
          vol[ok][oj][oi] = el == 1 ? cc.area1(cc.x1[0]) : (el == 2 ? cc.area2(0) : cc.area3());
          terms[ok][oj][oi] =
              vol[ok][oj][oi] * fine[((size_t)(k + ok) * snj + (j + oj)) * sni + (i + oi)];
        }
      }
  // restriction.hpp:103-111: off-centred terms first (FP symmetry)
  const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                      ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
  return ddiv((((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
               ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))),
              tvol);
}

// one thread per (descriptor, flux entry, coarse face cell).  DIFF = false: the hydrodynamic
// Metadata::Flux fields of the descriptor's fluid (conserved fluxes + the gas's interface
// pressure, cell-shaped arrays of the bound pack).  DIFF = true: the gas diffusion fluxes
// gas.diff.momentum / gas.diff.energy (src/gas/gas.cpp:277-285, also Metadata::WithFluxes and
// therefore corrected by the same Parthenon task), library-owned face-shaped arrays
// dflx[nb][4 S][fnk][fnj][fni] of the descriptor's direction; dust descriptors have none.
template <int GEOM, bool DIFF>
__global__ void __launch_bounds__(kThreads)
k_flux_correct(GridDev g, GridDev gc, FluidDev f0, FluidDev f1, const FluxCorDev *__restrict__ fc,
               double *__restrict__ dflx0, double *__restrict__ dflx1, double *__restrict__ dflx2) {
  const FluxCorDev d = fc[blockIdx.y];
  if (DIFF && d.fluid != AB200_GAS) return;
  const FluidDev &f = d.fluid == AB200_GAS ? f0 : f1;
  // conserved fluxes + pressure flux | momentum + energy diffusion fluxes
  const int nent = DIFF ? 4 * f.S : f.nvar + (d.fluid == AB200_GAS ? f.S : 0);
  const int nci = d.cie - d.cis + 1, ncj = d.cje - d.cjs + 1, nck = d.cke - d.cks + 1;
  const long long total = (long long)nent * nck * ncj * nci;
  const int el = d.dir + 1;
  const int snj = DIFF ? g.fnj : g.nj, sni = DIFF ? g.fni : g.ni;
  double *dflx = d.dir == 0 ? dflx0 : (d.dir == 1 ? dflx1 : dflx2);
  const size_t fcells = (size_t)g.fnk * g.fnj * g.fni;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int ci = (int)(r % nci) + d.cis; r /= nci;
    const int cj = (int)(r % ncj) + d.cjs; r /= ncj;
    const int ck = (int)(r % nck) + d.cks; r /= nck;
    const int n = (int)r;
    const int i = (ci - gc.is) * 2 + g.is;
    const int j = g.ndim > 1 ? (cj - gc.js) * 2 + g.js : g.js;
    const int k = g.ndim > 2 ? (ck - gc.ks) * 2 + g.ks : g.ks;
    const double *fine;
    double *dst;
    if (DIFF) {
      fine = dflx + ((size_t)d.fine_block * nent + n) * fcells;
      dst = dflx + ((size_t)d.coarse_block * nent + n) * fcells;
    } else {
      double *const *tab = n < f.nvar ? f.flux[d.dir] : f.pflux[d.dir];
      const int ent = n < f.nvar ? n : n - f.nvar;
      const int stride = n < f.nvar ? f.nvar : f.S;
      fine = tab[(size_t)d.fine_block * stride + ent];
      dst = tab[(size_t)d.coarse_block * stride + ent];
    }
    const double v = restrict_face<GEOM>(g, d.fine_block, el, k, j, i, fine, snj, sni);
    dst[((size_t)(d.dsk + (ck - d.cks)) * snj + (d.dsj + (cj - d.cjs))) * sni +
        (d.dsi + (ci - d.cis))] = v;
  }
}

}  // namespace ab200

using namespace ab200;

extern "C" {

int ab200_flux_correct(ab200_ctx *c, const ab200_fluxcor_desc *fc, int nd) {
  AB_REQUIRE(c && c->grid_set, AB200_ESTATE, "ab200_flux_correct: no grid bound");
  if (nd == 0) return AB200_OK;
  AB_REQUIRE(fc && nd > 0, AB200_EINVAL, "ab200_flux_correct: bad descriptor list");
  AB_CUDA(cudaSetDevice(c->device));
  AB_TRY(ensure_coarse_grid(c));
  const GridDev &g = c->g, &gc = c->gc;
  std::vector<FluxCorDev> h(nd);
  long long maxcells = 1;
  for (int q = 0; q < nd; ++q) {
    const ab200_fluxcor_desc &s = fc[q];
    AB_REQUIRE(s.fluid == AB200_GAS || s.fluid == AB200_DUST, AB200_EINVAL, "ab200_flux_correct: bad fluid");
    const FluidHost &fh = c->fl[s.fluid];
    AB_REQUIRE(fh.bound, AB200_ESTATE, "ab200_flux_correct: fluid not bound");
    AB_REQUIRE(s.dir >= 0 && s.dir < g.ndim, AB200_EINVAL, "ab200_flux_correct: bad direction");
    AB_REQUIRE(fh.d.flux[s.dir] && (s.fluid == AB200_DUST || fh.d.pflux[s.dir]), AB200_ESTATE,
               "ab200_flux_correct: no flux arrays (call ab200_calculate_fluxes first)");
    AB_REQUIRE(s.fine_block >= 0 && s.fine_block < g.nb && s.coarse_block >= 0 &&
                   s.coarse_block < g.nb,
               AB200_EINVAL, "ab200_flux_correct: bad block");
    AB_REQUIRE(s.cis >= 0 && s.cie < gc.ni + 1 && s.cjs >= 0 && s.cje < gc.nj + 1 && s.cks >= 0 &&
                   s.cke < gc.nk + 1 && s.cis <= s.cie && s.cjs <= s.cje && s.cks <= s.cke,
               AB200_EINVAL, "ab200_flux_correct: bad coarse box");
    AB_REQUIRE(s.dsi >= 0 && s.dsi + (s.cie - s.cis) < g.ni && s.dsj >= 0 &&
                   s.dsj + (s.cje - s.cjs) < g.nj && s.dsk >= 0 && s.dsk + (s.cke - s.cks) < g.nk,
               AB200_EINVAL, "ab200_flux_correct: destination outside the array");
    FluxCorDev &d = h[q];
    std::memset(&d, 0, sizeof d);
    d.fluid = s.fluid; d.fine_block = s.fine_block; d.coarse_block = s.coarse_block; d.dir = s.dir;
    d.cis = s.cis; d.cie = s.cie; d.cjs = s.cjs; d.cje = s.cje; d.cks = s.cks; d.cke = s.cke;
    d.dsi = s.dsi; d.dsj = s.dsj; d.dsk = s.dsk;
    maxcells = std::max(maxcells, (long long)(s.cie - s.cis + 1) * (s.cje - s.cjs + 1) *
                                      (s.cke - s.cks + 1) * (fh.d.nvar + fh.d.S));
  }
  void *dev = nullptr;
  AB_TRY(cached_descriptors(c, h.data(), sizeof(FluxCorDev) * (size_t)nd, nd, &dev));
  NvtxRange nvtx_("SendBoundBufs<flxcor_send> + SetBounds<flxcor_recv> [fused with restriction]");
  const unsigned gx = (unsigned)std::min<long long>((maxcells + kThreads - 1) / kThreads, 64);
  for (int q0 = 0; q0 < nd; q0 += kMaxGridY) {  // grid.y is a 16-bit dimension
    dim3 grid(gx, (unsigned)std::min(kMaxGridY, nd - q0));
    const FluxCorDev *dd = (const FluxCorDev *)dev + q0;
    // gas.diff.* are Metadata::WithFluxes fields too: once the diffusion operators are
    // configured and their flux arrays exist, the same task corrects them
    const bool diff = c->has_diffusion && c->d_dflx[0] && c->fl[AB200_GAS].bound;
#define AB_LAUNCH(G)                                                                          \
  case G:                                                                                     \
    k_flux_correct<G, false><<<grid, kThreads, 0, c->stream>>>(g, gc, c->fl[0].d, c->fl[1].d, \
                                                               dd, nullptr, nullptr, nullptr); \
    if (diff)                                                                                 \
      k_flux_correct<G, true><<<grid, kThreads, 0, c->stream>>>(                              \
          g, gc, c->fl[0].d, c->fl[1].d, dd, c->d_dflx[0], c->d_dflx[1], c->d_dflx[2]);       \
    break;
    switch (g.geom) {
      AB_LAUNCH(0) AB_LAUNCH(1) AB_LAUNCH(2) AB_LAUNCH(3) AB_LAUNCH(4) AB_LAUNCH(5)
    default:
      set_error("Coordinate type not recognized!");
      return AB200_EINVAL;
    }
#undef AB_LAUNCH
    c->launches += diff ? 2 : 1;
  }
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

int ab200_box_copy(ab200_ctx *c, const ab200_box_desc *bx, int nd) {
  AB_REQUIRE(c && c->grid_set, AB200_ESTATE, "ab200_box_copy: no grid bound");
  if (nd == 0) return AB200_OK;
  AB_REQUIRE(bx && nd > 0, AB200_EINVAL, "ab200_box_copy: bad descriptor list");
  AB_CUDA(cudaSetDevice(c->device));
  AB_TRY(ensure_coarse_grid(c));
  const GridDev &g = c->g, &gc = c->gc;
  std::vector<BoxDev> h(nd);
  long long maxcells = 1;
  for (int q = 0; q < nd; ++q) {
    const ab200_box_desc &s = bx[q];
    AB_REQUIRE(s.fluid == AB200_GAS || s.fluid == AB200_DUST, AB200_EINVAL, "ab200_box_copy: bad fluid");
    AB_REQUIRE(c->fl[s.fluid].bound, AB200_ESTATE, "ab200_box_copy: fluid not bound");
    const FluidDev &f = c->fl[s.fluid].d;
    AB_REQUIRE(s.ncomp >= 1 && s.ni >= 1 && s.nj >= 1 && s.nk >= 1, AB200_EINVAL, "ab200_box_copy: empty box");
    const GridDev &sg = s.src_coarse ? gc : g, &dg = s.dst_coarse ? gc : g;
    AB_REQUIRE(s.ssi >= 0 && s.ssj >= 0 && s.ssk >= 0 && s.ssi + s.ni <= sg.ni &&
                   s.ssj + s.nj <= sg.nj && s.ssk + s.nk <= sg.nk,
               AB200_EINVAL, "ab200_box_copy: source box outside the array");
    AB_REQUIRE(s.dsi >= 0 && s.dsj >= 0 && s.dsk >= 0 && s.dsi + s.ni <= dg.ni &&
                   s.dsj + s.nj <= dg.nj && s.dsk + s.nk <= dg.nk,
               AB200_EINVAL, "ab200_box_copy: destination box outside the array");
    AB_REQUIRE(s.src_coarse || (s.src_block >= 0 && s.src_block < g.nb && s.src_var0 >= 0 &&
                                s.src_var0 + s.ncomp <= f.nvar),
               AB200_EINVAL, "ab200_box_copy: bad source entries");
    AB_REQUIRE(s.dst_coarse || (s.dst_block >= 0 && s.dst_block < g.nb && s.dst_var0 >= 0 &&
                                s.dst_var0 + s.ncomp <= f.nvar),
               AB200_EINVAL, "ab200_box_copy: bad destination entries");
    BoxDev &d = h[q];
    std::memset(&d, 0, sizeof d);
    d.fluid = s.fluid; d.ncomp = s.ncomp;
    d.src_block = s.src_block; d.src_var0 = s.src_var0; d.src_coarse = s.src_coarse;
    d.dst_block = s.dst_block; d.dst_var0 = s.dst_var0; d.dst_coarse = s.dst_coarse;
    d.ssi = s.ssi; d.ssj = s.ssj; d.ssk = s.ssk; d.dsi = s.dsi; d.dsj = s.dsj; d.dsk = s.dsk;
    d.ni = s.ni; d.nj = s.nj; d.nk = s.nk;
    maxcells = std::max(maxcells, (long long)s.ni * s.nj * s.nk * s.ncomp);
  }
  void *dev = nullptr;
  AB_TRY(cached_descriptors(c, h.data(), sizeof(BoxDev) * (size_t)nd, nd, &dev));
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(sync_prim_home(c, f, 0));
  NvtxRange nvtx_("SendBoundBufs + SetBounds [multilevel, same GPU]");
  const unsigned gx = (unsigned)std::min<long long>((maxcells + kThreads - 1) / kThreads, 64);
  for (int q0 = 0; q0 < nd; q0 += kMaxGridY) {  // grid.y is a 16-bit dimension
    dim3 grid(gx, (unsigned)std::min(kMaxGridY, nd - q0));
    k_box_copy<<<grid, kThreads, 0, c->stream>>>(g, gc, c->fl[0].d, c->fl[1].d,
                                                 (const BoxDev *)dev + q0);
    c->launches++;
  }
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

int ab200_block_bcs(ab200_ctx *c, const ab200_block_bc_desc *bc, int nd) {
  AB_REQUIRE(c && c->grid_set, AB200_ESTATE, "ab200_block_bcs: no grid bound");
  if (nd == 0) return AB200_OK;
  AB_REQUIRE(bc && nd > 0, AB200_EINVAL, "ab200_block_bcs: bad descriptor list");
  AB_CUDA(cudaSetDevice(c->device));
  // the coarse index space (and its even-block-size requirement) matters only to descriptors
  // that name a coarse buffer; a uniform mesh applies its user conditions on fine arrays alone
  for (int q = 0; q < nd; ++q)
    if (bc[q].coarse || bc[q].coarse_entries) { AB_TRY(ensure_coarse_grid(c)); break; }
  const GridDev &g = c->g, &gc = c->gc;
  std::vector<BcDev> h(nd);
  bool has_dir[3] = {false, false, false}, user_dir[3] = {false, false, false},
       generic_dir[3] = {false, false, false};
  for (int q = 0; q < nd; ++q) {
    const ab200_block_bc_desc &s = bc[q];
    AB_REQUIRE(s.fluid == AB200_GAS || s.fluid == AB200_DUST, AB200_EINVAL, "ab200_block_bcs: bad fluid");
    AB_REQUIRE(c->fl[s.fluid].bound, AB200_ESTATE, "ab200_block_bcs: fluid not bound");
    const FluidDev &f = c->fl[s.fluid].d;
    AB_REQUIRE(s.face >= 0 && s.face < 2 * g.ndim, AB200_EINVAL, "ab200_block_bcs: bad face");
    const bool user = s.type == AB200_BC_EXTRAP || s.type == AB200_BC_INFLOW;
    AB_REQUIRE(s.type == AB200_BC_OUTFLOW || s.type == AB200_BC_REFLECT || user, AB200_EINVAL,
               "ab200_block_bcs: outflow, reflect, extrap or inflow");
    // strat.hpp registers `extrap` on the x1 and x3 faces and `inflow` on the x2 faces only
    // (src/pgen/problem_modifier.hpp:114-127); both act on the whole fluid at once
    AB_REQUIRE(s.type != AB200_BC_EXTRAP || s.face / 2 != 1, AB200_EINVAL,
               "ab200_block_bcs: AB200_BC_EXTRAP exists on x1 and x3 faces only");
    AB_REQUIRE(s.type != AB200_BC_INFLOW || s.face / 2 == 1, AB200_EINVAL,
               "ab200_block_bcs: AB200_BC_INFLOW exists on x2 faces only");
    AB_REQUIRE(!user || (s.var0 == 0 && s.ncomp == f.nvar), AB200_EINVAL,
               "ab200_block_bcs: a user condition takes every pack entry of the fluid");
    AB_REQUIRE(user || !s.coarse_entries, AB200_EINVAL,
               "ab200_block_bcs: coarse_entries is for user conditions (outflow / reflect take one "
               "descriptor per Variable)");
    AB_REQUIRE(s.type != AB200_BC_INFLOW || c->shear_bc_set, AB200_ESTATE,
               "ab200_block_bcs: call ab200_set_shear_bc_params before AB200_BC_INFLOW");
    AB_REQUIRE(s.block >= 0 && s.block < g.nb && s.var0 >= 0 && s.ncomp >= 1 &&
                   s.var0 + s.ncomp <= f.nvar,
               AB200_EINVAL, "ab200_block_bcs: bad entries");
    BcDev &d = h[q];
    std::memset(&d, 0, sizeof d);
    d.fluid = s.fluid; d.block = s.block; d.var0 = s.var0; d.ncomp = s.ncomp;
    d.face = s.face; d.type = s.type; d.coarse = s.coarse; d.ctab = s.coarse_entries;
    has_dir[s.face / 2] = true;
    (user ? user_dir : generic_dir)[s.face / 2] = true;
  }
  void *dev = nullptr;
  AB_TRY(cached_descriptors(c, h.data(), sizeof(BcDev) * (size_t)nd, nd, &dev));
  for (int f = 0; f < 2; ++f)
    if (c->fl[f].bound) AB_TRY(sync_prim_home(c, f, 0));
  NvtxRange nvtx_("ApplyBoundaryConditionsOnCoarseOrFineMD [per block]");
  const long long plane = (long long)std::max(g.ni * g.nj, std::max(g.nj * g.nk, g.ni * g.nk));
  const unsigned gx = (unsigned)std::min<long long>((plane * g.ng * 6 + kThreads - 1) / kThreads, 64);
  for (int dir = 0; dir < g.ndim; ++dir) {
    if (!has_dir[dir]) continue;
    for (int q0 = 0; q0 < nd; q0 += kMaxGridY) {  // grid.y is a 16-bit dimension
      dim3 grid(gx, (unsigned)std::min(kMaxGridY, nd - q0));
      // faces of one direction never overlap, so the two kernels of a direction commute
      if (generic_dir[dir]) {
        k_block_bcs<<<grid, kThreads, 0, c->stream>>>(g, gc, c->fl[0].d, c->fl[1].d,
                                                      (const BcDev *)dev + q0, dir);
        c->launches++;
      }
      if (user_dir[dir]) {
        k_block_user_bcs<<<grid, kThreads, 0, c->stream>>>(g, gc, c->fl[0].d, c->fl[1].d,
                                                           (const BcDev *)dev + q0, dir,
                                                           c->shear_q, c->shear_om0);
        c->launches++;
      }
    }
  }
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}


int ab200_set_shear_bc_params(ab200_ctx *c, double q, double om0) {
  AB_REQUIRE(c, AB200_EINVAL, "ab200_set_shear_bc_params: null context");
  c->shear_q = q;
  c->shear_om0 = om0;
  c->shear_bc_set = true;
  return AB200_OK;
}
}  // extern "C"
