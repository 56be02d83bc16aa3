// refine.cu -- the multilevel operators of the ghost exchange:
//   ab200_restrict    ArtemisUtils::RestrictAverage<GEOM>::Do<DIM, CC>
//                     (src/utils/refinement/restriction.hpp:41-114)
//   ab200_prolongate  ArtemisUtils::ProlongateSharedMinMod<GEOM>::Do<DIM, CC>
//                     (src/utils/refinement/prolongation.hpp:82-184, GetGridSpacings :39-67,
//                     GradMinMod :72-79)
// applied over descriptor lists (the analogue of the index ranges Parthenon's
// refinement::Restrict / Prolongate take from BndInfo, P:prolong_restrict/prolong_restrict.hpp).
// One thread per (descriptor, variable, coarse cell); grid.y = descriptor.  Both operators are
// pure streaming (restriction reads 2^ndim fine cells per coarse cell, prolongation writes
// them), so the bound is HBM; with the few ghost layers they run over they are launch-latency
// bound in practice, which is why a whole descriptor list goes into ONE launch.
//
// Geometry: the fine cells use the bound grid's metric tables; the coarse buffer of a block is
// a UniformCartesian of twice the cell width with the same number of ghost cells
// (P:coordinates/uniform_cartesian.hpp:41-55, P:mesh/meshblock.cpp:205-228) whose tables are
// built on first use.  The arithmetic follows the reference operation for operation (the
// strict build is bit-identical to it, including the pairing of the eight-term sums).
#include <cstring>

#include "ab200_ctx.cuh"
#include "tasks.cuh"

namespace ab200 {

struct RefineDev {
  int fluid, block, var0, nvar, kind;
  int cis, cie, cjs, cje, cks, cke;
  double *coarse;
};

int ensure_coarse_grid(ab200_ctx *c) {
  if (c->coarse_ready) return AB200_OK;
  const GridDev &g = c->g;
  GridDev &gc = c->gc;
  gc = g;
  const int act[3] = {1, g.ndim > 1, g.ndim > 2};
  const int nint[3] = {g.ie - g.is + 1, g.je - g.js + 1, g.ke - g.ks + 1};
  int cn[3], cs[3];
  for (int d = 0; d < 3; ++d) {
    AB_REQUIRE(!act[d] || nint[d] % 2 == 0, AB200_EINVAL,
               "multilevel operators need an even number of interior cells per active direction");
    cn[d] = act[d] ? nint[d] / 2 + 2 * g.ng : 1;  // P:mesh/meshblock.cpp:205-228
    cs[d] = act[d] ? g.ng : 0;
  }
  gc.ni = cn[0]; gc.nj = cn[1]; gc.nk = cn[2];
  gc.is = cs[0]; gc.js = cs[1]; gc.ks = cs[2];
  gc.ie = cs[0] + (act[0] ? nint[0] / 2 : 1) - 1;
  gc.je = cs[1] + (act[1] ? nint[1] / 2 : 1) - 1;
  gc.ke = cs[2] + (act[2] ? nint[2] / 2 : 1) - 1;
  std::vector<double> cx(3 * (size_t)g.nb), cd(3 * (size_t)g.nb);
  for (int b = 0; b < g.nb; ++b)
    for (int d = 0; d < 3; ++d) {
      // UniformCartesian(const UniformCartesian &src, int coarsen = 2)
      const int istart = act[d] ? g.ng : 0;
      const int coarsen = 2;
      double dx = c->h_dx[3 * b + d];
      double xm = c->h_xmin[3 * b + d];
      xm += istart * dx * (1 - coarsen);
      dx *= (d == 0 ? coarsen : (istart > 0 ? coarsen : 1));
      cx[3 * b + d] = xm;
      cd[3 * b + d] = dx;
    }
  AB_TRY(build_geom_tables_for(c, gc.t, g.geom, g.nb, gc.ni, gc.nj, gc.nk, cx.data(), cd.data()));
  c->coarse_ready = true;
  return AB200_OK;
}

template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_restrict(GridDev g, GridDev gc, FluidDev f0, FluidDev f1, const RefineDev *__restrict__ descs) {
  const RefineDev d = descs[blockIdx.y];
  const FluidDev &f = d.fluid == AB200_GAS ? f0 : f1;
  double *const *tab = d.kind == 0 ? f.prim : f.u0;
  const int nci = d.cie - d.cis + 1, ncj = d.cje - d.cjs + 1, nck = d.cke - d.cks + 1;
  const long long total = (long long)d.nvar * nck * ncj * nci;
  const bool inc2 = g.ndim > 1, inc3 = g.ndim > 2;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int ci = (int)(r % nci) + d.cis; r /= nci;
    const int cj = (int)(r % ncj) + d.cjs; r /= ncj;
    const int ck = (int)(r % nck) + d.cks; r /= nck;
    const int n = (int)r;
    const int i = (ci - gc.is) * 2 + g.is;
    const int j = inc2 ? (cj - gc.js) * 2 + g.js : g.js;
    const int k = inc3 ? (ck - gc.ks) * 2 + g.ks : g.ks;
    const double *fine = tab[(size_t)d.block * f.nvar + d.var0 + n];
    double vol[2][2][2], terms[2][2][2];
#pragma unroll
    for (int ok = 0; ok < 2; ++ok)
#pragma unroll
      for (int oj = 0; oj < 2; ++oj)
#pragma unroll
        for (int oi = 0; oi < 2; ++oi) {
          vol[ok][oj][oi] = 0.0;
          terms[ok][oj][oi] = 0.0;
          if ((ok == 0 || inc3) && (oj == 0 || inc2)) {
            Coords<GEOM> cc(g, d.block, k + ok, j + oj, i + oi);
            vol[ok][oj][oi] = cc.volume();
            terms[ok][oj][oi] =
                vol[ok][oj][oi] * fine[((size_t)(k + ok) * g.nj + (j + oj)) * g.ni + (i + oi)];
          }
        }
    // restriction.hpp:103-111: off-centred terms first (FP symmetry)
    const double tvol = ((vol[0][0][0] + vol[0][1][0]) + (vol[0][0][1] + vol[0][1][1])) +
                        ((vol[1][0][0] + vol[1][1][0]) + (vol[1][0][1] + vol[1][1][1]));
    d.coarse[(((size_t)n * gc.nk + ck) * gc.nj + cj) * gc.ni + ci] =
        ddiv((((terms[0][0][0] + terms[0][1][0]) + (terms[0][0][1] + terms[0][1][1])) +
              ((terms[1][0][0] + terms[1][1][0]) + (terms[1][0][1] + terms[1][1][1]))),
             tvol);
  }
}

// prolongation.hpp:72-79 (SIGN: P:config.hpp.in:86)
AB_D double grad_minmod(double fc, double fm, double fp, double dxm, double dxp) {
  const double gxm = ddiv(fc - fm, dxm);
  const double gxp = ddiv(fp - fc, dxp);
  const double sm = (gxm < 0.0) ? -1.0 : 1.0, sp = (gxp < 0.0) ? -1.0 : 1.0;
  return 0.5 * (sm + sp) * dmin(fabs(gxm), fabs(gxp));
}

// prolongation.hpp:39-67: centroid distances along DIM on both levels.  The centroids come from
// the host-built metric tables (the reference's own formulas evaluated without FMA contraction),
// so the differences below -- which cancel most of their leading digits -- are the reference's
// in both builds.
template <int GEOM, int DIM>
AB_D void grid_spacings(const GridDev &g, const GridDev &gc, int b, int k, int j, int i, int fk,
                        int fj, int fi, double &dxm, double &dxp, double &dxfm, double &dxfp) {
  const double *cv, *fv;
  int c, fc;
  if (DIM == 1) { cv = gc.t.x1v + (size_t)b * gc.ni; fv = g.t.x1v + (size_t)b * g.ni; c = i; fc = fi; }
  else if (DIM == 2) { cv = gc.t.x2v + (size_t)b * gc.nj; fv = g.t.x2v + (size_t)b * g.nj; c = j; fc = fj; }
  else { cv = gc.t.x3v + (size_t)b * gc.nk; fv = g.t.x3v + (size_t)b * g.nk; c = k; fc = fk; }
  const double xm = cv[c - 1], xc = cv[c], xp = cv[c + 1];
  const double fxm = fv[fc], fxp = fv[fc + 1];
  dxm = xc - xm;
  dxp = xp - xc;
  dxfm = xc - fxm;
  dxfp = fxp - xc;
}

template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_prolongate(GridDev g, GridDev gc, FluidDev f0, FluidDev f1, const RefineDev *__restrict__ descs) {
  const RefineDev d = descs[blockIdx.y];
  const FluidDev &f = d.fluid == AB200_GAS ? f0 : f1;
  double *const *tab = d.kind == 0 ? f.prim : f.u0;
  const int nci = d.cie - d.cis + 1, ncj = d.cje - d.cjs + 1, nck = d.cke - d.cks + 1;
  const long long total = (long long)d.nvar * nck * ncj * nci;
  const bool inc2 = g.ndim > 1, inc3 = g.ndim > 2;
  const size_t sj = gc.ni, sk = (size_t)gc.ni * gc.nj;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    long long r = t;
    const int i = (int)(r % nci) + d.cis; r /= nci;
    const int j = (int)(r % ncj) + d.cjs; r /= ncj;
    const int k = (int)(r % nck) + d.cks; r /= nck;
    const int n = (int)r;
    const int fi = (i - gc.is) * 2 + g.is;
    const int fj = inc2 ? (j - gc.js) * 2 + g.js : g.js;
    const int fk = inc3 ? (k - gc.ks) * 2 + g.ks : g.ks;
    const double *co = d.coarse + (((size_t)n * gc.nk + k) * gc.nj + j) * gc.ni + i;
    double *fine = tab[(size_t)d.block * f.nvar + d.var0 + n];
    const double fc = co[0];
    double dx1fm = 0, dx1fp = 0, gx1m = 0, gx1p = 0;
    {
      double dx1m, dx1p;
      grid_spacings<GEOM, 1>(g, gc, d.block, k, j, i, fk, fj, fi, dx1m, dx1p, dx1fm, dx1fp);
      const double gx = grad_minmod(fc, co[-1], co[1], dx1m, dx1p);
      gx1m = gx; gx1p = gx;
    }
    double dx2fm = 0, dx2fp = 0, gx2m = 0, gx2p = 0;
    if (inc2) {
      double dx2m, dx2p;
      grid_spacings<GEOM, 2>(g, gc, d.block, k, j, i, fk, fj, fi, dx2m, dx2p, dx2fm, dx2fp);
      const double gx = grad_minmod(fc, co[-(long long)sj], co[sj], dx2m, dx2p);
      gx2m = gx; gx2p = gx;
    }
    double dx3fm = 0, dx3fp = 0, gx3m = 0, gx3p = 0;
    if (inc3) {
      double dx3m, dx3p;
      grid_spacings<GEOM, 3>(g, gc, d.block, k, j, i, fk, fj, fi, dx3m, dx3p, dx3fm, dx3fp);
      const double gx = grad_minmod(fc, co[-(long long)sk], co[sk], dx3m, dx3p);
      gx3m = gx; gx3p = gx;
    }
    auto F = [&](int kk, int jj, int ii) -> double & {
      return fine[((size_t)kk * g.nj + jj) * g.ni + ii];
    };
    // prolongation.hpp:158-181: off-centred terms first (FP symmetry); unused terms are zero
    F(fk, fj, fi) = fc - (gx1m * dx1fm + gx2m * dx2fm + gx3m * dx3fm);
    F(fk, fj, fi + 1) = fc + (gx1p * dx1fp - gx2m * dx2fm - gx3m * dx3fm);
    if (inc2) {
      F(fk, fj + 1, fi) = fc - (gx1m * dx1fm - gx2p * dx2fp + gx3m * dx3fm);
      F(fk, fj + 1, fi + 1) = fc + (gx1p * dx1fp + gx2p * dx2fp - gx3m * dx3fm);
    }
    if (inc3) {
      F(fk + 1, fj, fi) = fc - (gx1m * dx1fm + gx2m * dx2fm - gx3p * dx3fp);
      F(fk + 1, fj, fi + 1) = fc + (gx1p * dx1fp - gx2m * dx2fm + gx3p * dx3fp);
      if (inc2) {
        F(fk + 1, fj + 1, fi) = fc - (gx1m * dx1fm - gx2p * dx2fp - gx3p * dx3fp);
        F(fk + 1, fj + 1, fi + 1) = fc + (gx1p * dx1fp + gx2p * dx2fp + gx3p * dx3fp);
      }
    }
  }
}

static int launch_refine(ab200_ctx *c, const ab200_refine_desc *descs, int n, bool prolong) {
  AB_REQUIRE(c && (descs || n == 0), AB200_EINVAL, "refine: null argument");
  AB_REQUIRE(c->grid_set, AB200_ESTATE, "no grid bound: call ab200_set_grid");
  if (n == 0) return AB200_OK;
  AB_CUDA(cudaSetDevice(c->device));
  AB_TRY(ensure_coarse_grid(c));
  for (int f = 0; f < 2; ++f) AB_TRY(sync_prim_home(c, f, 0));  // operators use the caller's arrays
  const GridDev &g = c->g, &gc = c->gc;
  const int act[3] = {1, g.ndim > 1, g.ndim > 2};
  const int grow = prolong ? 1 : 0;  // the minmod stencil reads one coarse neighbour per side
  std::vector<RefineDev> h(n);
  memset(h.data(), 0, sizeof(RefineDev) * (size_t)n);
  long long maxel = 0;
  for (int q = 0; q < n; ++q) {
    const ab200_refine_desc &d = descs[q];
    AB_REQUIRE(d.fluid == 0 || d.fluid == 1, AB200_EINVAL, "Fluid type not recognized!");
    AB_REQUIRE(c->fl[d.fluid].bound, AB200_ESTATE, "refine: descriptor names an unbound fluid");
    AB_REQUIRE(d.kind == AB200_REFINE_PRIM || d.kind == AB200_REFINE_CONS, AB200_EINVAL,
               "refine: unknown array kind");
    AB_REQUIRE(d.block >= 0 && d.block < g.nb && d.var0 >= 0 && d.nvar > 0 &&
                   d.var0 + d.nvar <= c->fl[d.fluid].d.nvar,
               AB200_EINVAL, "refine: descriptor block/variable out of range");
    AB_REQUIRE(d.coarse != nullptr, AB200_EINVAL, "refine: null coarse buffer");
    const int lo[3] = {d.cis, d.cjs, d.cks}, hi[3] = {d.cie, d.cje, d.cke};
    const int cn[3] = {gc.ni, gc.nj, gc.nk}, cs[3] = {gc.is, gc.js, gc.ks};
    const int fs[3] = {g.is, g.js, g.ks}, fn[3] = {g.ni, g.nj, g.nk};
    for (int a = 0; a < 3; ++a) {
      AB_REQUIRE(lo[a] <= hi[a], AB200_EINVAL, "refine: empty coarse index box");
      if (!act[a]) {
        AB_REQUIRE(lo[a] == 0 && hi[a] == 0, AB200_EINVAL,
                   "refine: coarse box must be {0,0} along an inactive direction");
        continue;
      }
      AB_REQUIRE(lo[a] - grow >= 0 && hi[a] + grow < cn[a], AB200_EINVAL,
                 "refine: coarse index box (plus the stencil) outside the coarse buffer");
      AB_REQUIRE((lo[a] - cs[a]) * 2 + fs[a] >= 0 && (hi[a] - cs[a]) * 2 + fs[a] + 1 < fn[a],
                 AB200_EINVAL, "refine: fine cells of the coarse box fall outside the block");
    }
    // member-wise into zero-filled storage: the padding bytes are part of the cache key
    RefineDev &o = h[q];
    o.fluid = d.fluid; o.block = d.block; o.var0 = d.var0; o.nvar = d.nvar; o.kind = d.kind;
    o.cis = d.cis; o.cie = d.cie; o.cjs = d.cjs; o.cje = d.cje; o.cks = d.cks; o.cke = d.cke;
    o.coarse = d.coarse;
    const long long el = (long long)d.nvar * (d.cie - d.cis + 1) * (d.cje - d.cjs + 1) *
                         (d.cke - d.cks + 1);
    if (el > maxel) maxel = el;
  }
  // descriptor lists are static between remeshes: cached on the device, keyed by content
  RefineDev *dd = nullptr;
  AB_TRY(cached_descriptors(c, h.data(), sizeof(RefineDev) * (size_t)n, n, (void **)&dd));
  unsigned gx = (unsigned)((maxel + kThreads - 1) / kThreads);
  if (gx > 1024) gx = 1024;
  dim3 grid(gx, (unsigned)n);
  const FluidDev &f0 = c->fl[0].d, &f1 = c->fl[1].d;
#define AB_LAUNCH(G)                                                                            \
  case G:                                                                                       \
    if (prolong) k_prolongate<G><<<grid, kThreads, 0, c->stream>>>(g, gc, f0, f1, dd);          \
    else k_restrict<G><<<grid, kThreads, 0, c->stream>>>(g, gc, f0, f1, dd);                    \
    break;
  switch (g.geom) {
    AB_LAUNCH(0) AB_LAUNCH(1) AB_LAUNCH(2) AB_LAUNCH(3) AB_LAUNCH(4) AB_LAUNCH(5)
  default:
    set_error("Coordinate type not recognized!");
    return AB200_EINVAL;
  }
#undef AB_LAUNCH
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

}  // namespace ab200

using namespace ab200;

extern "C" {

int ab200_coarse_shape(ab200_ctx *c, int *dims6) {
  AB_REQUIRE(c && dims6, AB200_EINVAL, "ab200_coarse_shape: null argument");
  AB_REQUIRE(c->grid_set, AB200_ESTATE, "no grid bound: call ab200_set_grid");
  AB_CUDA(cudaSetDevice(c->device));
  AB_TRY(ensure_coarse_grid(c));
  dims6[0] = c->gc.ni; dims6[1] = c->gc.nj; dims6[2] = c->gc.nk;
  dims6[3] = c->gc.is; dims6[4] = c->gc.js; dims6[5] = c->gc.ks;
  return AB200_OK;
}

int ab200_restrict(ab200_ctx *c, const ab200_refine_desc *descs, int n) {
  return launch_refine(c, descs, n, false);
}

int ab200_prolongate(ab200_ctx *c, const ab200_refine_desc *descs, int n) {
  return launch_refine(c, descs, n, true);
}

}  // extern "C"
