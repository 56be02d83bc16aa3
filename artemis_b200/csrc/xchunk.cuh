// xchunk.cuh -- the x1 directional pass of the fused stage as a WARP-AUTONOMOUS streaming kernel.
//
// Along x1 the stencil neighbours of a cell live in adjacent lanes, so the pass cannot march
// with a private register window like march.cuh does for x2/x3.  Instead one warp streams
// through the cells of one k-plane of a MeshBlock in memory order (rows js..je are contiguous:
// 64 x 72 cells = 144 chunks of 32), lane l of chunk t holding flat cell g = 32 t + l, and the
// reconstruct -> Riemann -> update chain is software-pipelined ACROSS LANES with a lag of one
// cell per stage:
//
//   S1  I(g-1)   = PPM interface between cells g-2 and g-1 from q(g-3..g)      (ppm.hpp:39-46)
//   S2  cell g-2 : monotonise with I(g-2) [lane l-1] and I(g-1) -> edges       (ppm.hpp:48-61)
//   S3  face g-2 : Riemann(upper edge of g-3 [lane l-1], lower edge of g-2) -> F(g-2)
//   S4  cell g-3 : update with F(g-3) [lane l-1] and F(g-2)
//
// "lane l-1" values travel through a 64-entry per-warp ring in shared memory (lane 0 reads what
// lane 31 left there in the previous chunk); the stages are separated by __syncwarp() only --
// no CTA barrier, no TMA, no halo recomputation: every interface value, monotonisation and
// Riemann solve of a row is done exactly once, and rows never interact because the ghost
// columns (i < is, i > ie) separate them; results that belong to ghost columns are dropped.
// Loads and stores are 256-byte contiguous per warp.  This is the FIRST pass of a stage, so it
// also applies the low-storage combination gam0*u0 + gam1*u1 (or the folded u1 <- u0 copy) and
// the curvilinear source terms: fluid_fluxes.hpp:107-126, artemis_integrator.hpp:95-106,
// fluid_fluxes.hpp:365-415.
#pragma once
#include "march.cuh"

namespace ab200 {

constexpr int kXcWarps = 4;  // warps per CTA (independent of each other)

// (r02: forcing 4 CTAs per SM -- 128 registers, 16 warps -- was measured 10 % SLOWER per cycle
// than the 160 registers / 12 warps the compiler picks: the pass lives on per-thread ILP)
template <int GEOM, int FLUID, int RS, int RC>
__global__ void __launch_bounds__(kXcWarps * 32)
k_xchunk_pass(GridDev g, FluidDev f, FusedArgs a) {
  constexpr int DIR = 1;
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr bool CART = (GEOM == AB200_CARTESIAN);
  constexpr bool PPM = (RC == AB200_PPM);
  constexpr int NV = gas ? 6 : 4;
  constexpr int NF = gas ? 8 : 4;
#if defined(AB200_FAST_MATH)
  constexpr bool HOIST = CART;
#else
  constexpr bool HOIST = false;
#endif
  extern __shared__ __align__(16) double xc_smem[];
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  double *sI = xc_smem + (size_t)wid * (2 * NV + NF) * 64;  // [NV][64] interface values
  double *sQ = sI + NV * 64;                                // [NV][64] upper-edge states
  double *sF = sQ + NV * 64;                                // [NF][64] face fluxes

  const int njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int S = f.S, nvar = f.nvar;
  // work item of this warp: (block, species, plane, segment of the plane's rows)
  const int nseg = a.np;  // segments per plane
  int w = blockIdx.x * kXcWarps + wid;
  if (w >= (a.blist ? a.nbl : g.nb) * S * nkr * nseg) return;
  const int seg = w % nseg; w /= nseg;
  const int k = g.ks + w % nkr; w /= nkr;
  const int n = w % S;
  const int b = a.blist ? a.blist[w / S] : w / S;
  const int rows_per_seg = (njr + nseg - 1) / nseg;
  const int j0 = g.js + seg * rows_per_seg;
  const int j1 = min(j0 + rows_per_seg, g.je + 1);  // exclusive
  if (j0 >= j1) return;
  const int ni = g.ni;
  const int gbeg = (k * g.nj + j0) * ni;          // first flat cell of the range
  const int ncell = (j1 - j0) * ni;               // cells in the range (whole rows)
  const int nchunk = (ncell + 3 + 31) / 32;       // +3: the pipeline lag
  const int gmax = g.ni * g.nj * g.nk - 1;

  const double dt = a.dt_dev ? *a.dt_dev : a.dt;
  const double bdt = a.beta * dt;
  const EosConsts eos{f.gm1, f.igm1, f.gamma, f.alpha};
  const bool need_u1 = !a.copy_u1;
  const bool need_u0 = a.copy_u1 || a.gam0 != 0.0;

  // reconstruction order for DIR = 1 is the pack order
  const int ci[6] = {n, S + 3 * n, S + 3 * n + 1, S + 3 * n + 2, 4 * S + n, 5 * S + n};
  const double *pq[NV];
  double *pu[NV], *pv[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    pq[v] = f.prim[(size_t)b * nvar + ci[v]];
    pu[v] = f.u0[(size_t)b * nvar + ci[v]];
    pv[v] = f.u1[(size_t)b * nvar + ci[v]];
  }
  double rinv = 0.0;
  if (HOIST) {
    const double *xf = g.t.x1f + (size_t)b * (g.ni + 1);
    rinv = ddiv(bdt, xf[g.is + 1] - xf[g.is]);
  }
  auto clampg = [&](int x) { return x < 0 ? 0 : (x > gmax ? gmax : x); };

  // column index of flat cell (gbeg + lane - 3), kept incrementally (ni need not divide 32)
  int ic = (lane - 3) % ni, jc = j0 + (lane - 3) / ni;
  if (ic < 0) { ic += ni; jc -= 1; }
  const int step_i = 32 % ni, step_j = 32 / ni;

  double qn[NV];  // q(g) of the next chunk, loaded one chunk ahead
#pragma unroll
  for (int v = 0; v < NV; ++v) qn[v] = __ldg(pq[v] + clampg(gbeg + lane));

  for (int t = 0; t < nchunk; ++t) {
    const int gl = gbeg + 32 * t + lane;  // this lane's newest cell
    const int p = ((t & 1) << 5) | lane;  // ring position
    const int pm1 = (p - 1) & 63;
    double q0[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) q0[v] = qn[v];
    if (t + 1 < nchunk) {
#pragma unroll
      for (int v = 0; v < NV; ++v) qn[v] = __ldg(pq[v] + clampg(gl + 32));
#pragma unroll
      for (int v = 0; v < NV; ++v) prefetch_l2(pq[v] + clampg(gl + 32 * 6));
    }
    // cell to update in S4 and its conserved inputs (issued early)
    const int gu = gl - 3;
    const bool upd = (gu >= gbeg) && (gu < gbeg + ncell) && (ic >= g.is) && (ic <= g.ie);
    double v0[NV], v1[NV];
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      v0[m] = (upd && need_u0) ? __ldg(pu[m] + gu) : 0.0;
      v1[m] = (upd && need_u1) ? __ldg(pv[m] + gu) : 0.0;
    }
    // ---- S1/S2: edges of cell g-2 ------------------------------------------------------------
    double qm2[NV], qr[NV], ql[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) qm2[v] = __ldg(pq[v] + clampg(gl - 2));
    if (PPM) {
      double Iup[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double qm3 = __ldg(pq[v] + clampg(gl - 3));
        const double qm1 = __ldg(pq[v] + clampg(gl - 1));
        Iup[v] = ppm_iface(qm3, qm2[v], qm1, q0[v]);
        sI[v * 64 + p] = Iup[v];
      }
      __syncwarp();
#pragma unroll
      for (int v = 0; v < NV; ++v) ppm_mono(sI[v * 64 + pm1], qm2[v], Iup[v], ql[v], qr[v]);
    } else if (RC == AB200_PLM) {
      double gx[6] = {0, 0, 0, 0, 0, 0};
      if (!CART) {
        int i2 = ic + 1;  // column of cell g-2
        int j2 = jc;
        if (i2 >= ni) { i2 -= ni; j2 += 1; }
        i2 = i2 < 1 ? 1 : (i2 > ni - 2 ? ni - 2 : i2);  // ghost-column results are dropped
        plmg_geom<GEOM, DIR>(g, b, k, max(0, min(j2, g.nj - 1)), i2, gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double qm3 = __ldg(pq[v] + clampg(gl - 3));
        const double qm1 = __ldg(pq[v] + clampg(gl - 1));
        if (CART) plm(qm3, qm2[v], qm1, ql[v], qr[v]);
        else plm_g(qm3, qm2[v], qm1, ql[v], qr[v], gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);
      }
    } else {
#pragma unroll
      for (int v = 0; v < NV; ++v) { ql[v] = qm2[v]; qr[v] = qm2[v]; }
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) sQ[v * 64 + p] = ql[v];
    __syncwarp();
    // ---- S3: Riemann at the lower face of cell g-2 --------------------------------------------
    double wl[NV], lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int v = 0; v < NV; ++v) wl[v] = sQ[v * 64 + pm1];
    Riemann<RS, FLUID>::solve(eos, wl, qr, lo);
    if (!CART) {  // ScaleMomentumFlux, fluid_fluxes.hpp:32-70
      int i2 = ic + 1, j2 = jc;
      if (i2 >= ni) { i2 -= ni; j2 += 1; }
      Coords<GEOM> cf(g, b, k, max(0, min(j2, g.nj - 1)), i2);
      double hs[3];
      cf.template face_scale<DIR>(hs);
#pragma unroll
      for (int m = 1; m <= 3; ++m) lo[m] *= hs[m - 1];
    }
#pragma unroll
    for (int m = 0; m < NF; ++m) sF[m * 64 + p] = lo[m];
    __syncwarp();
    // ---- S4: update cell g-3 (lower face F(g-3) from lane l-1, upper face F(g-2) own) -----------
    if (upd) {
      double fl[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int m = 0; m < NF; ++m) fl[m] = sF[m * 64 + pm1];
      const double *hi = lo;
      if (!CART && a.tap) {  // density-flux tap: lower face of every interior cell + the face at ie+1
        double *pf = f.dflux[0][(size_t)b * S + n];
        __stcg(pf + gu, fl[0]);
        if (ic == g.ie) __stcg(pf + gu + 1, hi[0]);
      }
      double u[6];
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        if (a.copy_u1) {  // stage 1 with DeepCopyConservedData folded in: u1 <- u0
          __stcg(pv[m] + gu, v0[m]);
          u[m] = v0[m];
        } else {
          u[m] = (a.gam0 == 0.0) ? a.gam1 * v1[m] : a.gam0 * v0[m] + a.gam1 * v1[m];
        }
      }
      Coords<GEOM> cc(g, b, k, jc, ic);
      if (HOIST) {
#pragma unroll
        for (int m = 0; m < NV; ++m) u[m] += (fl[m] - hi[m]) * rinv;
        if (gas) {
          u[1] += rinv * (fl[6] - hi[6]);
          u[5] -= rinv * 0.5 * (fl[6] + hi[6]) * (hi[7] - fl[7]);
        }
      } else {
        const double a0 = cc.area1(cc.x1[0]), a1 = cc.area1(cc.x1[1]);
        const double vol = cc.volume();
#ifdef AB200_FAST_MATH
        const double wv = ddiv(bdt, vol);
#define AB_UPD(x) ((x) * wv)
#else
#define AB_UPD(x) ((x) * bdt / vol)
#endif
#pragma unroll
        for (int m = 0; m < NV; ++m) u[m] += AB_UPD(a0 * fl[m] - a1 * hi[m]);
        if (gas) {  // FluxSource, direction 1 (fluid_fluxes.hpp:365-392)
          const double dxd = cc.x1[1] - cc.x1[0];
          u[1] += ddiv(bdt, dxd) * (fl[6] - hi[6]);
#ifdef AB200_FAST_MATH
          u[5] -= wv * 0.5 * (fl[6] + hi[6]) * (a1 * hi[7] - a0 * fl[7]);
#else
          u[5] -= bdt / vol * 0.5 * (fl[6] + hi[6]) * (a1 * hi[7] - a0 * fl[7]);
#endif
        }
#undef AB_UPD
        // coordinate source terms (fluid_fluxes.hpp:395-415), added once per stage
        const double wc0 = __ldg(pq[0] + gu);
        const double vel[3] = {__ldg(pq[1] + gu), __ldg(pq[2] + gu), __ldg(pq[3] + gu)};
        double vf[3];
        cc.rotation_velocity(a.omf, vf);
        const double rdt = wc0 * bdt;
        const double s0q = sqr(vel[0] + vf[0]), s1q = sqr(vel[1] + vf[1]),
                     s2q = sqr(vel[2] + vf[2]);
        if (Coords<GEOM>::x1dep) {
          double dh[3];
          cc.conn1(dh);
          u[1] += rdt * (dh[0] * s0q + dh[1] * s1q + dh[2] * s2q);
        }
        if (Coords<GEOM>::x2dep && g.ndim >= 2) {
          double dh[3];
          cc.conn2(dh);
          u[2] += rdt * (dh[0] * s0q + dh[1] * s1q + dh[2] * s2q);
        }
      }
#pragma unroll
      for (int m = 0; m < NV; ++m) __stcg(pu[m] + gu, u[m]);
    }
    // advance the column/row bookkeeping of cell g-3 by one chunk
    ic += step_i;
    jc += step_j;
    if (ic >= ni) { ic -= ni; jc += 1; }
  }
}

}  // namespace ab200
