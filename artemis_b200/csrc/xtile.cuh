// xtile.cuh -- the x1 directional pass of the fused stage as a THREAD-PER-CELL tile kernel.
//
// xchunk.cuh streams a plane through ONE warp with a three-stage lane pipeline: 160 registers,
// 12 warps per SM, and on B200 the pass is bound by exposed load / shared-memory latency
// (profiles/r01d: 38 % issue slots, 3.5 long-scoreboard stall cycles per issue, 52 % of HBM).
// This kernel spends thread-level parallelism instead: a CTA owns R whole rows of one k-plane of
// one MeshBlock, one thread per cell INCLUDING one halo cell at either end of every row
// (cells -1 .. nx1), short-lived registers only (<= 72 -> 27 warps per SM), three phases:
//
//   A  every thread reconstructs ITS cell (all variables) straight from global memory -- the
//      5-point stencils of neighbouring threads overlap in L1 -- keeps the lower-edge state and
//      publishes the upper-edge state;
//   B  threads of cells 0 .. nx1 solve the Riemann problem at their LOWER face (left state
//      from shared memory) and publish the 8 face quantities;
//   C  threads of cells 0 .. nx1-1 pick up their upper face and update the zone.
//
// Both PPM interface values of a cell are computed by the cell's own thread (the shared value
// is the same expression of the same operands on both sides, ppm.hpp:39-46), so there is no
// interface-value exchange and only two CTA barriers; three CTAs per SM overlap them.  Every
// Riemann solve is done once; the redundancy is the halo cells (2 of nx1 + 2 threads per row).
// As the FIRST pass of a stage it also applies gam0*u0 + gam1*u1 (or the folded u1 <- u0 copy)
// and the curvilinear source terms: fluid_fluxes.hpp:107-126, artemis_integrator.hpp:95-106,
// fluid_fluxes.hpp:365-415.
#pragma once
#include "march.cuh"

namespace ab200 {

constexpr int kXtThreads = 288;   // 4 rows of 64 + 2 cells, rounded up to whole warps

template <int GEOM, int FLUID, int RS, int RC>
__global__ void __launch_bounds__(kXtThreads, 3)
k_xtile_pass(GridDev g, FluidDev f, FusedArgs a) {
  constexpr int DIR = 1;
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr bool CART = (GEOM == AB200_CARTESIAN);
  constexpr int NV = gas ? 6 : 4;
  constexpr int NF = gas ? 8 : 4;
#if defined(AB200_FAST_MATH)
  constexpr bool HOIST = CART;
#else
  constexpr bool HOIST = false;
#endif
  extern __shared__ __align__(16) double xt_smem[];
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int W = nir + 2;        // threads per row: cells -1 .. nir
  const int FS = nir + 1;       // faces per row
  const int R = a.np;           // rows per CTA
  double *sQ = xt_smem;                 // [NV][R][FS] upper-edge states (left states of the faces)
  double *sF = xt_smem + NV * R * FS;   // [NF][R][FS] face quantities
  const int S = f.S, nvar = f.nvar;
  // work item of this CTA: (block, species, plane, row group)
  const int ngrp = (njr + R - 1) / R;
  int w = blockIdx.x;
  const int grp = w % ngrp; w /= ngrp;
  const int k = g.ks + w % nkr; w /= nkr;
  const int n = w % S;
  const int b = a.blist ? a.blist[w / S] : w / S;
  const int t = threadIdx.x;
  const int r = t / W, c = t - r * W - 1;
  const int j = g.js + grp * R + r;
  const bool row_ok = (r < R) && (j <= g.je);
  const int i = g.is + c;

  const double dt = a.dt_dev ? *a.dt_dev : a.dt;
  const double bdt = a.beta * dt;
  const EosConsts eos{f.gm1, f.igm1, f.gamma, f.alpha};
  const int ci[6] = {n, S + 3 * n, S + 3 * n + 1, S + 3 * n + 2, 4 * S + n, 5 * S + n};
  const int off = (k * g.nj + (row_ok ? j : g.js)) * g.ni + i;

  // ---- A: reconstruct the own cell ------------------------------------------------------------
  double qr[NV];
  if (row_ok) {
    double gx[6] = {0, 0, 0, 0, 0, 0};
    if (RC == AB200_PLM && !CART) {
      const int i2 = i < 1 ? 1 : (i > g.ni - 2 ? g.ni - 2 : i);
      plmg_geom<GEOM, DIR>(g, b, k, j, i2, gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);
    }
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double *q = f.prim[(size_t)b * nvar + ci[v]] + off;
      double ql;
      if (RC == AB200_PPM) {
        const double qm2 = __ldg(q - 2), qm1 = __ldg(q - 1), q0 = __ldg(q), qp1 = __ldg(q + 1),
                     qp2 = __ldg(q + 2);
        const double ilo = ppm_iface(qm2, qm1, q0, qp1);
        const double iup = ppm_iface(qm1, q0, qp1, qp2);
        ppm_mono(ilo, q0, iup, ql, qr[v]);
      } else if (RC == AB200_PLM) {
        if (CART) plm(__ldg(q - 1), __ldg(q), __ldg(q + 1), ql, qr[v]);
        else plm_g(__ldg(q - 1), __ldg(q), __ldg(q + 1), ql, qr[v], gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);
      } else {
        ql = __ldg(q);
        qr[v] = ql;
      }
      if (c < nir) sQ[(v * R + r) * FS + c + 1] = ql;
    }
  }
  __syncthreads();
  // ---- B: Riemann at the lower face of cells 0 .. nir -------------------------------------------
  double lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (row_ok && c >= 0) {
    double wl[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) wl[v] = sQ[(v * R + r) * FS + c];
    Riemann<RS, FLUID>::solve(eos, wl, qr, lo);
    if (!CART) {  // ScaleMomentumFlux, fluid_fluxes.hpp:32-70
      Coords<GEOM> cf(g, b, k, j, i);
      double hs[3];
      cf.template face_scale<DIR>(hs);
#pragma unroll
      for (int m = 1; m <= 3; ++m) lo[m] *= hs[m - 1];
    }
#pragma unroll
    for (int m = 0; m < NF; ++m) sF[(m * R + r) * FS + c] = lo[m];
  }
  __syncthreads();
  // ---- C: update cells 0 .. nir-1 (lower face own, upper face from the neighbour) ---------------
  if (row_ok && c >= 0 && c < nir) {
    double hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int m = 0; m < NF; ++m) hi[m] = sF[(m * R + r) * FS + c + 1];
    const double *fl = lo;
    double u[6];
#pragma unroll
    for (int m = 0; m < NV; ++m) {
      double *pu = f.u0[(size_t)b * nvar + ci[m]] + off;
      double *pv = f.u1[(size_t)b * nvar + ci[m]] + off;
      if (a.copy_u1) {  // stage 1 with DeepCopyConservedData folded in: u1 <- u0
        const double v0 = __ldg(pu);
        __stcg(pv, v0);
        u[m] = v0;
      } else {
        u[m] = (a.gam0 == 0.0) ? a.gam1 * __ldg(pv) : a.gam0 * __ldg(pu) + a.gam1 * __ldg(pv);
      }
    }
    Coords<GEOM> cc(g, b, k, j, i);
    if (HOIST) {
      const double rinv = ddiv(bdt, cc.x1[1] - cc.x1[0]);
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
      const double wc0 = __ldg(f.prim[(size_t)b * nvar + ci[0]] + off);
      const double vel[3] = {__ldg(f.prim[(size_t)b * nvar + ci[1]] + off),
                             __ldg(f.prim[(size_t)b * nvar + ci[2]] + off),
                             __ldg(f.prim[(size_t)b * nvar + ci[3]] + off)};
      double vf[3];
      cc.rotation_velocity(a.omf, vf);
      const double rdt = wc0 * bdt;
      const double s0q = sqr(vel[0] + vf[0]), s1q = sqr(vel[1] + vf[1]), s2q = sqr(vel[2] + vf[2]);
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
    for (int m = 0; m < NV; ++m) __stcg(f.u0[(size_t)b * nvar + ci[m]] + off, u[m]);
  }
}

}  // namespace ab200
