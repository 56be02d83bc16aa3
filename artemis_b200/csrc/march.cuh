// march.cuh -- the x2 / x3 directional passes of the fused stage as MARCHING kernels.
//
// One thread owns one pencil along the pass direction (lanes run along i, so every load and
// store of a warp is a contiguous 256-byte row segment) and walks it cell by cell with the
// reconstruction stencil held in a rotating REGISTER window of four cells:
//
//   step c:  window = q(c-1), q(c), q(c+1), q(c+2)
//            PPM: interface value I(c|c+1) -- computed ONCE per face and shared by the two
//                 cells it bounds (ppm.hpp:39-46: the 4th-order interface value and its first
//                 clamp are symmetric in the two cells, so this is bit-identical to PPM4
//                 evaluated per cell); q(c-1) is dead now, its slot receives q(c+3)
//            monotonise cell c  -> lower edge qr(c), upper edge ql(c+1)      ppm.hpp:48-61
//            Riemann at face c  [ql(c) kept from step c-1]  -> F(c)
//            finish cell c-1 with F(c); start cell c with F(c)
//
// The loop is unrolled by four with the register names rotated statically, so the window
// never moves; every primitive is read from HBM exactly once per pass, every interface value,
// monotonisation and Riemann solve is done exactly once; there is no shared memory and no
// barrier.  It replaces, per direction, K2/K3 + the direction's share of K4/K5 (+K6/K7/
// K12-interior on the last pass): fluid_fluxes.hpp:129-210, artemis_integrator.hpp:95-106,
// fluid_fluxes.hpp:365-392, fill_derived.cpp:55-72,129-164,217-274.
#pragma once
#include "fused.cuh"

namespace ab200 {

#ifndef AB200_MARCH_THREADS
#define AB200_MARCH_THREADS 128
#endif
constexpr int kMarchThreads = AB200_MARCH_THREADS;
#ifndef AB200_MARCH_PREFETCH
#define AB200_MARCH_PREFETCH 6
#endif
constexpr int kMarchPrefetch = AB200_MARCH_PREFETCH;  // steps ahead of the register window
AB_D void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#ifndef AB200_MARCH_MIN_BLOCKS
#define AB200_MARCH_MIN_BLOCKS 2
#endif

// PPM interface value between cells c and c+1 from q(c-1), q(c), q(c+1), q(c+2), with the
// first limiter (clamp to the two adjacent cell values): ppm.hpp:39-46
AB_D double ppm_iface(double qm, double q0, double q1, double q2) {
#ifdef AB200_FAST_MATH
  double v = (7. * (q0 + q1) - (qm + q2)) * (1.0 / 12.0);
#else
  double v = (7. * (q0 + q1) - (qm + q2)) / 12.0;
#endif
  v = dmax(v, dmin(q0, q1));
  v = dmin(v, dmax(q0, q1));
  return v;
}
// PPM monotonisation of cell q_i given its two limited interface values: ppm.hpp:48-61
AB_D void ppm_mono(double qlv, double q_i, double qrv, double &ql_ip1, double &qr_i) {
  const double qc = qrv - q_i;
  const double qd = qlv - q_i;
  if ((qc * qd) >= 0.0) {
    qlv = q_i;
    qrv = q_i;
  } else {
    if (fabs(qc) >= 2.0 * fabs(qd)) qrv = q_i - 2.0 * qd;
    if (fabs(qd) >= 2.0 * fabs(qc)) qlv = q_i - 2.0 * qc;
  }
  ql_ip1 = qrv;
  qr_i = qlv;
}

// AB200_MARCH_MAXNREG (experiments): cap the registers directly instead of through the
// occupancy hint (ptxas maps both (160, 2) and (96, 3) launch bounds to 168 registers)
#ifdef AB200_MARCH_MAXNREG
#define AB200_MARCH_BOUNDS __maxnreg__(AB200_MARCH_MAXNREG)
#else
#define AB200_MARCH_BOUNDS __launch_bounds__(kMarchThreads, AB200_MARCH_MIN_BLOCKS)
#endif
template <int GEOM, int FLUID, int RS, int RC, int DIR, bool LAST>
__global__ void AB200_MARCH_BOUNDS
k_march_pass(GridDev g, FluidDev f, FusedArgs a) {
  static_assert(DIR == 2 || DIR == 3, "marching passes cover x2 and x3");
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr bool CART = (GEOM == AB200_CARTESIAN);
  constexpr bool PPM = (RC == AB200_PPM);
  constexpr int NV = gas ? 6 : 4;

  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int L = DIR == 2 ? njr : nkr;
  const int nL = DIR == 2 ? g.nj : g.nk;
  const int s0 = DIR == 2 ? g.js : g.ks;
  const int ntr = DIR == 2 ? nkr : njr;  // transverse (non-i) extent
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= nir * ntr) return;
  const int b = a.blist ? a.blist[blockIdx.y] : blockIdx.y, n = blockIdx.z;
  const int i = g.is + col % nir;
  const int tr = col / nir;
  const int S = f.S, nvar = f.nvar;
  // element offsets inside one [nk][nj][ni] array fit 32 bits
  const int st = DIR == 2 ? g.ni : g.ni * g.nj;
  const int base = DIR == 2 ? ((g.ks + tr) * g.nj) * g.ni + i : (g.js + tr) * g.ni + i;
  const int jfix = DIR == 2 ? 0 : g.js + tr, kfix = DIR == 2 ? g.ks + tr : 0;

  const double dt = a.dt_dev ? *a.dt_dev : a.dt;
  const double bdt = a.beta * dt;
  const EosConsts eos{f.gm1, f.igm1, f.gamma, f.alpha};

  // primitives in reconstruction order (rho, v_normal, v_t1, v_t2, P, sie), hllc.hpp:66-73;
  // conserved in pack order (rho, m1, m2, m3, E, u)
  const int idx[6] = {n, S + 3 * n + (DIR - 1), S + 3 * n + ((DIR - 1) + 1) % 3,
                      S + 3 * n + ((DIR - 1) + 2) % 3, 4 * S + n, 5 * S + n};
  const int ci[6] = {n, S + 3 * n, S + 3 * n + 1, S + 3 * n + 2, 4 * S + n, 5 * S + n};
  double *pq[NV], *pu[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    pq[v] = f.prim[(size_t)b * nvar + idx[v]] + base;
    pu[v] = f.u0[(size_t)b * nvar + ci[v]] + base;
  }
  // position of pack-order primitive m inside the reconstruction-order pointer list
  // (velocity component comp = m-1 sits at 1 + (comp - (DIR-1) + 3) % 3)
  constexpr int wslot[6] = {0, 1 + (0 - (DIR - 1) + 3) % 3, 1 + (1 - (DIR - 1) + 3) % 3,
                            1 + (2 - (DIR - 1) + 3) % 3, 4, 5};

  // Cartesian: faces areas and the volume factor are separable; in the FAST build the
  // update uses one reciprocal cell width per pencil (differs from the per-cell value only
  // in the last bit of xmin + idx*dx differences)
#if defined(AB200_FAST_MATH)
  constexpr bool HOIST = CART;
#else
  constexpr bool HOIST = false;
#endif
  double rinv = 0.0;
  if (HOIST) {
    const double *xf = DIR == 2 ? g.t.x2f + (size_t)b * (g.nj + 1) : g.t.x3f + (size_t)b * (g.nk + 1);
    rinv = ddiv(bdt, xf[s0 + 1] - xf[s0]);
  }

  auto ldq = [&](int v, int c) -> double {
    int cc = c < 0 ? 0 : (c > nL - 1 ? nL - 1 : c);  // PLM/PCM never use the clamped cells
    return __ldg(pq[v] + cc * st);
  };

  double W0[NV], W1[NV], W2[NV], W3[NV];  // rotating window
  double IA[NV], IB[NV];                  // PPM interface values (lower / upper), alternating
  double QA[NV], QB[NV];                  // upper-edge state of the previous / current cell
  double UA[NV], UB[NV];                  // u0 of the cell being finished / prefetched
  double AA[8], AB[8];                    // lower-face contributions of the open cell, alternating
#pragma unroll
  for (int v = 0; v < NV; ++v) {
    W0[v] = ldq(v, s0 - 3);
    W1[v] = ldq(v, s0 - 2);
    W2[v] = ldq(v, s0 - 1);
    W3[v] = ldq(v, s0);
    IA[v] = IB[v] = QA[v] = QB[v] = UA[v] = UB[v] = 0.0;
  }
#pragma unroll
  for (int m = 0; m < 8; ++m) AA[m] = AB[m] = 0.0;
  const int cend = s0 + L;
  double tmin = 1.79769313486231570815e+308;
  double tden = 0.0;                   // Cartesian fast build: max of sum_d (|v_d| + c_s) / dx_d
  double rdx[3] = {0.0, 0.0, 0.0};     // 1 / dx_d of this pencil's block (0 beyond ndim)
  if (HOIST && LAST) {
    const double *x1f = g.t.x1f + (size_t)b * (g.ni + 1), *x2f = g.t.x2f + (size_t)b * (g.nj + 1);
    const double *x3f = g.t.x3f + (size_t)b * (g.nk + 1);
    rdx[0] = drcp(x1f[g.is + 1] - x1f[g.is]);
    if (g.ndim >= 2) rdx[1] = drcp(x2f[g.js + 1] - x2f[g.js]);
    if (g.ndim >= 3) rdx[2] = drcp(x3f[g.ks + 1] - x3f[g.ks]);
  }

  // one marching step for cell c; Wa..Wd = q(c-1), q(c), q(c+1), q(c+2)
  auto step = [&](const int c, double(&Wa)[NV], double(&Wb)[NV], double(&Wc)[NV],
                  double(&Wd)[NV], double(&Ilo)[NV], double(&Iup)[NV], double(&Qprev)[NV],
                  double(&Qcur)[NV], double(&Ucur)[NV], double(&Unext)[NV],
                  double(&acc)[8], double(&Anext)[8]) {
    if (c > cend) return;
    const int j = DIR == 2 ? c : jfix, k = DIR == 3 ? c : kfix;
    double qr[NV];
    if (PPM) {
#pragma unroll
      for (int v = 0; v < NV; ++v) Iup[v] = ppm_iface(Wa[v], Wb[v], Wc[v], Wd[v]);
      if (c >= s0 - 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) ppm_mono(Ilo[v], Wb[v], Iup[v], Qcur[v], qr[v]);
      }
    } else if (RC == AB200_PLM) {
      double gx[6] = {0, 0, 0, 0, 0, 0};
      if (!CART) plmg_geom<GEOM, DIR>(g, b, k, j, i, gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);
      if (c >= s0 - 1) {
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (CART) plm(Wa[v], Wb[v], Wc[v], Qcur[v], qr[v]);
          else plm_g(Wa[v], Wb[v], Wc[v], Qcur[v], qr[v], gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);
        }
      }
    } else {
#pragma unroll
      for (int v = 0; v < NV; ++v) { Qcur[v] = Wb[v]; qr[v] = Wb[v]; }
    }
    // q(c-1) is dead: its slot receives q(c+3); u0(c) is fetched one step ahead of its use
    if (c < cend) {
      if (c + kMarchPrefetch < nL) {  // pull the lines of a later step into L2 now
#pragma unroll
        for (int v = 0; v < NV; ++v) prefetch_l2(pq[v] + (c + kMarchPrefetch) * st);
        if (c + kMarchPrefetch - 3 < cend) {
#pragma unroll
          for (int v = 0; v < NV; ++v) prefetch_l2(pu[v] + (c + kMarchPrefetch - 3) * st);
        }
      }
#pragma unroll
      for (int v = 0; v < NV; ++v) Wa[v] = ldq(v, c + 3);
      if (c >= s0) {
#pragma unroll
        for (int v = 0; v < NV; ++v) Unext[v] = __ldg(pu[v] + c * st);
      }
    }
    if (c < s0) return;
    // ---- Riemann at the lower face of cell c ------------------------------------------------
    double lo_[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    double(&lo)[8] = HOIST ? Anext : lo_;  // Cartesian fast path: F(c) IS the next step's acc
    Riemann<RS, FLUID>::solve(eos, Qprev, qr, lo);
    if (!CART) {  // ScaleMomentumFlux, fluid_fluxes.hpp:32-70
      Coords<GEOM> cf(g, b, k, j, i);
      double hs[3];
      cf.template face_scale<DIR>(hs);
#pragma unroll
      for (int m = 1; m <= 3; ++m) lo[m] *= hs[(DIR - 1 + (m - 1)) % 3];
    }
    // mass-flux tap (curvilinear meshes only -- RotatingFrameImpl is their only reader -- so the
    // Cartesian kernels carry no trace of it: a run-time test alone cost 5 % per cycle there)
    if (!CART && a.tap) __stcg(f.dflux[DIR - 1][(size_t)b * S + n] + base + c * st, lo[0]);
    // ---- finish cell c-1: ApplyUpdate + FluxSource of direction DIR --------------------------
    // (artemis_integrator.hpp:95-106, fluid_fluxes.hpp:365-392)
    if (c >= s0 + 1) {
      const int off = (c - 1) * st;
      double u[6];
      Coords<GEOM> cc(g, b, DIR == 3 ? c - 1 : kfix, DIR == 2 ? c - 1 : jfix, i);
      if (HOIST) {
        u[0] = Ucur[0] + (acc[0] - lo[0]) * rinv;
#pragma unroll
        for (int m = 1; m <= 3; ++m) {
          const int comp = (DIR - 1 + (m - 1)) % 3;
          u[1 + comp] = Ucur[1 + comp] + (acc[m] - lo[m]) * rinv;
        }
        if (gas) {
          u[4] = Ucur[4] + (acc[4] - lo[4]) * rinv;
          u[5] = Ucur[5] + (acc[5] - lo[5]) * rinv;
          u[1 + (DIR - 1)] += rinv * (acc[6] - lo[6]);
          u[5] -= rinv * 0.5 * (acc[6] + lo[6]) * (lo[7] - acc[7]);
        }
      } else {
        const double a1 = DIR == 2 ? cc.area2(1) : cc.area3();
        const double vol = cc.volume();
#ifdef AB200_FAST_MATH
        const double wv = ddiv(bdt, vol);
#define AB_UPD(x) ((x) * wv)
#else
#define AB_UPD(x) ((x) * bdt / vol)
#endif
        u[0] = Ucur[0] + AB_UPD(acc[0] - a1 * lo[0]);
#pragma unroll
        for (int m = 1; m <= 3; ++m) {
          const int comp = (DIR - 1 + (m - 1)) % 3;
          u[1 + comp] = Ucur[1 + comp] + AB_UPD(acc[m] - a1 * lo[m]);
        }
        if (gas) {
          u[4] = Ucur[4] + AB_UPD(acc[4] - a1 * lo[4]);
          u[5] = Ucur[5] + AB_UPD(acc[5] - a1 * lo[5]);
          const double dxd = DIR == 2 ? cc.x2[1] - cc.x2[0] : cc.x3[1] - cc.x3[0];
          u[1 + (DIR - 1)] += ddiv(bdt, dxd) * (acc[6] - lo[6]);
#ifdef AB200_FAST_MATH
          u[5] -= wv * 0.5 * (acc[6] + lo[6]) * (a1 * lo[7] - acc[7]);
#else
          u[5] -= bdt / vol * 0.5 * (acc[6] + lo[6]) * (a1 * lo[7] - acc[7]);
#endif
        }
#undef AB_UPD
      }
      if (!LAST) {
#pragma unroll
        for (int m = 0; m < NV; ++m) __stcg(pu[m] + off, u[m]);
      } else if (HOIST) {
        // Cartesian fast build: SetAuxillaryFields + ConsToPrim + PrimToCons + the CFL estimate
        // (fill_derived.cpp:55-72, 129-164, 217-274; gas.cpp:411-433) with ONE reciprocal of the
        // floored density and one inverse root (h == 1; the generic branch below spends twelve
        // Newton sequences on the same zone, three of them dividing by the constant 1).  The
        // timestep is reduced as max over zones of sum_d (|v_d| + c_s) / dx_d and inverted once.
        const double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
        const double rwd = drcp(w_d);
        const double v1 = u[1] * rwd, v2 = u[2] * rwd, v3 = u[3] * rwd;
        const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
        __stcg(pq[wslot[0]] + off, w_d);
        __stcg(pq[wslot[1]] + off, v1);
        __stcg(pq[wslot[2]] + off, v2);
        __stcg(pq[wslot[3]] + off, v3);
        __stcg(pu[0] + off, w_d);
        __stcg(pu[1] + off, w_d * v1);
        __stcg(pu[2] + off, w_d * v2);
        __stcg(pu[3] + off, w_d * v3);
        double cs = 0.0;
        if (gas) {
          const double ue = u[4] - ke;
          double w_s = ((ue > f.de_switch * u[4]) ? ue : u[5]) * rwd;
          w_s = dmax(w_s, f.siefloor);
          const double u_u = w_s * w_d;
          __stcg(pq[wslot[5]] + off, w_s);
          __stcg(pq[wslot[4]] + off, dmax(0.0, f.gm1 * w_d * w_s));
          __stcg(pu[5] + off, u_u);
          __stcg(pu[4] + off, u_u + ke);
          if (a.dt_min) cs = dsqrt(dmax(0.0, (f.gm1 + 1) * f.gm1 * w_d * w_s) * rwd);
        }
        if (a.dt_min)
          tden = dmax(tden, (fabs(v1) + cs) * rdx[0] + (fabs(v2) + cs) * rdx[1] +
                                (fabs(v3) + cs) * rdx[2]);
      } else {
        const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
        if (gas)  // SetAuxillaryFields (fill_derived.cpp:55-72)
          u[5] = set_aux_cell(u[0], u[1], u[2], u[3], u[4], u[5], hx, f.dfloor, f.siefloor,
                              f.de_switch);
        // ConsToPrim (fill_derived.cpp:129-164)
        double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
#ifdef AB200_FAST_MATH
        const double rwd = drcp(w_d);
        const double v1 = CART ? u[1] * rwd : ddiv(u[1], w_d * hx[0]);
        const double v2 = CART ? u[2] * rwd : ddiv(u[2], w_d * hx[1]);
        const double v3 = CART ? u[3] * rwd : ddiv(u[3], w_d * hx[2]);
#else
        const double v1 = u[1] / (w_d * hx[0]), v2 = u[2] / (w_d * hx[1]),
                     v3 = u[3] / (w_d * hx[2]);
#endif
        // PrimToCons on the just-computed primitives (fill_derived.cpp:217-274)
        w_d = (w_d > f.dfloor) ? w_d : f.dfloor;
        __stcg(pq[wslot[0]] + off, w_d);
        __stcg(pq[wslot[1]] + off, v1);
        __stcg(pq[wslot[2]] + off, v2);
        __stcg(pq[wslot[3]] + off, v3);
        __stcg(pu[0] + off, w_d);
        __stcg(pu[1] + off, w_d * v1 * hx[0]);
        __stcg(pu[2] + off, w_d * v2 * hx[1]);
        __stcg(pu[3] + off, w_d * v3 * hx[2]);
        if (gas) {
#ifdef AB200_FAST_MATH
          double w_s = u[5] * rwd;
#else
          double w_s = u[5] / ((u[0] > f.dfloor) ? u[0] : f.dfloor);
#endif
          w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
          const double u_u = w_s * w_d;
          __stcg(pq[wslot[5]] + off, w_s);
          __stcg(pq[wslot[4]] + off, dmax(0.0, f.gm1 * w_d * w_s));
          __stcg(pu[5] + off, u_u);
          const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
          __stcg(pu[4] + off, u_u + ke);
          if (a.dt_min) {  // EstimateTimestepMesh folded in (src/gas/gas.cpp:411-433)
            const double vel[3] = {v1, v2, v3};
            tmin = dmin(tmin, cell_dt<GEOM, FLUID>(g, f, cc, w_d, vel, w_s));
          }
        } else if (a.dt_min) {
          const double vel[3] = {v1, v2, v3};
          tmin = dmin(tmin, cell_dt<GEOM, FLUID>(g, f, cc, w_d, vel, 0.0));
        }
      }
    }
    // ---- start cell c: its lower-face contributions --------------------------------------------
    if (!HOIST && c < cend) {
      Coords<GEOM> cs(g, b, k, j, i);
      const double a0 = DIR == 2 ? cs.area2(0) : cs.area3();
#pragma unroll
      for (int m = 0; m < 6; ++m) Anext[m] = a0 * lo[m];
      Anext[6] = lo[6];
      Anext[7] = a0 * lo[7];
    }
  };

  for (int c = s0 - 2; c <= cend; c += 4) {
    step(c + 0, W0, W1, W2, W3, IA, IB, QA, QB, UA, UB, AA, AB);
    step(c + 1, W1, W2, W3, W0, IB, IA, QB, QA, UB, UA, AB, AA);
    step(c + 2, W2, W3, W0, W1, IA, IB, QA, QB, UA, UB, AA, AB);
    step(c + 3, W3, W0, W1, W2, IB, IA, QB, QA, UB, UA, AB, AA);
  }
  if (a.dt_min) {  // warp-shuffle min, one atomic per warp
    if (HOIST && LAST && tden > 0.0) tmin = drcp(tden);
    if (__activemask() == 0xffffffffu) {
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) tmin = dmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
      if ((threadIdx.x & 31) == 0)
        atomicMin(a.dt_min, (unsigned long long)__double_as_longlong(tmin));
    } else {
      atomicMin(a.dt_min, (unsigned long long)__double_as_longlong(tmin));
    }
  }
}

}  // namespace ab200
