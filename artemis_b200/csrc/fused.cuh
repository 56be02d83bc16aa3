// fused.cuh -- the fused fast path: ONE kernel per direction that does
//   reconstruct -> Riemann -> flux difference -> conserved update (+ coordinate sources)
// without ever materialising a flux array, and whose LAST pass also folds in
// SetAuxillaryFields + ConsToPrim + the interior part of PrimToCons.
//
// It replaces, per stage, K1-K7 and the interior of K12, i.e. the
// task chain src/artemis_driver.cpp:184-255 (+ interior of :261).
//
// Work decomposition ("pencil tiles"): a CTA owns NP pencils of one MeshBlock along the pass
// direction.  Every cell of a pencil -- interior cells 0..L-1 plus one halo cell at each end
// -- gets one thread.  Data flow inside the CTA:
//   phase A  each thread reconstructs ITS cell once (all variables); keeps the lower-face
//            state qr in registers, publishes the upper-face state ql to shared memory;
//   phase B  threads of cells 0..L solve the Riemann problem at their LOWER face once
//            (ql from shared memory, qr from registers) and publish the 8 face quantities;
//   phase C  threads of cells 0..L-1 pick up the UPPER face from shared memory and update.
// So every reconstruction and every Riemann solve is done exactly once; the only redundancy
// is 2 idle lanes per pencil in phase C.  Global loads are coalesced along i in all three
// directions (thread->item order is i-fastest); stencil re-reads hit L1.
#pragma once
#include "tasks.cuh"

namespace ab200 {

constexpr int kFusedMaxThreads = 512;

struct FusedArgs {
  double gam0, gam1, beta, dt, omf;
  const double *dt_dev;  // if non-null: dt = *dt_dev
  int first, last, copy_u1;
  int np;       // pencils per CTA
  int npencils; // total pencils
};

template <int GEOM, int FLUID, int RS, int RC, int DIR>
__global__ void __launch_bounds__(kFusedMaxThreads, 1)
k_fused_pass(GridDev g, FluidDev f, FusedArgs a) {
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr bool CART = (GEOM == AB200_CARTESIAN);
  constexpr int NV = gas ? 6 : 4;   // reconstructed variables per species
  constexpr int NF = gas ? 8 : 4;   // face quantities per species
  extern __shared__ double smem[];

  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int L = DIR == 1 ? nir : (DIR == 2 ? njr : nkr);
  const int NP = a.np;
  const int nslots = NP * (L + 1);
  double *s_ql = smem;                 // [NV][nslots]
  double *s_fx = smem + NV * nslots;   // [NF][nslots]

  // ---- thread -> (pencil p, cell c) ---------------------------------------------------
  const int t = threadIdx.x;
  int p, c;
  if (DIR == 1) { p = t / (L + 2); c = t % (L + 2) - 1; }
  else { p = t % NP; c = t / NP - 1; }
  const long long pid = (long long)blockIdx.x * NP + p;
  const bool active = (t < NP * (L + 2)) && (pid < a.npencils);
  // pencil -> block and transverse indices
  int b = 0, k = g.ks, j = g.js, i = g.is;
  if (active) {
    long long r = pid;
    if (DIR == 1) { j = (int)(r % njr) + g.js; r /= njr; k = (int)(r % nkr) + g.ks; r /= nkr; i = g.is + c; }
    if (DIR == 2) { i = (int)(r % nir) + g.is; r /= nir; k = (int)(r % nkr) + g.ks; r /= nkr; j = g.js + c; }
    if (DIR == 3) { i = (int)(r % nir) + g.is; r /= nir; j = (int)(r % njr) + g.js; r /= njr; k = g.ks + c; }
    b = (int)r;
  }
  const ptrdiff_t st = DIR == 1 ? 1 : (DIR == 2 ? g.ni : (ptrdiff_t)g.ni * g.nj);
  const size_t off = ((size_t)k * g.nj + j) * g.ni + i;
  // slot of face `fc` (0..L) of pencil p; consecutive lanes -> consecutive doubles
  auto slot = [&](int fc) { return DIR == 1 ? p * (L + 1) + fc : fc * NP + p; };

  const double dt = a.dt_dev ? *a.dt_dev : a.dt;
  const double bdt = a.beta * dt;
  const int S = f.S;
  const size_t e = (size_t)b * f.nvar;

  // PLM_G geometry of this cell along DIR
  double gx[6] = {0, 0, 0, 0, 0, 0};
  if (!CART && RC == AB200_PLM && active)
    plmg_geom<GEOM, DIR>(g, b, k, j, i, gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);

  for (int n = 0; n < S; ++n) {
    int idx[6];
    idx[0] = n;
    idx[1] = S + 3 * n + (DIR - 1);
    idx[2] = S + 3 * n + ((DIR - 1) + 1) % 3;
    idx[3] = S + 3 * n + ((DIR - 1) + 2) % 3;
    idx[4] = 4 * S + n;
    idx[5] = 5 * S + n;

    // ---- phase A: reconstruct this cell ------------------------------------------------
    double qr[NV], wc[NV];
    if (active) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const double *q = f.prim[e + idx[v]] + off;
        double ql;
        recon_cell<RC, CART>(q, st, ql, qr[v], gx[0], gx[1], gx[2], gx[3], gx[4], gx[5]);
        wc[v] = q[0];
        if (c < L) s_ql[v * nslots + slot(c + 1)] = ql;
      }
    }
    __syncthreads();

    // ---- phase B: Riemann solve at the lower face of cells 0..L --------------------------
    double lo[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    if (active && c >= 0) {
      double wl[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) wl[v] = s_ql[v * nslots + slot(c)];
      Riemann<RS, FLUID>::solve(f.gm1, wl, qr, lo);
      if (!CART) {  // ScaleMomentumFlux, fluid_fluxes.hpp:32-70
        Coords<GEOM> cf(g, b, k, j, i);
        double hs[3];
        cf.template face_scale<DIR>(hs);
#pragma unroll
        for (int m = 1; m <= 3; ++m) lo[m] *= hs[(DIR - 1 + (m - 1)) % 3];
      }
#pragma unroll
      for (int m = 0; m < NF; ++m) s_fx[m * nslots + slot(c)] = (m < 4 || gas) ? lo[m] : 0.0;
    }
    __syncthreads();

    // ---- phase C: flux difference + sources + (last pass) C2P/P2C ------------------------
    if (active && c >= 0 && c < L) {
      double hi[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
      for (int m = 0; m < NF; ++m) hi[m] = s_fx[m * nslots + slot(c + 1)];
      Coords<GEOM> cc(g, b, k, j, i);
      double a0, a1;
      if (DIR == 1) { a0 = cc.area1(cc.x1[0]); a1 = cc.area1(cc.x1[1]); }
      else if (DIR == 2) { a0 = cc.area2(0); a1 = cc.area2(1); }
      else { a0 = cc.area3(); a1 = cc.area3(); }
      const double vol = cc.volume();
      // conserved values in pack order: rho, m1, m2, m3, (E, u)
      const int ci[6] = {n, S + 3 * n, S + 3 * n + 1, S + 3 * n + 2, 4 * S + n, 5 * S + n};
      double u[6];
#pragma unroll
      for (int m = 0; m < NV; ++m) {
        double *p0 = f.u0[e + ci[m]] + off;
        if (a.first) {
          if (a.copy_u1) {  // stage 1 with DeepCopyConservedData folded in: u1 <- u0
            const double v0 = *p0;
            f.u1[e + ci[m]][off] = v0;
            u[m] = v0;
          } else {
            const double v1 = f.u1[e + ci[m]][off];
            u[m] = (a.gam0 == 0.0) ? a.gam1 * v1 : a.gam0 * *p0 + a.gam1 * v1;
          }
        } else {
          u[m] = *p0;
        }
      }
      // ApplyUpdate, direction DIR (artemis_integrator.hpp:95-106)
      u[0] += (a0 * lo[0] - a1 * hi[0]) * bdt / vol;
#pragma unroll
      for (int m = 1; m <= 3; ++m) {
        const int comp = (DIR - 1 + (m - 1)) % 3;
        u[1 + comp] += (a0 * lo[m] - a1 * hi[m]) * bdt / vol;
      }
      if (gas) {
        u[4] += (a0 * lo[4] - a1 * hi[4]) * bdt / vol;
        u[5] += (a0 * lo[5] - a1 * hi[5]) * bdt / vol;
        // FluxSource, direction DIR (fluid_fluxes.hpp:365-392)
        const double dxd = DIR == 1 ? cc.x1[1] - cc.x1[0]
                                    : (DIR == 2 ? cc.x2[1] - cc.x2[0] : cc.x3[1] - cc.x3[0]);
        u[1 + (DIR - 1)] += bdt / dxd * (lo[6] - hi[6]);
        u[5] -= bdt / vol * 0.5 * (lo[6] + hi[6]) * (a1 * hi[7] - a0 * lo[7]);
      }
      // coordinate source terms (fluid_fluxes.hpp:395-415), added once in the first pass
      if (!CART && a.first) {
        // wc[] is in permuted order: velocities back to (1,2,3)
        double vel[3];
#pragma unroll
        for (int m = 0; m < 3; ++m) vel[(DIR - 1 + m) % 3] = wc[1 + m];
        double vf[3];
        cc.rotation_velocity(a.omf, vf);
        const double rdt = wc[0] * bdt;
        const double s0 = sqr(vel[0] + vf[0]), s1 = sqr(vel[1] + vf[1]), s2 = sqr(vel[2] + vf[2]);
        if (Coords<GEOM>::x1dep) {
          double dh[3];
          cc.conn1(dh);
          u[1] += rdt * (dh[0] * s0 + dh[1] * s1 + dh[2] * s2);
        }
        if (Coords<GEOM>::x2dep && g.ndim >= 2) {
          double dh[3];
          cc.conn2(dh);
          u[2] += rdt * (dh[0] * s0 + dh[1] * s1 + dh[2] * s2);
        }
      }
      if (!a.last) {
#pragma unroll
        for (int m = 0; m < NV; ++m) f.u0[e + ci[m]][off] = u[m];
      } else {
        const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
        if (gas)  // SetAuxillaryFields (fill_derived.cpp:55-72)
          u[5] = set_aux_cell(u[0], u[1], u[2], u[3], u[4], u[5], hx, f.dfloor, f.siefloor,
                              f.de_switch);
        // ConsToPrim (fill_derived.cpp:129-164)
        double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
        const double v1 = u[1] / (w_d * hx[0]), v2 = u[2] / (w_d * hx[1]),
                     v3 = u[3] / (w_d * hx[2]);
        // PrimToCons on the just-computed primitives (fill_derived.cpp:217-274)
        w_d = (w_d > f.dfloor) ? w_d : f.dfloor;
        f.prim[e + ci[0]][off] = w_d;
        f.prim[e + ci[1]][off] = v1;
        f.prim[e + ci[2]][off] = v2;
        f.prim[e + ci[3]][off] = v3;
        f.u0[e + ci[0]][off] = w_d;
        f.u0[e + ci[1]][off] = w_d * v1 * hx[0];
        f.u0[e + ci[2]][off] = w_d * v2 * hx[1];
        f.u0[e + ci[3]][off] = w_d * v3 * hx[2];
        if (gas) {
          double w_s = u[5] / ((u[0] > f.dfloor) ? u[0] : f.dfloor);
          w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
          w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
          const double u_u = w_s * w_d;
          f.prim[e + ci[5]][off] = w_s;
          f.prim[e + ci[4]][off] = dmax(0.0, f.gm1 * w_d * w_s);
          f.u0[e + ci[5]][off] = u_u;
          const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
          f.u0[e + ci[4]][off] = u_u + ke;
        }
      }
    }
    if (n + 1 < S) __syncthreads();
  }
}

}  // namespace ab200
