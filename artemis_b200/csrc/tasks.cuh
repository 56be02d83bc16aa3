// tasks.cuh -- one kernel per reference task function (the un-fused drop-in path).
// These mirror K1-K7, K12-K15 of SURVEY.md section 2.3 one to one, reading and writing the
// same arrays the reference kernels do, so the Artemis driver can swap any single task.
#pragma once
#include "ab200_ctx.cuh"

namespace ab200 {

constexpr int kThreads = 256;

struct CellIdx {
  int b, k, j, i;
};
// flat -> (b,k,j,i) over a [nb][nkr][njr][nir] box starting at (k0,j0,i0)
AB_D CellIdx decode(long long t, int nir, int njr, int nkr, int i0, int j0, int k0) {
  CellIdx c;
  c.i = (int)(t % nir) + i0; t /= nir;
  c.j = (int)(t % njr) + j0; t /= njr;
  c.k = (int)(t % nkr) + k0; t /= nkr;
  c.b = (int)t;
  return c;
}

// PLM_G inputs for the cell at (k,j,i) along DIR: plm.hpp:92-101, 123-132, 154-163
template <int GEOM, int DIR>
AB_D void plmg_geom(const GridDev &g, int b, int k, int j, int i, double &xm, double &xc,
                    double &xp, double &xf0, double &xf1, double &w) {
  const GeomTab &t = g.t;
  if (DIR == 1) {
    const int o = b * g.ni + i, of = b * (g.ni + 1) + i;
    xm = t.x1v[o - 1]; xc = t.x1v[o]; xp = t.x1v[o + 1];
    xf0 = t.x1f[of]; xf1 = t.x1f[of + 1];
  } else if (DIR == 2) {
    const int o = b * g.nj + j, of = b * (g.nj + 1) + j;
    xm = t.x2v[o - 1]; xc = t.x2v[o]; xp = t.x2v[o + 1];
    xf0 = t.x2f[of]; xf1 = t.x2f[of + 1];
  } else {
    const int o = b * g.nk + k, of = b * (g.nk + 1) + k;
    xm = t.x3v[o - 1]; xc = t.x3v[o]; xp = t.x3v[o + 1];
    xf0 = t.x3f[of]; xf1 = t.x3f[of + 1];
  }
  Coords<GEOM> cc(g, b, k, j, i);
  double ww[3];
  cc.widths(ww);
  w = ww[DIR - 1];
}

// ----------------------------------------------------------------------------------------
// K1-K3: CalculateFluxesImpl (fluid_fluxes.hpp:76-213), one thread per face.
// ----------------------------------------------------------------------------------------
template <int GEOM, int FLUID, int RS, int RC, int DIR>
__global__ void __launch_bounds__(kThreads)
k_calculate_fluxes(GridDev g, FluidDev f) {
  constexpr bool gas = (FLUID == AB200_GAS);
  constexpr bool CART = (GEOM == AB200_CARTESIAN);
  constexpr int NV = gas ? 6 : 4;
  const int nir = g.ie - g.is + 1 + (DIR == 1), njr = g.je - g.js + 1 + (DIR == 2),
            nkr = g.ke - g.ks + 1 + (DIR == 3);
  const long long total = (long long)g.nb * nkr * njr * nir;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  const ptrdiff_t st = DIR == 1 ? 1 : (DIR == 2 ? g.ni : (ptrdiff_t)g.ni * g.nj);
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const int S = f.S;
  double gl[6] = {0, 0, 0, 0, 0, 0}, gr[6] = {0, 0, 0, 0, 0, 0};
  if (!CART && RC == AB200_PLM) {
    const int dk = DIR == 3, dj = DIR == 2, di = DIR == 1;
    plmg_geom<GEOM, DIR>(g, c.b, c.k - dk, c.j - dj, c.i - di, gl[0], gl[1], gl[2], gl[3],
                         gl[4], gl[5]);
    plmg_geom<GEOM, DIR>(g, c.b, c.k, c.j, c.i, gr[0], gr[1], gr[2], gr[3], gr[4], gr[5]);
  }
  double hs[3] = {1.0, 1.0, 1.0};
  if (!CART) {
    Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
    cc.template face_scale<DIR>(hs);
  }
  for (int n = 0; n < S; ++n) {
    // pack indices with the direction permutation of hllc.hpp:66-73
    int idx[6];
    idx[0] = n;
    idx[1] = S + 3 * n + (DIR - 1);
    idx[2] = S + 3 * n + ((DIR - 1) + 1) % 3;
    idx[3] = S + 3 * n + ((DIR - 1) + 2) % 3;
    idx[4] = 4 * S + n;
    idx[5] = 5 * S + n;
    double wl[6], wr[6], out[8], dummy;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const double *q = f.prim[(size_t)c.b * f.nvar + idx[v]] + off;
      // left cell (c - st): its upper-face value is wl; right cell c: lower-face value wr
      recon_cell<RC, CART>(q - st, st, wl[v], dummy, gl[0], gl[1], gl[2], gl[3], gl[4], gl[5]);
      recon_cell<RC, CART>(q, st, dummy, wr[v], gr[0], gr[1], gr[2], gr[3], gr[4], gr[5]);
    }
    Riemann<RS, FLUID>::solve(EosConsts{f.gm1, f.igm1, f.gamma, f.alpha}, wl, wr, out);
    // ScaleMomentumFlux: component IVX*=hx1, IVY*=hx2, IVZ*=hx3 (fluid_fluxes.hpp:64-66)
    if (!CART) {
#pragma unroll
      for (int m = 1; m <= 3; ++m) out[m] *= hs[(DIR - 1 + (m - 1)) % 3];
    }
    double *const *fx = f.flux[DIR - 1] + (size_t)c.b * f.nvar;
    fx[idx[0]][off] = out[0];
    fx[idx[1]][off] = out[1];
    fx[idx[2]][off] = out[2];
    fx[idx[3]][off] = out[3];
    if (gas) {
      fx[idx[4]][off] = out[4];
      fx[idx[5]][off] = out[5];
      f.pflux[DIR - 1][(size_t)c.b * S + n][off] = out[6];
      const size_t foff = ((size_t)c.k * g.fnj + c.j) * g.fni + c.i;
      f.vface[DIR - 1][(size_t)c.b * S + n][foff] = out[7];
    }
  }
}

// ----------------------------------------------------------------------------------------
// K4: ApplyUpdate (artemis_integrator.hpp:56-110)
// ----------------------------------------------------------------------------------------
template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_apply_update(GridDev g, FluidDev f, double gam0, double gam1, double beta_dt) {
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long total = (long long)g.nb * nkr * njr * nir;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  const bool multi_d = g.ndim > 1, three_d = g.ndim > 2;
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double ax1[2] = {cc.area1(cc.x1[0]), cc.area1(cc.x1[1])};
  double ax2[2] = {0.0, 0.0}, ax3[2] = {0.0, 0.0};
  if (multi_d) { ax2[0] = cc.area2(0); ax2[1] = cc.area2(1); }
  if (three_d) { ax3[0] = cc.area3(); ax3[1] = cc.area3(); }
  const double vol = cc.volume();
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const size_t sj = g.ni, sk = (size_t)g.ni * g.nj;
  for (int n = 0; n < f.nvar; ++n) {
    const size_t e = (size_t)c.b * f.nvar + n;
    const double *f1 = f.flux[0][e] + off;
    double divf = (ax1[0] * f1[0] - ax1[1] * f1[1]);
    if (multi_d) {
      const double *f2 = f.flux[1][e] + off;
      divf += (ax2[0] * f2[0] - ax2[1] * f2[sj]);
    }
    if (three_d) {
      const double *f3 = f.flux[2][e] + off;
      divf += (ax3[0] * f3[0] - ax3[1] * f3[sk]);
    }
    double *u0 = f.u0[e] + off;
    const double *u1 = f.u1[e] + off;
    *u0 = gam0 * *u0 + gam1 * *u1 + divf * beta_dt / vol;
  }
}

// ----------------------------------------------------------------------------------------
// K5: FluxSourceImpl (fluid_fluxes.hpp:298-420), interior cells
// ----------------------------------------------------------------------------------------
template <int GEOM, int FLUID>
__global__ void __launch_bounds__(kThreads)
k_flux_source(GridDev g, FluidDev f, double omf, double dt) {
  constexpr bool gas = (FLUID == AB200_GAS);
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long total = (long long)g.nb * nkr * njr * nir;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  const bool multi_d = g.ndim >= 2, three_d = g.ndim == 3;
  const bool x1dep = Coords<GEOM>::x1dep;
  const bool x2dep = Coords<GEOM>::x2dep && multi_d;
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  double dhdx1[3] = {0, 0, 0}, dhdx2[3] = {0, 0, 0};
  if (x1dep) cc.conn1(dhdx1);
  if (x2dep) cc.conn2(dhdx2);
  const double ax1[2] = {cc.area1(cc.x1[0]), cc.area1(cc.x1[1])};
  double ax2[2] = {0.0, 0.0}, ax3[2] = {0.0, 0.0};
  if (multi_d) { ax2[0] = cc.area2(0); ax2[1] = cc.area2(1); }
  if (three_d) { ax3[0] = cc.area3(); ax3[1] = cc.area3(); }
  const double vol = cc.volume();
  const double dx[3] = {cc.x1[1] - cc.x1[0], cc.x2[1] - cc.x2[0], cc.x3[1] - cc.x3[0]};
  double vf[3];
  cc.rotation_velocity(omf, vf);
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const size_t foff = ((size_t)c.k * g.fnj + c.j) * g.fni + c.i;
  const size_t sj = g.ni, sk = (size_t)g.ni * g.nj;
  const size_t fsj = g.fni, fsk = (size_t)g.fni * g.fnj;
  const int S = f.S;
  for (int n = 0; n < S; ++n) {
    const size_t e = (size_t)c.b * f.nvar;
    double *mx = f.u0[e + S + 3 * n + 0] + off;
    double *my = f.u0[e + S + 3 * n + 1] + off;
    double *mz = f.u0[e + S + 3 * n + 2] + off;
    if (gas) {
      double *eg = f.u0[e + 5 * S + n] + off;
      const double *p1 = f.pflux[0][(size_t)c.b * S + n] + off;
      const double *v1 = f.vface[0][(size_t)c.b * S + n] + foff;
      *mx += dt / dx[0] * (p1[0] - p1[1]);
      *eg -= dt / vol * 0.5 * (p1[0] + p1[1]) * (ax1[1] * v1[1] - ax1[0] * v1[0]);
      if (multi_d) {
        const double *p2 = f.pflux[1][(size_t)c.b * S + n] + off;
        const double *v2 = f.vface[1][(size_t)c.b * S + n] + foff;
        *my += dt / dx[1] * (p2[0] - p2[sj]);
        *eg -= dt / vol * 0.5 * (p2[0] + p2[sj]) * (ax2[1] * v2[fsj] - ax2[0] * v2[0]);
      }
      if (three_d) {
        const double *p3 = f.pflux[2][(size_t)c.b * S + n] + off;
        const double *v3 = f.vface[2][(size_t)c.b * S + n] + foff;
        *mz += dt / dx[2] * (p3[0] - p3[sk]);
        *eg -= dt / vol * 0.5 * (p3[0] + p3[sk]) * (ax3[1] * v3[fsk] - ax3[0] * v3[0]);
      }
    }
    if (x1dep || x2dep) {
      const double dens = f.prim[e + n][off];
      const double rdt = dens * dt;
      const double vx = f.prim[e + S + 3 * n + 0][off];
      const double vy = f.prim[e + S + 3 * n + 1][off];
      const double vz = f.prim[e + S + 3 * n + 2][off];
      if (x1dep)
        *mx += rdt * (dhdx1[0] * sqr(vx + vf[0]) + dhdx1[1] * sqr(vy + vf[1]) +
                      dhdx1[2] * sqr(vz + vf[2]));
      if (x2dep)
        *my += rdt * (dhdx2[0] * sqr(vx + vf[0]) + dhdx2[1] * sqr(vy + vf[1]) +
                      dhdx2[2] * sqr(vz + vf[2]));
    }
  }
}

// ----------------------------------------------------------------------------------------
// Per-cell pieces of src/derived/fill_derived.cpp, shared with the fused kernels
// ----------------------------------------------------------------------------------------
// SetAuxillaryFields body (fill_derived.cpp:55-72 + artemis_utils.hpp:50-66): returns new u_u
AB_D double set_aux_cell(double dens, double m1, double m2, double m3, double e_cons,
                         double u_u, const double hx[3], double dflr, double sieflr,
                         double de_switch) {
  double u_d = dens;
  u_d = (u_d > dflr) ? u_d : dflr;
  const double ud2 = dmax(dens, dflr);
  const double rv1 = ddiv(m1, hx[0]), rv2 = ddiv(m2, hx[1]), rv3 = ddiv(m3, hx[2]);
  const double ke = ddiv(0.5 * (sqr(rv1) + sqr(rv2) + sqr(rv3)), ud2);
  const double ue_cons = e_cons - ke;
  double sie = ddiv((ue_cons > de_switch * e_cons) ? ue_cons : u_u, ud2);
  sie = dmax(sie, sieflr);
  double r = sie * u_d;
  const double uflr = sieflr * u_d;
  r = (r > uflr) ? r : uflr;
  return r;
}

template <int GEOM>
__global__ void __launch_bounds__(kThreads) k_set_aux(GridDev g, FluidDev f) {
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long total = (long long)g.nb * nkr * njr * nir;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const int S = f.S;
  const size_t e = (size_t)c.b * f.nvar;
  for (int n = 0; n < S; ++n) {
    double *u_u = f.u0[e + 5 * S + n] + off;
    *u_u = set_aux_cell(f.u0[e + n][off], f.u0[e + S + 3 * n][off], f.u0[e + S + 3 * n + 1][off],
                        f.u0[e + S + 3 * n + 2][off], f.u0[e + 4 * S + n][off], *u_u, hx,
                        f.dfloor, f.siefloor, f.de_switch);
  }
}

// K7: ConsToPrim (fill_derived.cpp:120-166), interior
template <int GEOM, int FLUID>
__global__ void __launch_bounds__(kThreads) k_cons_to_prim(GridDev g, FluidDev f) {
  constexpr bool gas = (FLUID == AB200_GAS);
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long total = (long long)g.nb * nkr * njr * nir;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const int S = f.S;
  const size_t e = (size_t)c.b * f.nvar;
  for (int n = 0; n < S; ++n) {
    const double u_d = f.u0[e + n][off];
    const double w_d = (u_d > f.dfloor) ? u_d : f.dfloor;
    f.prim[e + n][off] = w_d;
#pragma unroll
    for (int d = 0; d < 3; ++d)
      f.prim[e + S + 3 * n + d][off] = f.u0[e + S + 3 * n + d][off] / (w_d * hx[d]);
    if (gas) {
      const double w_s = f.u0[e + 5 * S + n][off] / w_d;
      f.prim[e + 5 * S + n][off] = (w_s > f.siefloor) ? w_s : f.siefloor;
    }
  }
}

// PrimToCons body for one cell (fill_derived.cpp:217-274)
template <int FLUID>
AB_D void prim_to_cons_cell(const FluidDev &f, size_t e, size_t off, int n, const double hx[3]) {
  constexpr bool gas = (FLUID == AB200_GAS);
  const int S = f.S;
  double w_d = f.prim[e + n][off];
  w_d = (w_d > f.dfloor) ? w_d : f.dfloor;
  f.prim[e + n][off] = w_d;
  const double u_d = w_d;
  f.u0[e + n][off] = u_d;
  const double vel1 = f.prim[e + S + 3 * n + 0][off];
  const double vel2 = f.prim[e + S + 3 * n + 1][off];
  const double vel3 = f.prim[e + S + 3 * n + 2][off];
  f.u0[e + S + 3 * n + 0][off] = w_d * vel1 * hx[0];
  f.u0[e + S + 3 * n + 1][off] = w_d * vel2 * hx[1];
  f.u0[e + S + 3 * n + 2][off] = w_d * vel3 * hx[2];
  if (gas) {
    double w_s = f.prim[e + 5 * S + n][off];
    w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
    f.prim[e + 5 * S + n][off] = w_s;
    const double u_u = w_s * u_d;
    f.u0[e + 5 * S + n][off] = u_u;
    // singularity-eos 1.9.1 eos_ideal.hpp:91-95 PressureFromDensityInternalEnergy
    f.prim[e + 4 * S + n][off] = dmax(0.0, f.gm1 * w_d * w_s);
    const double ke = 0.5 * w_d * (sqr(vel1) + sqr(vel2) + sqr(vel3));
    f.u0[e + 4 * S + n][off] = u_u + ke;
  }
}

// K12: PrimToCons over the entire domain, or over ghost zones only (ghosts_only)
template <int GEOM, int FLUID>
__global__ void __launch_bounds__(kThreads)
k_prim_to_cons(GridDev g, FluidDev f, int ghosts_only) {
  const long long total = (long long)g.nb * g.nk * g.nj * g.ni;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= total) return;
  const CellIdx c = decode(t, g.ni, g.nj, g.nk, 0, 0, 0);
  if (ghosts_only && c.i >= g.is && c.i <= g.ie && c.j >= g.js && c.j <= g.je &&
      c.k >= g.ks && c.k <= g.ke)
    return;
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const size_t e = (size_t)c.b * f.nvar;
  for (int n = 0; n < f.S; ++n) prim_to_cons_cell<FLUID>(f, e, off, n, hx);
}

// K15: DeepCopyConservedData (artemis_integrator.hpp:42-49), entire domain
__global__ void __launch_bounds__(kThreads) k_deep_copy(GridDev g, FluidDev f);

// K13/K14: EstimateTimestepMesh (src/gas/gas.cpp:411-433, src/dust/dust.cpp:256-272)
// one zone: 1 / sum_d (|v_d| + c_s) / dl_d   (dust: c_s = 0)
template <int GEOM, int FLUID>
AB_D double cell_dt(const GridDev &g, const FluidDev &f, const Coords<GEOM> &cc, double dens,
                    const double vel[3], double sie) {
  constexpr bool gas = (FLUID == AB200_GAS);
  double dx[3];
  cc.widths(dx);
  double cs = 0.0;
  if (gas) {
    // eos_ideal.hpp:136-140 BulkModulusFromDensityInternalEnergy
    const double bulk = dmax(0.0, (f.gm1 + 1) * f.gm1 * dens * sie);
    cs = dsqrt(ddiv(bulk, dens));
  }
  double denom = 0.0;
#pragma unroll
  for (int d = 0; d < 3; d++) {
    if (d < g.ndim) {
      const double av = fabs(vel[d]);
      denom += gas ? ddiv(av + cs, dx[d]) : ddiv(av, dx[d]);
    }
  }
  return drcp(denom);
}

template <int GEOM, int FLUID>
__global__ void __launch_bounds__(kThreads)
k_estimate_dt(GridDev g, FluidDev f, double *partial) {
  constexpr bool gas = (FLUID == AB200_GAS);
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long total = (long long)g.nb * nkr * njr * nir;
  double ldt = 1.79769313486231570815e+308;  // Big<Real>()
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
    Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
    const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
    const size_t e = (size_t)c.b * f.nvar;
    const int S = f.S;
    for (int n = 0; n < S; ++n) {
      const double dens = f.prim[e + n][off];
      const double sie = gas ? f.prim[e + 5 * S + n][off] : 0.0;
      const double vel[3] = {f.prim[e + S + 3 * n + 0][off], f.prim[e + S + 3 * n + 1][off],
                             f.prim[e + S + 3 * n + 2][off]};
      ldt = dmin(ldt, cell_dt<GEOM, FLUID>(g, f, cc, dens, vel, sie));
    }
  }
  // warp-shuffle min reduction, then one value per CTA
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) ldt = dmin(ldt, __shfl_xor_sync(0xffffffffu, ldt, o));
  __shared__ double sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = ldt;
  __syncthreads();
  if (threadIdx.x < 32) {
    double v = threadIdx.x < kThreads / 32 ? sm[threadIdx.x] : 1.79769313486231570815e+308;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) partial[blockIdx.x] = v;
  }
}

}  // namespace ab200
