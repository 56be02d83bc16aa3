// diffusion.cu -- viscous stress and heat conduction of the gas.
//
// Reference: src/utils/diffusion/{diffusion_coeff,momentum_diffusion,thermal_diffusion,
// diffusion}.hpp, dispatched by Gas::{ZeroDiffusionFlux,ViscousFlux,ThermalFlux,DiffusionUpdate}
// and Gas::EstimateTimestepMesh (src/gas/gas.cpp:437-467, 524-642).  The reference runs FIVE
// kernels per stage for the fluxes (zero, and per direction a pencil-marching team kernel each
// for viscosity and conduction, each re-evaluating div(u) and the coefficient of every row) and
// one for the update.  Here:
//
//   k_diffusion_coeff   pre-pass: dynamic viscosity, conductivity (the pow() calls) and div(u) of
//                       every zone ONCE (3 doubles per zone and species)
//   k_diffusion_flux    one thread per zone of the interior extended by one face layer: the
//                       fluxes through its (up to three) lower faces, i.e. every face of the
//                       interior is computed exactly once; writes 4 face values per species and
//                       direction (gas.diff.momentum, gas.diff.energy), coalesced along i.
//                       Zeroing is the store itself.
//   k_diffusion_update  Gas::DiffusionUpdate, one thread per interior zone.
//   k_diffusion_dt      the two diffusive timestep limits in one pass over the primitives.
//
// All geometry is position-only and comes from Coords<GEOM> (host-built metric tables: the trig
// of cell centroids that CoordsBase::Distance needs is tabulated with libm), so the strict build
// reproduces the reference bit for bit wherever the coefficient law does not call pow() with a
// non-trivial exponent (CUDA's pow and glibc's differ by <= 2 ulp there).
#include <type_traits>

#include "diffcoef.cuh"

namespace ab200 {

template <typename F>
static int dispatch_geom_d(int geom, F &&fn) {
  switch (geom) {
  case 0: return fn(std::integral_constant<int, 0>{});
  case 1: return fn(std::integral_constant<int, 1>{});
  case 2: return fn(std::integral_constant<int, 2>{});
  case 3: return fn(std::integral_constant<int, 3>{});
  case 4: return fn(std::integral_constant<int, 4>{});
  case 5: return fn(std::integral_constant<int, 5>{});
  }
  set_error("Coordinate type not recognized!");
  return AB200_EINVAL;
}

// Per-thread geometry of the 3 x 3 x 3 neighbourhood of one zone.  Everything the strain tensor
// needs is position-only and separable: centroids x1v(i), x2v(j), x3v(k) (host tables: the same
// expressions evaluated once per index instead of once per use -- each costs a division), the
// centroid trig, hx3v(i, j).  A face kernel then assembles the volume-averaged scale factors and
// the Cartesian image of any neighbour from cached pieces with compile-time offsets (NDIM is a
// template parameter, so inactive directions fold away).  Same operands, same operation order
// as Coords<GEOM>::{x?v, hx?v, to_cart}: the strict build stays bit-identical.
template <int GEOM, int NDIM>
struct GeoCache {
  static constexpr bool multid = NDIM >= 2, threed = NDIM == 3;
  static constexpr bool sph = Coords<GEOM>::sph, sph23 = Coords<GEOM>::sph23;
  double x1[3], x2[3], x3[3];      // centroids at index - 1, index, index + 1
  double s2[3], c2[3], s3[3], c3[3];  // sin / cos of x2v, x3v
  double h3[3][3];                 // hx3v at [dj + 1][di + 1] (spherical 2-D / 3-D)
  double c1_0[3], c2_0[3];         // GetConnX1 / GetConnX2 of the zone itself
  AB_D GeoCache(const GridDev &g, int b, int k, int j, int i) {
    const GeomTab &t = g.t;
#pragma unroll
    for (int q = 0; q < 3; ++q) {
      const int ii = i + q - 1, jj = multid ? j + q - 1 : j, kk = threed ? k + q - 1 : k;
      x1[q] = t.x1v[(size_t)b * g.ni + ii];
      x2[q] = t.x2v[(size_t)b * g.nj + jj];
      x3[q] = t.x3v[(size_t)b * g.nk + kk];
      s2[q] = t.sinv[(size_t)b * g.nj + jj];
      c2[q] = t.cosv[(size_t)b * g.nj + jj];
      s3[q] = t.sin3v[(size_t)b * g.nk + kk];
      c3[q] = t.cos3v[(size_t)b * g.nk + kk];
    }
    if (sph23) {
#pragma unroll
      for (int qj = 0; qj < 3; ++qj)
#pragma unroll
        for (int qi = 0; qi < 3; ++qi) {
          if (!multid && qj != 1) { h3[qj][qi] = 0.0; continue; }
          const Coords<GEOM> c(g, b, k, multid ? j + qj - 1 : j, i + qi - 1);
          h3[qj][qi] = c.hx3v();
        }
    }
    const Coords<GEOM> c(g, b, k, j, i);
    c.conn1(c1_0);
    c.conn2(c2_0);
  }
  // volume-averaged scale factor h_c of the neighbour at offset (oj, oi) (no k dependence)
  AB_D double hx(int c, int oj, int oi) const {
    if (c == 1) return (GEOM == AB200_CYLINDRICAL || sph23) ? x1[oi + 1] : 1.0;
    if (c == 2) {
      if (GEOM == AB200_AXISYMMETRIC) return x1[oi + 1];
      if (sph23) return h3[oj + 1][oi + 1];
    }
    return 1.0;
  }
  // is h_c identically 1 in this geometry?  (x / 1.0 == x: the division is skipped)
  static AB_D constexpr bool unit_h(int c) {
    return c == 0 || (c == 1 && !(GEOM == AB200_CYLINDRICAL || sph23)) ||
           (c == 2 && !(GEOM == AB200_AXISYMMETRIC || sph23));
  }
  // ConvertCoordsToCart of the neighbour's centroid (see Coords<GEOM>::to_cart)
  AB_D void cart(int ok, int oj, int oi, double xo[3]) const {
    const double a = x1[oi + 1];
    if (GEOM == AB200_CYLINDRICAL) {
      xo[0] = a * c2[oj + 1]; xo[1] = a * s2[oj + 1]; xo[2] = x3[ok + 1];
    } else if (GEOM == AB200_AXISYMMETRIC) {
      xo[0] = a * c3[ok + 1]; xo[1] = a * s3[ok + 1]; xo[2] = x2[oj + 1];
    } else if (sph) {
      const double cp = (GEOM == AB200_SPHERICAL3D) ? c3[ok + 1] : 1.0;
      const double sp = (GEOM == AB200_SPHERICAL3D) ? s3[ok + 1] : 0.0;
      const double ct = (GEOM == AB200_SPHERICAL1D) ? 0.0 : c2[oj + 1];
      const double st = (GEOM == AB200_SPHERICAL1D) ? 1.0 : s2[oj + 1];
      xo[0] = a * st * cp; xo[1] = a * st * sp; xo[2] = a * ct;
    } else {
      xo[0] = a; xo[1] = x2[oj + 1]; xo[2] = x3[ok + 1];
    }
  }
  // CoordsBase::Distance between the centroids of two neighbours (geometry.hpp:398-403)
  AB_D double dist(int ak, int aj, int ai, int bk, int bj, int bi) const {
    double p[3], q[3];
    cart(ak, aj, ai, p);
    cart(bk, bj, bi, q);
    return dsqrt(sqr(p[0] - q[0]) + sqr(p[1] - q[1]) + sqr(p[2] - q[2]));
  }
};
// x / h_c with the unit scale factors folded away
template <int GEOM, int NDIM>
AB_D double hdiv(const GeoCache<GEOM, NDIM> &gc, double x, int c, int oj, int oi) {
  return GeoCache<GEOM, NDIM>::unit_h(c) ? x : ddiv(x, gc.hx(c, oj, oi));
}

// FaceAverage selection of StressTensorFaceX* / ThermalFluxImpl: both means are evaluated
// (diffusion_coeff.hpp:148-160, momentum_diffusion.hpp:398-399)
AB_D double face_avg(int avg_type, double m1, double m2) {
  const double avg = (avg_type == AB200_AVG_ARITHMETIC) ? 1.0 : 0.0;
  const double havg = (avg_type == AB200_AVG_HARMONIC) ? 1.0 : 0.0;
  return avg * (0.5 * (m1 + m2)) + havg * ddiv(2.0 * m1 * m2, m1 + m2);
}

#define PRD(v, kk, jj, ii) f.prim[eb + (v)][((size_t)(kk) * g.nj + (jj)) * g.ni + (ii)]

// DiffusionCoeff<viscosity_plaw / viscosity_alpha>::Get from the primitives of one zone
template <int GEOM>
AB_D double visc_mu(const GridDev &g, const FluidDev &f, const DiffDev &dd, int b, int n, int k,
                    int j, int i) {
  const size_t eb = (size_t)b * f.nvar;
  const Coords<GEOM> c(g, b, k, j, i);
  return visc_mu_val<GEOM>(c, dd, f.gm1, PRD(n, k, j, i), PRD(5 * f.S + n, k, j, i));
}
// DiffusionCoeff<conductivity_plaw / thermaldiff_plaw>::Get, diffusion_coeff.hpp:270-384
AB_D double cond_kappa(const GridDev &g, const FluidDev &f, const DiffDev &dd, int b, int n,
                       int k, int j, int i) {
  const size_t eb = (size_t)b * f.nvar;
  const double dens = PRD(n, k, j, i), sie = PRD(5 * f.S + n, k, j, i);
  const double T = dmax(0.0, sie / dd.cv);
  if (dd.cond_type == AB200_COND_CONDUCTIVITY)
    return dd.cond * pow(T / dd.t_ref, dd.temp_exp) * pow(dens / dd.rho_ref, dd.rho_exp);
  return dd.kappa * pow(T / dd.t_ref, dd.temp_exp) * pow(dens / dd.rho_ref, dd.rho_exp) * dens *
         dd.cv;
}
// VelocityDivergence, momentum_diffusion.hpp:553-590
template <int GEOM>
AB_D double vel_div(const GridDev &g, const FluidDev &f, int b, int n, int k, int j, int i) {
  const size_t eb = (size_t)b * f.nvar;
  const int multid = (g.ndim >= 2), threed = (g.ndim == 3);
  const double md = multid, td = threed;
  const Coords<GEOM> c(g, b, k, j, i);
  const double vol = c.volume();
  const double a1[2] = {c.area1(c.x1[0]), c.area1(c.x1[1])};
  const double a2[2] = {multid ? c.area2(0) : 0.0, multid ? c.area2(1) : 0.0};
  const double a3[2] = {threed ? c.area3() : 0.0, threed ? c.area3() : 0.0};
  const int v1 = f.S + 3 * n, v2 = v1 + 1, v3 = v1 + 2;
  const double divv = a1[1] * (PRD(v1, k, j, i) + PRD(v1, k, j, i + 1)) -
                      a1[0] * (PRD(v1, k, j, i) + PRD(v1, k, j, i - 1)) +
                      md * a2[1] * (PRD(v2, k, j, i) + PRD(v2, k, j + multid, i)) -
                      md * a2[0] * (PRD(v2, k, j, i) + PRD(v2, k, j - multid, i)) +
                      td * a3[1] * (PRD(v3, k, j, i) + PRD(v3, k + threed, j, i)) -
                      td * a3[0] * (PRD(v3, k, j, i) + PRD(v3, k - threed, j, i));
  return ddiv(divv, 2.0 * vol);
}
// Viscous flux through the LOWER face of zone (k,j,i) in direction D+1:
// StrainTensorFace<XDIR> + StressTensorFaceX{1,2,3}, momentum_diffusion.hpp:27-551.
// mu0 / div0, mum / divm: coefficient and div(u) of the zone itself and of the zone across the
// face, from the pre-pass (k_diffusion_coeff).
template <int GEOM, int NDIM, int D>
AB_D void visc_face(const GridDev &g, const FluidDev &f, const DiffDev &dd,
                    const GeoCache<GEOM, NDIM> &gc, int b, int n, int k, int j, int i, double mu0,
                    double div0, double mum, double divm, double out[4]) {
  const size_t eb = (size_t)b * f.nvar;
  constexpr int offs[3] = {1, NDIM >= 2, NDIM == 3};
  const int vi[3] = {f.S + 3 * n, f.S + 3 * n + 1, f.S + 3 * n + 2};
  constexpr int dk[3] = {0, 0, 1}, dj[3] = {0, 1, 0}, di[3] = {1, 0, 0};
  constexpr int mk = -dk[D], mj = -dj[D], mi = -di[D];  // the zone across the face
  double v[3], vm[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    v[c] = hdiv(gc, PRD(vi[c], k, j, i), c, 0, 0);
    vm[c] = hdiv(gc, PRD(vi[c], k + mk, j + mj, i + mi), c, mj, mi);
  }
  double hxf[3];
  {
    const Coords<GEOM> cc(g, b, k, j, i);
    cc.template face_scale<D + 1>(hxf);
  }
  const double dxD = gc.dist(0, 0, 0, mk, mj, mi);
  // dh_D/dx_k = {GetConnX1()[D], GetConnX2()[D], 0}; across the face only the factor that
  // depends on the marching index changes (conn1: x1 faces, conn2: theta faces)
  double c1m[3] = {gc.c1_0[0], gc.c1_0[1], gc.c1_0[2]}, c2m[3] = {gc.c2_0[0], gc.c2_0[1], gc.c2_0[2]};
  if (D == 0 || D == 1) {
    const Coords<GEOM> cmc(g, b, k + mk, j + mj, i + mi);
    if (D == 0) cmc.conn1(c1m);
    if (D == 1) cmc.conn2(c2m);
  }
  double flx[3];
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    if (c == D) {
      const double dv = v[D] - vm[D];
      const double src = v[0] * gc.c1_0[D] + v[1] * gc.c2_0[D] + v[2] * 0.0;
      const double srm = vm[0] * c1m[D] + vm[1] * c2m[D] + vm[2] * 0.0;
      flx[c] = ddiv(2 * dv, dxD) + 0.5 * (src + srm);
    } else {
      const int o = offs[c];
      const int ok = o * dk[c], oj = o * dj[c], oi = o * di[c];
      const double dxc = o ? gc.dist(-ok, -oj, -oi, ok, oj, oi) : 1e-99;
      const double dxc_m = o ? gc.dist(mk - ok, mj - oj, mi - oi, mk + ok, mj + oj, mi + oi) : 1e-99;
      const double dvt = hdiv(gc, PRD(vi[D], k + ok, j + oj, i + oi), D, oj, oi) -
                         hdiv(gc, PRD(vi[D], k - ok, j - oj, i - oi), D, -oj, -oi);
      const double dvt_m =
          hdiv(gc, PRD(vi[D], k + mk + ok, j + mj + oj, i + mi + oi), D, mj + oj, mi + oi) -
          hdiv(gc, PRD(vi[D], k + mk - ok, j + mj - oj, i + mi - oi), D, mj - oj, mi - oi);
      const double dv = v[c] - vm[c];
      const double r = ddiv(hxf[c], hxf[D]);
      flx[c] = (double)o * 0.5 * (ddiv(dvt, dxc) + ddiv(dvt_m, dxc_m)) + ddiv((r * r) * dv, dxD);
    }
  }
  const double mus = face_avg(dd.visc_avg, mu0, mum);
  const double divs = (D == 0) ? div0 + divm : divm + div0;
  const double hDf = hxf[D];
  double fc[3];
#pragma unroll
  for (int c = 0; c < 3; ++c)
    fc[c] = (c == D) ? hDf * mus * (flx[c] - 1. / 3 * (1. - dd.eta) * divs) : hDf * mus * flx[c];
  out[0] = fc[0]; out[1] = fc[1]; out[2] = fc[2];
  // (the same quotients as v[] / vm[] above: the reference re-divides, the bits are equal)
  out[3] = 0.5 * (v[0] + vm[0]) * fc[0] + 0.5 * (v[1] + vm[1]) * fc[1] +
           0.5 * (v[2] + vm[2]) * fc[2];
}
// Heat flux through the lower face of zone (k,j,i) in direction D+1, thermal_diffusion.hpp:62-218
template <int GEOM, int NDIM, int D>
AB_D double cond_face(const GridDev &g, const FluidDev &f, const DiffDev &dd,
                      const GeoCache<GEOM, NDIM> &gc, int b, int n, int k, int j, int i,
                      double kap0, double kapm) {
  const size_t eb = (size_t)b * f.nvar;
  constexpr int mk = -(D == 2), mj = -(D == 1), mi = -(D == 0);
  const double dxD = gc.dist(0, 0, 0, mk, mj, mi);
  const double T = dmax(0.0, ddiv(PRD(5 * f.S + n, k, j, i), dd.cv));
  const double Tm = dmax(0.0, ddiv(PRD(5 * f.S + n, k + mk, j + mj, i + mi), dd.cv));
  const double kc = face_avg(dd.cond_avg, kap0, kapm);
  return ddiv(kc * (T - Tm), dxD);
}

// per-zone inputs of the face kernels, computed ONCE per zone by k_diffusion_coeff (the reference
// re-evaluates them for every pencil row of every direction): [3][nb][S][cells] = dynamic
// viscosity, conductivity, div(u)
struct CoefDev {
  double *mu, *kap, *div;
};

template <int GEOM, int NDIM, int D>
AB_D void face_fluxes(const GridDev &g, const FluidDev &f, const DiffDev &dd, const CoefDev &cf,
                      const GeoCache<GEOM, NDIM> &gc, int b, int n, int k, int j, int i) {
  constexpr int dk = (D == 2), dj = (D == 1), di = (D == 0);
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  const size_t co = ((size_t)b * f.S + n) * cells;
  const size_t o0 = co + ((size_t)k * g.nj + j) * g.ni + i;
  const size_t om = co + ((size_t)(k - dk) * g.nj + (j - dj)) * g.ni + (i - di);
  double o[4] = {0.0, 0.0, 0.0, 0.0}, acc[4] = {0.0, 0.0, 0.0, 0.0};
  if (dd.visc_type != AB200_VISC_NONE) {
    visc_face<GEOM, NDIM, D>(g, f, dd, gc, b, n, k, j, i, cf.mu[o0], cf.div[o0], cf.mu[om],
                             cf.div[om], o);
#pragma unroll
    for (int m = 0; m < 4; ++m) acc[m] += o[m];
  }
  if (dd.cond_type != AB200_COND_NONE)
    acc[3] += cond_face<GEOM, NDIM, D>(g, f, dd, gc, b, n, k, j, i, cf.kap[o0], cf.kap[om]);
  const int S = f.S;
  const size_t fcells = (size_t)g.fni * g.fnj * g.fnk;
  const size_t fo = ((size_t)k * g.fnj + j) * g.fni + i;
  double *base = dd.flx[D] + (size_t)b * 4 * S * fcells + fo;
  base[(size_t)(3 * n + 0) * fcells] = acc[0];
  base[(size_t)(3 * n + 1) * fcells] = acc[1];
  base[(size_t)(3 * n + 2) * fcells] = acc[2];
  base[(size_t)(3 * S + n) * fcells] = acc[3];
}

// pre-pass: transport coefficients and div(u) of every zone a face kernel touches = the interior
// extended by one zone on every side of the active directions
template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_diffusion_coeff(GridDev g, FluidDev f, DiffDev dd, CoefDev cf) {
  const int multid = (g.ndim >= 2), threed = (g.ndim == 3);
  const int nir = g.ie - g.is + 3, njr = g.je - g.js + 1 + 2 * multid,
            nkr = g.ke - g.ks + 1 + 2 * threed;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is - 1, g.js - multid, g.ks - threed);
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  for (int n = 0; n < f.S; ++n) {
    const size_t o = ((size_t)c.b * f.S + n) * cells + ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
    if (dd.visc_type != AB200_VISC_NONE) {
      cf.mu[o] = visc_mu<GEOM>(g, f, dd, c.b, n, c.k, c.j, c.i);
      cf.div[o] = vel_div<GEOM>(g, f, c.b, n, c.k, c.j, c.i);
    }
    if (dd.cond_type != AB200_COND_NONE) cf.kap[o] = cond_kappa(g, f, dd, c.b, n, c.k, c.j, c.i);
  }
}

// Gas::ZeroDiffusionFlux + ViscousFlux + ThermalFlux: x1 faces is..ie+1, x2 faces js..je+1,
// x3 faces ks..ke+1 of the interior zones
constexpr int kFluxThreads = 128;  // 120-170 registers: three CTAs of four warps per SM
template <int GEOM, int NDIM>
__global__ void __launch_bounds__(kFluxThreads)
k_diffusion_flux(GridDev g, FluidDev f, DiffDev dd, CoefDev cf) {
  constexpr int multid = (NDIM >= 2), threed = (NDIM == 3);
  const int nir = g.ie - g.is + 2, njr = g.je - g.js + 1 + multid, nkr = g.ke - g.ks + 1 + threed;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  const bool in_i = c.i <= g.ie, in_j = c.j <= g.je, in_k = c.k <= g.ke;
  // a corner / edge zone of the extended box bounds no interior face
  if ((int)!in_i + (int)!in_j + (int)!in_k > 1) return;
  const GeoCache<GEOM, NDIM> gc(g, c.b, c.k, c.j, c.i);
  for (int n = 0; n < f.S; ++n) {
    if (in_j && in_k) face_fluxes<GEOM, NDIM, 0>(g, f, dd, cf, gc, c.b, n, c.k, c.j, c.i);
    if (multid && in_i && in_k) face_fluxes<GEOM, NDIM, 1>(g, f, dd, cf, gc, c.b, n, c.k, c.j, c.i);
    if (threed && in_i && in_j) face_fluxes<GEOM, NDIM, 2>(g, f, dd, cf, gc, c.b, n, c.k, c.j, c.i);
  }
}

// Diffusion::DiffusionUpdateImpl, diffusion.hpp:113-242
template <int GEOM>
__global__ void __launch_bounds__(kThreads, 3)
k_diffusion_update(GridDev g, FluidDev f, DiffDev dd, double dt_host, const double *dt_dev,
                   double beta) {
  const double dt = dt_dev ? beta * *dt_dev : dt_host;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  const int b = c.b, k = c.k, j = c.j, i = c.i;
  const int multi_d = (g.ndim > 1), three_d = (g.ndim > 2);
  const double md = multi_d, td = three_d;
  const bool do_viscosity = dd.visc_type != AB200_VISC_NONE;
  const Coords<GEOM> cc(g, b, k, j, i);
  const double x1dep = Coords<GEOM>::x1dep ? 1.0 : 0.0;
  const double x2dep = (Coords<GEOM>::x2dep && multi_d) ? 1.0 : 0.0;
  const double x3dep = 0.0;
  const double ax1[2] = {cc.area1(cc.x1[0]), cc.area1(cc.x1[1])};
  const double ax2[2] = {multi_d ? cc.area2(0) : 0.0, multi_d ? cc.area2(1) : 0.0};
  const double ax3[2] = {three_d ? cc.area3() : 0.0, three_d ? cc.area3() : 0.0};
  double dhdx1[3] = {0, 0, 0}, dhdx2[3] = {0, 0, 0};
  const double dhdx3[3] = {0, 0, 0};
  if (Coords<GEOM>::x1dep) cc.conn1(dhdx1);
  if (Coords<GEOM>::x2dep && multi_d) cc.conn2(dhdx2);
  const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
  const double vol = cc.volume();
  const int S = f.S;
  const size_t fcells = (size_t)g.fni * g.fnj * g.fnk;
  const size_t fo = ((size_t)k * g.fnj + j) * g.fni + i;
  const size_t fsj = (size_t)g.fni * multi_d, fsk = (size_t)g.fni * g.fnj * three_d;
  const double *F1 = dd.flx[0] + (size_t)b * 4 * S * fcells + fo;
  const double *F2 = (multi_d ? dd.flx[1] : dd.flx[0]) + (size_t)b * 4 * S * fcells + fo;
  const double *F3 = (three_d ? dd.flx[2] : dd.flx[0]) + (size_t)b * 4 * S * fcells + fo;
  const size_t off = ((size_t)k * g.nj + j) * g.ni + i;
  const size_t eb = (size_t)b * f.nvar;
  for (int n = 0; n < S; ++n) {
    const size_t m1 = (size_t)(3 * n) * fcells, m2 = (size_t)(3 * n + 1) * fcells,
                 m3 = (size_t)(3 * n + 2) * fcells, ien = (size_t)(3 * S + n) * fcells;
    double divfxm = 0., divfym = 0., divfzm = 0.;
    if (do_viscosity) {
      const double s1 = 0.5 * (F1[m1] + F1[m1 + 1]);
      const double s2 = 0.5 * (F2[m2] + F2[m2 + fsj]);
      const double s3 = 0.5 * (F3[m3] + F3[m3 + fsk]);
      divfxm = (ax1[0] * F1[m1] - ax1[1] * F1[m1 + 1]) +
               md * (ax2[0] * F2[m1] - ax2[1] * F2[m1 + fsj]) +
               td * (ax3[0] * F3[m1] - ax3[1] * F3[m1 + fsk]);
      divfxm /= vol;
      double src = dhdx1[0] * s1 + md * dhdx1[1] * s2 + td * dhdx1[2] * s3;
      divfxm += x1dep * src;
      divfym = (ax1[0] * F1[m2] - ax1[1] * F1[m2 + 1]) +
               md * (ax2[0] * F2[m2] - ax2[1] * F2[m2 + fsj]) +
               td * (ax3[0] * F3[m2] - ax3[1] * F3[m2 + fsk]);
      divfym /= vol;
      src = dhdx2[0] * s1 + md * dhdx2[1] * s2 + td * dhdx2[2] * s3;
      divfym += x2dep * src;
      divfzm = (ax1[0] * F1[m3] - ax1[1] * F1[m3 + 1]) +
               md * (ax2[0] * F2[m3] - ax2[1] * F2[m3 + fsj]) +
               td * (ax3[0] * F3[m3] - ax3[1] * F3[m3 + fsk]);
      divfzm /= vol;
      src = dhdx3[0] * s1 + md * dhdx3[1] * s2 + td * dhdx3[2] * s3;
      divfzm += x3dep * src;
    }
    double divfe = (ax1[0] * F1[ien] - ax1[1] * F1[ien + 1]) +
                   md * (ax2[0] * F2[ien] - ax2[1] * F2[ien + fsj]) +
                   td * (ax3[0] * F3[ien] - ax3[1] * F3[ien + fsk]);
    divfe /= vol;
    f.u0[eb + S + 3 * n + 0][off] -= dt * divfxm;
    f.u0[eb + S + 3 * n + 1][off] -= dt * divfym;
    f.u0[eb + S + 3 * n + 2][off] -= dt * divfzm;
    f.u0[eb + 4 * S + n][off] -= dt * divfe;
    f.u0[eb + 5 * S + n][off] -=
        dt * divfe - dt * (divfxm * f.prim[eb + S + 3 * n + 0][off] / hx[0] +
                           divfym * f.prim[eb + S + 3 * n + 1][off] / hx[1] +
                           divfzm * f.prim[eb + S + 3 * n + 2][off] / hx[2]);
  }
}

// Diffusion::EstimateTimestep for both operators in one pass (diffusion.hpp:64-111); partial:
// [0..grid) viscous minima, [512..512+grid) conductive minima, BEFORE the 1/(2 ndim) factor
template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_diffusion_dt(GridDev g, FluidDev f, DiffDev dd, double *partial) {
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long total = (long long)g.nb * nkr * njr * nir;
  const double big = 1.79769313486231570815e+308;
  double dtv = big, dtc = big;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
    const Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
    double w[3];
    cc.widths(w);
    double min_dx = big;
    for (int d = 0; d < g.ndim; d++) min_dx = dmin(min_dx, w[d]);
    const size_t eb = (size_t)c.b * f.nvar;
    for (int n = 0; n < f.S; ++n) {
      const double dens = PRD(n, c.k, c.j, c.i);
      if (dd.visc_type != AB200_VISC_NONE) {
        double mu = visc_mu<GEOM>(g, f, dd, c.b, n, c.k, c.j, c.i);
        mu *= (1.0 + ((dd.eta > 1.0) ? 1.0 : 0.0) * (dd.eta - 1.0)) / dens;
        dtv = dmin(dtv, sqr(min_dx) / (mu + 1e-99));
      }
      if (dd.cond_type != AB200_COND_NONE) {
        double mu = cond_kappa(g, f, dd, c.b, n, c.k, c.j, c.i);
        if (dd.cond_type == AB200_COND_CONDUCTIVITY) mu /= (dens * dd.cv);
        dtc = dmin(dtc, sqr(min_dx) / (mu + 1e-99));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    dtv = dmin(dtv, __shfl_xor_sync(0xffffffffu, dtv, o));
    dtc = dmin(dtc, __shfl_xor_sync(0xffffffffu, dtc, o));
  }
  __shared__ double sm[2][kThreads / 32];
  if ((threadIdx.x & 31) == 0) { sm[0][threadIdx.x >> 5] = dtv; sm[1][threadIdx.x >> 5] = dtc; }
  __syncthreads();
  if (threadIdx.x < 32) {
    double a = threadIdx.x < kThreads / 32 ? sm[0][threadIdx.x] : big;
    double c2 = threadIdx.x < kThreads / 32 ? sm[1][threadIdx.x] : big;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a = dmin(a, __shfl_xor_sync(0xffffffffu, a, o));
      c2 = dmin(c2, __shfl_xor_sync(0xffffffffu, c2, o));
    }
    if (threadIdx.x == 0) { partial[blockIdx.x] = a; partial[512 + blockIdx.x] = c2; }
  }
}
#undef PRD

// min over the per-CTA partials; visc_dt = min / (2 ndim), cond_dt likewise, only for the
// configured operators (gas.cpp:437-464); out = [combine: min(out,] cfl * min(visc_dt, cond_dt)
__global__ void k_finish_diffusion_dt(const double *partial, int n, int ndim, int has_v, int has_c,
                                      double cfl, double *out, int combine) {
  const double big = 1.79769313486231570815e+308;
  double a = big, c2 = big;
  for (int i = threadIdx.x; i < n; i += 32) {
    a = dmin(a, partial[i]);
    c2 = dmin(c2, partial[512 + i]);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a = dmin(a, __shfl_xor_sync(0xffffffffu, a, o));
    c2 = dmin(c2, __shfl_xor_sync(0xffffffffu, c2, o));
  }
  if (threadIdx.x == 0) {
    if (has_v) a = a / (2.0 * ndim);
    if (has_c) c2 = c2 / (2.0 * ndim);
    const double v = cfl * dmin(a, c2);
    *out = combine ? dmin(*out, v) : v;
  }
}

static size_t dflx_count(const ab200_ctx *c) {
  const GridDev &g = c->g;
  return (size_t)g.nb * 4 * c->fl[AB200_GAS].d.S * g.fni * g.fnj * g.fnk;
}

static size_t dcoef_count(const ab200_ctx *c) {
  const GridDev &g = c->g;
  return (size_t)g.nb * c->fl[AB200_GAS].d.S * g.ni * g.nj * g.nk;
}

static int ensure_dflx(ab200_ctx *c) {
  const size_t nc = dcoef_count(c);
  if (!c->d_dcoef || c->dcoef_elems != nc) {
    if (c->d_dcoef) cudaFree(c->d_dcoef);
    c->d_dcoef = nullptr;
    AB_CUDA(cudaMalloc((void **)&c->d_dcoef, 3 * nc * sizeof(double)));
    AB_CUDA(cudaMemsetAsync(c->d_dcoef, 0, 3 * nc * sizeof(double), c->stream));
    c->dcoef_elems = nc;
  }
  const size_t n = dflx_count(c);
  for (int d = 0; d < c->g.ndim; ++d) {
    if (c->d_dflx[d] && c->dflx_elems == n) continue;
    if (c->d_dflx[d]) cudaFree(c->d_dflx[d]);
    c->d_dflx[d] = nullptr;
    AB_CUDA(cudaMalloc((void **)&c->d_dflx[d], n * sizeof(double)));
    AB_CUDA(cudaMemsetAsync(c->d_dflx[d], 0, n * sizeof(double), c->stream));
  }
  c->dflx_elems = n;
  return AB200_OK;
}

#define AB_ENTER_D(c)                                                                    \
  AB_REQUIRE((c) != nullptr, AB200_EINVAL, "null context");                              \
  AB_REQUIRE((c)->grid_set, AB200_ESTATE, "no grid bound: call ab200_set_grid");         \
  AB_REQUIRE((c)->has_diffusion, AB200_ESTATE,                                           \
             "diffusion is not configured: call ab200_configure_diffusion");             \
  AB_REQUIRE((c)->fl[AB200_GAS].bound, AB200_ESTATE, "diffusion: no gas bound");         \
  AB_CUDA(cudaSetDevice((c)->device));

int launch_diffusion_flux(ab200_ctx *c) {
  AB_ENTER_D(c)
  AB_REQUIRE(c->g.ng >= 2, AB200_EINVAL, "diffusion needs at least 2 ghost cells");
  AB_TRY(sync_prim_home(c, AB200_GAS, 0));
  AB_TRY(ensure_dflx(c));
  NvtxRange nvtx_("Gas::ZeroDiffusionFlux + ViscousFlux + ThermalFlux [fused]");
  const GridDev &g = c->g;
  const int multid = g.ndim >= 2, threed = g.ndim == 3;
  const long long n = (long long)g.nb * (g.ke - g.ks + 1 + threed) * (g.je - g.js + 1 + multid) *
                      (g.ie - g.is + 2);
  const unsigned grid = (unsigned)((n + kFluxThreads - 1) / kFluxThreads);
  const DiffDev dd = diff_dev(c);
  const CoefDev cf{c->d_dcoef, c->d_dcoef + c->dcoef_elems, c->d_dcoef + 2 * c->dcoef_elems};
  const long long nco = (long long)g.nb * (g.ke - g.ks + 1 + 2 * threed) *
                        (g.je - g.js + 1 + 2 * multid) * (g.ie - g.is + 3);
  const unsigned gridc = (unsigned)((nco + kThreads - 1) / kThreads);
  int rc = dispatch_geom_d(g.geom, [&](auto G) {
    constexpr int GG = decltype(G)::value;
    k_diffusion_coeff<GG><<<gridc, kThreads, 0, c->stream>>>(g, c->fl[AB200_GAS].d, dd, cf);
    const FluidDev &fd = c->fl[AB200_GAS].d;
    // the dimensionalities each coordinate system admits (spherical1D / 2D / 3D fix theirs)
    constexpr bool only1 = GG == AB200_SPHERICAL1D, only2 = GG == AB200_SPHERICAL2D,
                   only3 = GG == AB200_SPHERICAL3D;
    if (g.ndim == 1) {
      if constexpr (!only2 && !only3) k_diffusion_flux<GG, 1><<<grid, kFluxThreads, 0, c->stream>>>(g, fd, dd, cf);
      else { set_error("diffusion: coordinate system and ndim do not match"); return AB200_EINVAL; }
    } else if (g.ndim == 2) {
      if constexpr (!only1 && !only3) k_diffusion_flux<GG, 2><<<grid, kFluxThreads, 0, c->stream>>>(g, fd, dd, cf);
      else { set_error("diffusion: coordinate system and ndim do not match"); return AB200_EINVAL; }
    } else {
      if constexpr (!only1 && !only2) k_diffusion_flux<GG, 3><<<grid, kFluxThreads, 0, c->stream>>>(g, fd, dd, cf);
      else { set_error("diffusion: coordinate system and ndim do not match"); return AB200_EINVAL; }
    }
    return AB200_OK;
  });
  c->launches += 2;
  AB_CUDA(cudaGetLastError());
  return rc;
}

int launch_diffusion_update(ab200_ctx *c, double dt, const double *dt_dev, double beta) {
  AB_ENTER_D(c)
  AB_REQUIRE(c->d_dflx[0] && c->dflx_elems == dflx_count(c), AB200_ESTATE,
             "ab200_diffusion_update: no diffusion fluxes (call ab200_diffusion_flux first)");
  AB_TRY(sync_prim_home(c, AB200_GAS, 0));
  NvtxRange nvtx_("Gas::DiffusionUpdate");
  const GridDev &g = c->g;
  const long long n = (long long)g.nb * (g.ke - g.ks + 1) * (g.je - g.js + 1) * (g.ie - g.is + 1);
  const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
  const DiffDev dd = diff_dev(c);
  int rc = dispatch_geom_d(g.geom, [&](auto G) {
    k_diffusion_update<decltype(G)::value><<<grid, kThreads, 0, c->stream>>>(g, c->fl[AB200_GAS].d, dd, dt, dt_dev, beta);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}

// d_out = [combine: min(d_out,] cfl * min(visc_dt, cond_dt)
int launch_diffusion_dt(ab200_ctx *c, double *d_out, int combine) {
  AB_ENTER_D(c)
  AB_TRY(sync_prim_home(c, AB200_GAS, 0));
  NvtxRange nvtx_("Diffusion::EstimateTimestep");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[AB200_GAS].d;
  const long long total = (long long)g.nb * (g.ke - g.ks + 1) * (g.je - g.js + 1) * (g.ie - g.is + 1);
  int grid = (int)((total + kThreads - 1) / kThreads);
  if (grid > 512) grid = 512;
  double *partial = c->d_red + 1024;  // [1024..2047]: free between the hydro partials and the scalars
  const DiffDev dd = diff_dev(c);
  int rc = dispatch_geom_d(g.geom, [&](auto G) {
    k_diffusion_dt<decltype(G)::value><<<grid, kThreads, 0, c->stream>>>(g, f, dd, partial);
    return AB200_OK;
  });
  k_finish_diffusion_dt<<<1, 32, 0, c->stream>>>(partial, grid, g.ndim, dd.visc_type != AB200_VISC_NONE,
                                                 dd.cond_type != AB200_COND_NONE, f.cfl, d_out, combine);
  c->launches += 2;
  AB_CUDA(cudaGetLastError());
  return rc;
}

}  // namespace ab200

using namespace ab200;

extern "C" {

int ab200_configure_diffusion(ab200_ctx *c, const ab200_diffusion_desc *dd) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  if (!dd || (dd->visc_type == AB200_VISC_NONE && dd->cond_type == AB200_COND_NONE)) {
    c->diffusion = ab200_diffusion_desc{};
    c->has_diffusion = false;
    return AB200_OK;
  }
  AB_REQUIRE(dd->visc_type >= AB200_VISC_NONE && dd->visc_type <= AB200_VISC_ALPHA, AB200_EINVAL,
             "Invalid viscosity type");
  AB_REQUIRE(dd->cond_type >= AB200_COND_NONE && dd->cond_type <= AB200_COND_DIFFUSIVITY,
             AB200_EINVAL, "Invalid conductivity type");
  AB_REQUIRE((dd->visc_avg == AB200_AVG_ARITHMETIC || dd->visc_avg == AB200_AVG_HARMONIC) &&
                 (dd->cond_avg == AB200_AVG_ARITHMETIC || dd->cond_avg == AB200_AVG_HARMONIC),
             AB200_EINVAL, "Invalid diffusion coefficient averaging method");
  AB_REQUIRE(dd->cond_type == AB200_COND_NONE || dd->cv > 0.0, AB200_EINVAL,
             "ab200_configure_diffusion: conduction needs the specific heat cv > 0");
  c->diffusion = *dd;
  c->has_diffusion = true;
  return AB200_OK;
}

int ab200_diffusion_flux(ab200_ctx *c) { return launch_diffusion_flux(c); }

int ab200_diffusion_update(ab200_ctx *c, double dt) {
  return launch_diffusion_update(c, dt, nullptr, 0.0);
}

int ab200_diffusion_timestep(ab200_ctx *c, double *dt_host) {
  AB_REQUIRE(dt_host, AB200_EINVAL, "ab200_diffusion_timestep: null output");
  AB_TRY(launch_diffusion_dt(c, c->d_red + 2050, 0));
  AB_CUDA(cudaMemcpyAsync(c->h_pinned, c->d_red + 2050, sizeof(double), cudaMemcpyDeviceToHost,
                          c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  *dt_host = c->h_pinned[0];
  return AB200_OK;
}

int ab200_diffusion_flux_array(ab200_ctx *c, int dir, double **dev_ptr, size_t *count) {
  AB_REQUIRE(c && dev_ptr && count, AB200_EINVAL, "ab200_diffusion_flux_array: null argument");
  AB_REQUIRE(dir >= 1 && dir <= 3, AB200_EINVAL, "ab200_diffusion_flux_array: dir is 1..3");
  AB_REQUIRE(c->d_dflx[dir - 1], AB200_ESTATE,
             "ab200_diffusion_flux_array: no diffusion fluxes in this direction yet");
  *dev_ptr = c->d_dflx[dir - 1];
  *count = c->dflx_elems;
  return AB200_OK;
}

}  // extern "C"
