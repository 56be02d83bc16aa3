// sources.cu -- pointwise source terms that sit between FluxSource and SetAuxillaryFields in
// the stage task list (src/artemis_driver.cpp:217-248): they read the
// STAGE-START primitives (ConsToPrim has not run yet) and add to the conserved state of the
// interior zones.  One kernel per source over every bound fluid and species, one thread per
// zone, coalesced along i; HBM-bound streaming by construction (read rho, v: 4 doubles per
// species; read-modify-write m, E: 4).  With a source enabled the stage is driven as
//   ab200_fused_stage(.. | AB200_STAGE_DEFER_C2P) -> ab200_<source> ... -> ab200_finish_stage
// because the conserved state must exist in memory between the update and C2P.
//
//   ab200_uniform_gravity  Gravity::UniformGravity<GEOM>     src/gravity/uniform.cpp:28-90
//   ab200_shearing_box     RotatingFrame::ShearingBoxImpl    src/rotating_frame/rotating_frame_impl.hpp:28-94
//   ab200_drag_source      Drag::DragSource<GEOM> in full: simple_dust coupling (implicit gas-
//                          dust drag, constant or Stokes stopping times) or self coupling, damping
//                          zones of gas and dust, damping towards the viscous inflow velocity
//                          src/drag/drag.cpp:88-165, drag.hpp:144-482
//   ab200_drag_simple      the constant-stopping-time, no-damping special case of the above
//   ab200_point_mass_gravity  Gravity::PointMassGravity<GEOM> (softened, off-centre, mass sink)
//                          src/gravity/point_mass.cpp:26-196
//   ab200_rotating_frame   RotatingFrame::RotatingFrameImpl<GEOM> (every curvilinear system)
//                          src/rotating_frame/rotating_frame_impl.hpp:96-199 -- reads the MASS
//                          fluxes of the stage: the fused passes tap them into FluidDev::dflux
//                          (AB200_STAGE_TAP_DFLUX), the task path leaves them in the flux arrays
#include <type_traits>

#include "diffcoef.cuh"

namespace ab200 {

template <typename F>
static int dispatch_geom_s(int geom, F &&fn) {
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

struct TwoFluids {
  FluidDev f[2];
  int on[2];
};

template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_uniform_gravity(GridDev g, TwoFluids tf, double dt_host, const double *dt_dev, double beta,
                  double gx1, double gx2, double gx3) {
  const double dt = dt_dev ? beta * *dt_dev : dt_host;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
  const double gv[3] = {gx1, gx2, gx3};
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
#pragma unroll
  for (int fl = 0; fl < 2; ++fl) {
    if (!tf.on[fl]) continue;
    const FluidDev &f = tf.f[fl];
    const int S = f.S;
    const size_t e = (size_t)c.b * f.nvar;
    for (int n = 0; n < S; ++n) {
      const double rdt = dt * f.prim[e + n][off];
#pragma unroll
      for (int d = 0; d < 3; ++d) f.u0[e + S + 3 * n + d][off] += rdt * hx[d] * gv[d];
      if (fl == AB200_GAS)
        f.u0[e + 4 * S + n][off] += rdt * (f.prim[e + S + 3 * n + 0][off] * gx1 +
                                           f.prim[e + S + 3 * n + 1][off] * gx2 +
                                           f.prim[e + S + 3 * n + 2][off] * gx3);
    }
  }
}

__global__ void __launch_bounds__(kThreads)
k_shearing_box(GridDev g, TwoFluids tf, double dt_host, const double *dt_dev, double beta,
               double om0, double qshear) {
  const double dt = dt_dev ? beta * *dt_dev : dt_host;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  Coords<AB200_CARTESIAN> cc(g, c.b, c.k, c.j, c.i);
  const double three_d = (g.ndim == 3) ? 1.0 : 0.0;
  const double omsq = om0 * om0;
  const double dx = cc.x1[1] - cc.x1[0];
  const double dz = cc.x3[1] - cc.x3[0];
  const double phi_xm1 = -qshear * omsq * cc.x1[0] * cc.x1[0];
  const double phi_xp1 = -qshear * omsq * cc.x1[1] * cc.x1[1];
  const double phi_zm1 = 0.5 * omsq * cc.x3[0] * cc.x3[0];
  const double phi_zp1 = 0.5 * omsq * cc.x3[1] * cc.x3[1];
  const double dpx = (phi_xp1 - phi_xm1) / dx;
  const double dpz = three_d * ((phi_zp1 - phi_zm1) / dz);
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
#pragma unroll
  for (int fl = 0; fl < 2; ++fl) {
    if (!tf.on[fl]) continue;
    const FluidDev &f = tf.f[fl];
    const int S = f.S;
    const size_t e = (size_t)c.b * f.nvar;
    for (int n = 0; n < S; ++n) {
      const double dens = f.prim[e + n][off];
      const double v1 = f.prim[e + S + 3 * n + 0][off], v2 = f.prim[e + S + 3 * n + 1][off],
                   v3 = f.prim[e + S + 3 * n + 2][off];
      const double rdt = dens * dt;
      f.u0[e + S + 3 * n + 0][off] -= rdt * (dpx - 2.0 * om0 * v2);
      f.u0[e + S + 3 * n + 1][off] -= rdt * 2.0 * om0 * v1;
      f.u0[e + S + 3 * n + 2][off] -= rdt * dpz;
      if (fl == AB200_GAS) f.u0[e + 4 * S + n][off] -= rdt * (v1 * dpx + v3 * dpz);
    }
  }
}

struct PointMass {
  double gm, pos[3], rsft2, sink_rate, sink_rad;  // sink_rate: per unit time (x dt in the kernel)
};

// Gravity::PointMassGravity<GEOM>, src/gravity/point_mass.cpp:62-193.  The acceleration is a
// function of position only; the conversions to / from Cartesian use the host-built trig tables
// of the cell centroids, so the strict build reproduces the reference bit for bit.
template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_point_mass(GridDev g, TwoFluids tf, double dt_host, const double *dt_dev, double beta,
             PointMass pm) {
  const double dt = dt_dev ? beta * *dt_dev : dt_host;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
  const double gm = pm.gm;
  double gx1 = 0.0, gx2 = 0.0, gx3 = 0.0, dr;
  if (GEOM == AB200_SPHERICAL1D || GEOM == AB200_SPHERICAL2D) {  // :76-80
    const double rad2 = sqr(cc.x1v()) + pm.rsft2;
    gx1 = -gm / rad2;
    dr = sqrt(rad2);
  } else if (GEOM == AB200_AXISYMMETRIC) {  // :81-88 with axisymmetric.hpp:115-133
    const double x0 = cc.x1v(), x1 = cc.x2v();
    const double rs = sqrt(x0 * x0 + x1 * x1);
    const double ct = x1 / (rs + 1e-99);
    const double st = x0 / (rs + 1e-99);
    dr = rs;
    const double rad2 = sqr(dr) + pm.rsft2;
    const double gg = -gm / rad2;
    gx1 = gg * st;
    gx2 = gg * ct;
  } else {  // :89-114
    double dxc[3], e[3][3];
    cc.to_cart(dxc, e);
#pragma unroll
    for (int n = 0; n < 3; n++) dxc[n] -= pm.pos[n];
    const double R = sqrt(dxc[0] * dxc[0] + dxc[1] * dxc[1]);  // geometry.hpp:262-269
    dr = sqrt(R * R + dxc[2] * dxc[2]);
    const double rad2 = sqr(dr) + pm.rsft2;
    const double idr3 = 1.0 / (sqrt(rad2) * rad2);
    const double multi_d = (g.ndim >= 2) ? 1.0 : 0.0, three_d = (g.ndim == 3) ? 1.0 : 0.0;
    const double gv[3] = {-gm * dxc[0] * idr3, multi_d * (-gm * dxc[1] * idr3),
                          three_d * (-gm * dxc[2] * idr3)};
    gx1 = gv[0] * e[0][0] + gv[1] * e[0][1] + gv[2] * e[0][2];
    gx2 = gv[0] * e[1][0] + gv[1] * e[1][1] + gv[2] * e[1][2];
    gx3 = gv[0] * e[2][0] + gv[1] * e[2][1] + gv[2] * e[2][2];
  }
  // mass accretion :146-148 (quad_ramp(x) = x^2, gravity.hpp:116); sink_rate * dt
  const double srate = dt * pm.sink_rate;
  const double sramp = srate * sqr((dr - pm.sink_rad) / pm.sink_rad);
  double fd = dmin(0.5, sramp / (1.0 + sramp));
  fd *= ((srate > 0.0) && (dr <= pm.sink_rad)) ? 1.0 : 0.0;
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
#pragma unroll
  for (int fl = 0; fl < 2; ++fl) {
    if (!tf.on[fl]) continue;
    const FluidDev &f = tf.f[fl];
    const int S = f.S;
    const size_t e = (size_t)c.b * f.nvar;
    for (int n = 0; n < S; ++n) {
      const double rho = f.prim[e + n][off];
      const double v1 = f.prim[e + S + 3 * n + 0][off], v2 = f.prim[e + S + 3 * n + 1][off],
                   v3 = f.prim[e + S + 3 * n + 2][off];
      double m1 = f.u0[e + S + 3 * n + 0][off], m2 = f.u0[e + S + 3 * n + 1][off],
             m3 = f.u0[e + S + 3 * n + 2][off];
      m1 += dt * rho * hx[0] * gx1;
      m2 += dt * rho * hx[1] * gx2;
      m3 += dt * rho * hx[2] * gx3;
      if (fl == AB200_GAS) {
        const double sie = f.prim[e + 5 * S + n][off];
        const double tote = rho * (sie + 0.5 * (sqr(v1) + sqr(v2) + sqr(v3)));
        double en = f.u0[e + 4 * S + n][off];
        en += dt * rho * (v1 * gx1 + v2 * gx2 + v3 * gx3);
        en -= fd * tote;
        f.u0[e + 4 * S + n][off] = en;
      }
      f.u0[e + n][off] -= fd * rho;
      m1 -= fd * hx[0] * rho * v1;
      m2 -= fd * hx[1] * rho * v2;
      m3 -= fd * hx[2] * rho * v3;
      f.u0[e + S + 3 * n + 0][off] = m1;
      f.u0[e + S + 3 * n + 1][off] = m2;
      f.u0[e + S + 3 * n + 2][off] = m3;
    }
  }
}

// where the mass fluxes of one fluid live: table [nb][stride] per direction, entry n < S
struct MassFlux {
  double *const *tab[3];
  int stride;
};
struct TwoMassFlux {
  MassFlux m[2];
};

// RotatingFrame::RotatingFrameImpl<GEOM>, rotating_frame_impl.hpp:96-199: angular-momentum
// conserving form -- the mass fluxes through the six faces weighted by +-(<R^2>_face - <R^2>).
template <int GEOM>
__global__ void __launch_bounds__(kThreads, 3)
k_rotating_frame(GridDev g, TwoFluids tf, TwoMassFlux mf, double dt_host, const double *dt_dev,
                 double beta, double om0) {
  const double dt = dt_dev ? beta * *dt_dev : dt_host;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double multi_d = (g.ndim >= 2) ? 1.0 : 0.0, three_d = (g.ndim == 3) ? 1.0 : 0.0;
  const double omdt = om0 * dt;
  const double om2dt = omdt * om0;
  double xcyl0, e[3][3], bx[3][2];
  cc.to_cyl(xcyl0, e);
  cc.rf_weights(bx);
  const double ax1[2] = {cc.area1(cc.x1[0]), cc.area1(cc.x1[1])};
  const double ax2[2] = {g.ndim >= 2 ? cc.area2(0) : 0.0, g.ndim >= 2 ? cc.area2(1) : 0.0};
  const double ax3[2] = {g.ndim == 3 ? cc.area3() : 0.0, g.ndim == 3 ? cc.area3() : 0.0};
  const double vol = cc.volume();
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const size_t sj = (size_t)g.ni, sk = (size_t)g.ni * g.nj;
#pragma unroll
  for (int fl = 0; fl < 2; ++fl) {
    if (!tf.on[fl]) continue;
    const FluidDev &f = tf.f[fl];
    const MassFlux &m = mf.m[fl];
    const int S = f.S;
    const size_t eb = (size_t)c.b * f.nvar;
    for (int n = 0; n < S; ++n) {
      const size_t fe = (size_t)c.b * m.stride + n;
      const double *f1 = m.tab[0][fe];
      const double f1m = f1[off], f1p = f1[off + 1];
      double f2m = 0.0, f2p = 0.0, f3m = 0.0, f3p = 0.0;
      if (g.ndim >= 2) { const double *f2 = m.tab[1][fe]; f2m = f2[off]; f2p = f2[off + sj]; }
      if (g.ndim == 3) { const double *f3 = m.tab[2][fe]; f3m = f3[off]; f3p = f3[off + sk]; }
      const double divf = (f1m * ax1[0] * bx[0][0] + f1p * ax1[1] * bx[0][1]) +
                          multi_d * (f2m * ax2[0] * bx[1][0] + f2p * ax2[1] * bx[1][1]) +
                          three_d * (f3m * ax3[0] * bx[2][0] + f3p * ax3[1] * bx[2][1]);
      f.u0[eb + S + 3 * n + 0][off] -= omdt * (divf / vol) * e[0][1];
      f.u0[eb + S + 3 * n + 1][off] -= omdt * (divf / vol) * e[1][1];
      f.u0[eb + S + 3 * n + 2][off] -= omdt * (divf / vol) * e[2][1];
      if (fl == AB200_GAS) {
        const double fx[3] = {0.5 * (f1m + f1p), multi_d * 0.5 * (f2m + f2p),
                              three_d * 0.5 * (f3m + f3p)};
        f.u0[eb + 4 * S + n][off] +=
            om2dt * xcyl0 * (fx[0] * e[0][0] + fx[1] * e[1][0] + fx[2] * e[2][0]);
      }
    }
  }
}

constexpr int kMaxDustSpecies = 16;
struct DragDev {  // ab200_drag_desc as the kernel reads it
  int coupling, model, damp_to_visc;
  double tau[kMaxDustSpecies], scale, grain_density, sizes[kMaxDustSpecies];
  double g_ix[3], g_ox[3], g_irate[3], g_orate[3];
  double d_ix[3], d_ox[3], d_irate[3], d_orate[3];
  double xmin[3], xmax[3];
};

// quadratic damping ramp of one direction (drag.hpp:195-199)
AB_D double damp_ramp(double dt, double x, double ix, double ox, double irate, double orate,
                      double xmin, double xmax) {
  const double ri = (x - ix) / (ix - xmin), ro = (x - ox) / (ox - xmax);
  return dt * (irate * ((x < ix ? 1.0 : 0.0) * (ri * ri)) + orate * ((x > ox ? 1.0 : 0.0) * (ro * ro)));
}

// Drag::SelfDragSourceImpl / SimpleDragSourceImpl (drag.hpp:144-482): the implicit two-pass
// update with the damping ramps bg / bd and the gas target velocity vt; with no damping zone and
// no viscous target they are zeros and the operation order (hence the strict build's bits) is the
// reference's either way.
template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_drag(GridDev g, TwoFluids tf, double dt_host, const double *dt_dev, double beta, DragDev dp,
       DiffDev dd) {
  const double dt = dt_dev ? beta * *dt_dev : dt_host;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  if (t >= (long long)g.nb * nkr * njr * nir) return;
  const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
  Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
  const double xv[3] = {cc.x1v(), cc.x2v(), cc.x3v()};
  const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
  const bool use_visc = dp.damp_to_visc && dd.visc_type != AB200_VISC_NONE;
#ifdef AB200_FAST_MATH
  // default build: the cylindrical basis only feeds the viscous target velocity, and the six
  // damping ramps (two IEEE divisions each) are identically zero without a damping zone --
  // both are uniform over the launch
  const bool need_basis = use_visc;
  bool damp = false;
#pragma unroll
  for (int d = 0; d < 3; ++d)
    damp = damp || dp.g_irate[d] != 0.0 || dp.g_orate[d] != 0.0 || dp.d_irate[d] != 0.0 ||
           dp.d_orate[d] != 0.0;
#else
  const bool need_basis = true, damp = true;
#endif
  double xcyl0 = 1.0, e[3][3] = {{0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}, {0.0, 0.0, 0.0}};
  if (!need_basis) {
  } else if (GEOM == AB200_CARTESIAN) {  // geometry.hpp:284-302
    const double R = sqrt(xv[0] * xv[0] + xv[1] * xv[1]);
    const double cp = xv[0] / (R + 1e-99), sp = xv[1] / (R + 1e-99);
    e[0][0] = cp;  e[0][1] = -sp; e[0][2] = 0.0;
    e[1][0] = sp;  e[1][1] = cp;  e[1][2] = 0.0;
    e[2][0] = 0.0; e[2][1] = 0.0; e[2][2] = 1.0;
    xcyl0 = R;
  } else {
    cc.to_cyl(xcyl0, e);
  }
  const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
  const FluidDev &fg = tf.f[AB200_GAS], &fd_ = tf.f[AB200_DUST];
  const int Sg = tf.on[AB200_GAS] ? fg.S : 0, Sd = tf.on[AB200_DUST] ? fd_.S : 0;
  const size_t eg = (size_t)c.b * fg.nvar, ed = (size_t)c.b * fd_.nvar;
  const double big = 1.79769313486231570815e+308;
  const double dsc[3] = {1.0, g.ndim >= 2 ? 1.0 : 0.0, g.ndim == 3 ? 1.0 : 0.0};
  double bg[3] = {0.0, 0.0, 0.0}, bd[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    if (!damp) break;
    const double rg = damp_ramp(dt, xv[d], dp.g_ix[d], dp.g_ox[d], dp.g_irate[d], dp.g_orate[d],
                                dp.xmin[d], dp.xmax[d]);
    const double rd = damp_ramp(dt, xv[d], dp.d_ix[d], dp.d_ox[d], dp.d_irate[d], dp.d_orate[d],
                                dp.xmin[d], dp.xmax[d]);
    bg[d] = (d == 0) ? rg : dsc[d] * rg;
    bd[d] = (d == 0) ? rd : dsc[d] * rd;
  }
  // ArtemisUtils::GetSpecificInternalEnergy of gas species n (artemis_utils.hpp:42-62)
  auto cons_sie = [&](int n) {
    const double u_d = dmax(fg.u0[eg + n][off], fg.dfloor);
    const double rv1 = fg.u0[eg + Sg + 3 * n + 0][off] / hx[0];
    const double rv2 = fg.u0[eg + Sg + 3 * n + 1][off] / hx[1];
    const double rv3 = fg.u0[eg + Sg + 3 * n + 2][off] / hx[2];
    const double ke = 0.5 * (sqr(rv1) + sqr(rv2) + sqr(rv3)) / u_d;
    const double e_cons = fg.u0[eg + 4 * Sg + n][off];
    const double ue_cons = e_cons - ke;
    const double sie = (ue_cons > fg.de_switch * e_cons) ? ue_cons / u_d
                                                         : fg.u0[eg + 5 * Sg + n][off] / u_d;
    return dmax(sie, fg.siefloor);
  };
  if (dp.coupling == AB200_DRAG_SELF) {  // SelfDragSourceImpl, drag.hpp:150-291
    for (int n = 0; n < Sg; ++n) {
      const double dens = fg.u0[eg + n][off];
      double vg[3];
#pragma unroll
      for (int d = 0; d < 3; ++d) vg[d] = fg.u0[eg + Sg + 3 * n + d][off] / (hx[d] * dens);
      const double sieg = cons_sie(n);
      const double mu = use_visc ? visc_mu_val<GEOM>(cc, dd, fg.gm1, dens, sieg) : 0.0;
      const double vR = -1.5 * mu / (xcyl0 * dens);
      const double vd[3] = {e[0][0] * vR, e[1][0] * vR, e[2][0] * vR};
      const double dm1 = -bg[0] * dens * (vg[0] - vd[0]) / (1.0 + bg[0]);
      const double dm2 = -bg[1] * dens * (vg[1] - vd[1]) / (1.0 + bg[1]);
      const double dm3 = -bg[2] * dens * (vg[2] - vd[2]) / (1.0 + bg[2]);
      fg.u0[eg + Sg + 3 * n + 0][off] += hx[0] * dm1;
      fg.u0[eg + Sg + 3 * n + 1][off] += hx[1] * dm2;
      fg.u0[eg + Sg + 3 * n + 2][off] += hx[2] * dm3;
      fg.u0[eg + 4 * Sg + n][off] += dm1 * (vg[0] + 0.5 * dm1 / dens) +
                                     dm2 * (vg[1] + 0.5 * dm2 / dens) +
                                     dm3 * (vg[2] + 0.5 * dm3 / dens);
    }
    for (int n = 0; n < Sd; ++n) {
#pragma unroll
      for (int d = 0; d < 3; ++d) {
        double *pm = fd_.u0[ed + Sd + 3 * n + d] + off;
        const double mom = *pm;
        *pm = mom - bd[d] * mom / (1.0 + bd[d]);
      }
    }
    return;
  }
#ifdef AB200_FAST_MATH
  // Default build, up to 4 dust species: the same implicit update with every operand loaded
  // ONCE before the first store (a store through double* may alias any later load, so the
  // statement order of the reference serialises one memory round trip per momentum component:
  // 1.0 TB/s measured on the 512^3 gas + 4 dust mesh), one reciprocal per density and per
  // implicit denominator instead of ~25 IEEE divisions per species, and the dust state kept in
  // registers between the two passes (4 slots, statically unrolled).  Agrees with the
  // reference's statement order (the strict build below) to a few ulp; the parity tests hold it
  // to 1e-12 per zone.
  auto drag_fast = [&](auto NSc) {
    constexpr int NS = decltype(NSc)::value;
    const bool cart = (GEOM == AB200_CARTESIAN);
    double rho[NS], mom[NS][3];
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      rho[n] = 1.0;
      mom[n][0] = mom[n][1] = mom[n][2] = 0.0;
      if (n < Sd) {
        rho[n] = fd_.u0[ed + n][off];
#pragma unroll
        for (int d = 0; d < 3; ++d) mom[n][d] = fd_.u0[ed + Sd + 3 * n + d][off];
      }
    }
    const double dg = fg.u0[eg][off];
    double mg[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) mg[d] = fg.u0[eg + Sg + d][off];
    const double eg0 = fg.u0[eg + 4 * Sg][off];
    const bool need_sie = use_visc || dp.model == AB200_DRAG_STOKES;
    const double sieg = need_sie ? cons_sie(0) : 0.0;
    double ihx[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) ihx[d] = cart ? 1.0 : drcp(hx[d]);
    const double rdg = drcp(dg);
    double vg[3];
#pragma unroll
    for (int d = 0; d < 3; ++d) vg[d] = mg[d] * ihx[d] * rdg;
    const double mu = use_visc ? visc_mu_val<GEOM>(cc, dd, fg.gm1, dg, sieg) : 0.0;
    const double vR = use_visc ? -1.5 * mu * drcp(xcyl0 * dg) : 0.0;
    const double vt[3] = {e[0][0] * vR, e[1][0] * vR, e[2][0] * vR};
    const double rvth = (dp.model == AB200_DRAG_STOKES)
                            ? drsqrt(8.0 / 3.14159265358979323846 * fg.gm1 * sieg) : 0.0;
    const bool bd_uniform = (bd[0] == bd[1]) && (bd[1] == bd[2]);
    double al[NS], rrho[NS], rden0[NS];
    double fd[3] = {0.0, 0.0, 0.0}, fvd[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      al[n] = rrho[n] = rden0[n] = 0.0;
      if (n < Sd) {
        // alpha = dt / t_stop (drag.hpp:356-366); t_stop <= 0 couples instantly
        const double tc = (dp.model == AB200_DRAG_STOKES)
                              ? dp.scale * dp.grain_density * rdg * dp.sizes[n] * rvth
                              : dp.scale * dp.tau[n];
        al[n] = dt * ((tc <= 0.0) ? big : drcp(tc));
        rrho[n] = drcp(rho[n]);
        rden0[n] = drcp(1.0 + al[n] + bd[0]);
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double rden = (d == 0 || bd_uniform) ? rden0[n] : drcp(1.0 + al[n] + bd[d]);
          const double vd = mom[n][d] * ihx[d] * rrho[n];
          const double rhop = rho[n] * al[n] * rden;
          fd[d] += rhop * (1.0 + bd[d]);
          fvd[d] += rhop * vd;
        }
      }
    }
    double vgp[3], delta_g[3] = {0.0, 0.0, 0.0};
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      vgp[d] = (dg * (vg[d] + bg[d] * vt[d]) + fvd[d]) * drcp(dg * (1.0 + bg[d]) + fd[d]);
      fvd[d] = 0.0;
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      if (n < Sd) {
#pragma unroll
        for (int d = 0; d < 3; ++d) {
          const double rden = (d == 0 || bd_uniform) ? rden0[n] : drcp(1.0 + al[n] + bd[d]);
          const double vd = mom[n][d] * ihx[d] * rrho[n];
          const double rhop = rho[n] * al[n] * rden;
          const double delta = rhop * ((vgp[d] - vd) + bd[d] * vgp[d]);
          delta_g[d] -= delta;
          const double delta_d = delta - bd[d] * rho[n] * rden * (vd + al[n] * vgp[d]);
          fvd[d] += rhop * ((vd - vt[d]) - bd[d] * vt[d]);
          mom[n][d] += hx[d] * delta_d;
        }
      }
    }
    double eadd = 0.0;
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double prefac = dg * bg[d] * drcp(1.0 + bg[d] + fd[d]);
      delta_g[d] -= prefac * (dg * (vg[d] - vt[d]) + fvd[d]);
      mg[d] += hx[d] * delta_g[d];
      eadd += 0.5 * (vg[d] + vgp[d]) * delta_g[d];
    }
#pragma unroll
    for (int n = 0; n < NS; ++n) {
      if (n < Sd) {
#pragma unroll
        for (int d = 0; d < 3; ++d) fd_.u0[ed + Sd + 3 * n + d][off] = mom[n][d];
      }
    }
#pragma unroll
    for (int d = 0; d < 3; ++d) fg.u0[eg + Sg + d][off] = mg[d];
    fg.u0[eg + 4 * Sg][off] = eg0 + eadd;
  };
  if (Sd <= 4) {   // more species take the generic statement order below
    drag_fast(std::integral_constant<int, 4>{});
    return;
  }
#endif
  // SimpleDragSourceImpl, drag.hpp:296-482 (gas species 0 couples to every dust species)
  const double dg = fg.u0[eg][off];
  double vg[3], fd[3] = {0.0, 0.0, 0.0}, fvd[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < 3; ++d) vg[d] = fg.u0[eg + Sg + d][off] / (hx[d] * dg);
  const double sieg = cons_sie(0);
  const double mu = use_visc ? visc_mu_val<GEOM>(cc, dd, fg.gm1, dg, sieg) : 0.0;
  const double vR = -1.5 * mu / (xcyl0 * dg);
  const double vt[3] = {e[0][0] * vR, e[1][0] * vR, e[2][0] * vR};
  const double vdt[3] = {0.0, 0.0, 0.0};
  double vth = 0.0;
  if (dp.model == AB200_DRAG_STOKES) vth = sqrt(8.0 / 3.14159265358979323846 * fg.gm1 * sieg);
  auto stopping_time = [&](int n) {
    if (dp.model == AB200_DRAG_STOKES) return dp.scale * dp.grain_density / dg * dp.sizes[n] / vth;
    return dp.scale * dp.tau[n];
  };
  for (int n = 0; n < Sd; ++n) {
    const double dens = fd_.u0[ed + n][off];
    const double tc = stopping_time(n);
    const double alpha = dt * ((tc <= 0.0) ? big : 1.0 / tc);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      const double vd = fd_.u0[ed + Sd + 3 * n + d][off] / (hx[d] * dens);
      const double rhop = dens * alpha / (1.0 + alpha + bd[d]);
      fd[d] += rhop * (1.0 + bd[d]);
      fvd[d] += rhop * (vd + bd[d] * vdt[d]);
    }
  }
  double vgp[3], delta_g[3] = {0.0, 0.0, 0.0};
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    vgp[d] = (dg * (vg[d] + bg[d] * vt[d]) + fvd[d]) / (dg * (1.0 + bg[d]) + fd[d]);
    fvd[d] = 0.;
  }
  for (int n = 0; n < Sd; ++n) {
    const double dens = fd_.u0[ed + n][off];
    const double tc = stopping_time(n);
    const double alpha = dt * ((tc <= 0.0) ? big : 1.0 / tc);
#pragma unroll
    for (int d = 0; d < 3; ++d) {
      double *pm = fd_.u0[ed + Sd + 3 * n + d] + off;
      const double vd = *pm / (hx[d] * dens);
      double delta_d = 0.;
      const double rhop = dens * alpha / (1.0 + alpha + bd[d]);
      const double delta = rhop * ((vgp[d] - vd + bd[d] * (vgp[d] - vdt[d])));
      delta_d += delta;
      delta_g[d] -= delta;
      delta_d -= bd[d] * dens / (1. + alpha + bd[d]) * (vd - vdt[d] + alpha * (vgp[d] - vdt[d]));
      fvd[d] += rhop * (vd - vt[d] + bd[d] * (vdt[d] - vt[d]));
      *pm += hx[d] * delta_d;
    }
  }
#pragma unroll
  for (int d = 0; d < 3; ++d) {
    const double prefac = dg * bg[d] / (1.0 + bg[d] + fd[d]);
    delta_g[d] -= prefac * (dg * (vg[d] - vt[d]) + fvd[d]);
    fg.u0[eg + Sg + d][off] += hx[d] * delta_g[d];
    fg.u0[eg + 4 * Sg][off] += 0.5 * (vg[d] + vgp[d]) * delta_g[d];
  }
}

// SetAuxillaryFields + ConsToPrim + interior PrimToCons (+ the CFL estimate) of ONE fluid in one
// pointwise kernel: the second half of a split stage (fill_derived.cpp:55-72, 129-164, 217-274;
// gas.cpp:411-433 / dust.cpp:256-272).  Same device functions, same operation order as the LAST
// marching pass (march.cuh), so the strict build stays bit-identical to the task kernels.
template <int GEOM, int FLUID>
__global__ void __launch_bounds__(kThreads)
k_finish_stage(GridDev g, FluidDev f, unsigned long long *dt_min) {
  constexpr bool gas = (FLUID == AB200_GAS);
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x;
  double tmin = 1.79769313486231570815e+308;
  if (t < (long long)g.nb * nkr * njr * nir) {
    const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
    Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
    const double hx[3] = {cc.hx1v(), cc.hx2v(), cc.hx3v()};
    const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
    const int S = f.S;
    const size_t e = (size_t)c.b * f.nvar;
    for (int n = 0; n < S; ++n) {
      double u[6];
      u[0] = f.u0[e + n][off];
#pragma unroll
      for (int d = 0; d < 3; ++d) u[1 + d] = f.u0[e + S + 3 * n + d][off];
      if (gas) {
        u[4] = f.u0[e + 4 * S + n][off];
        u[5] = f.u0[e + 5 * S + n][off];
        u[5] = set_aux_cell(u[0], u[1], u[2], u[3], u[4], u[5], hx, f.dfloor, f.siefloor, f.de_switch);
      }
      double w_d = (u[0] > f.dfloor) ? u[0] : f.dfloor;
      const double v1 = ddiv(u[1], w_d * hx[0]), v2 = ddiv(u[2], w_d * hx[1]),
                   v3 = ddiv(u[3], w_d * hx[2]);
      w_d = (w_d > f.dfloor) ? w_d : f.dfloor;
      f.prim[e + n][off] = w_d;
      f.prim[e + S + 3 * n + 0][off] = v1;
      f.prim[e + S + 3 * n + 1][off] = v2;
      f.prim[e + S + 3 * n + 2][off] = v3;
      f.u0[e + n][off] = w_d;
      f.u0[e + S + 3 * n + 0][off] = w_d * v1 * hx[0];
      f.u0[e + S + 3 * n + 1][off] = w_d * v2 * hx[1];
      f.u0[e + S + 3 * n + 2][off] = w_d * v3 * hx[2];
      const double vel[3] = {v1, v2, v3};
      if (gas) {
        double w_s = ddiv(u[5], (u[0] > f.dfloor) ? u[0] : f.dfloor);
        w_s = (w_s > f.siefloor) ? w_s : f.siefloor;
        const double u_u = w_s * w_d;
        f.prim[e + 5 * S + n][off] = w_s;
        f.prim[e + 4 * S + n][off] = dmax(0.0, f.gm1 * w_d * w_s);
        f.u0[e + 5 * S + n][off] = u_u;
        const double ke = 0.5 * w_d * (sqr(v1) + sqr(v2) + sqr(v3));
        f.u0[e + 4 * S + n][off] = u_u + ke;
        if (dt_min) tmin = dmin(tmin, cell_dt<GEOM, FLUID>(g, f, cc, w_d, vel, w_s));
      } else if (dt_min) {
        tmin = dmin(tmin, cell_dt<GEOM, FLUID>(g, f, cc, w_d, vel, 0.0));
      }
    }
  }
  if (dt_min) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) tmin = dmin(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
    if ((threadIdx.x & 31) == 0) atomicMin(dt_min, (unsigned long long)__double_as_longlong(tmin));
  }
}

int launch_finish_stage(ab200_ctx *c, int fluid, unsigned long long *dt_min) {
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fluid].d;
  const long long n = (long long)g.nb * (g.ke - g.ks + 1) * (g.je - g.js + 1) * (g.ie - g.is + 1);
  const unsigned grid = (unsigned)((n + kThreads - 1) / kThreads);
  NvtxRange nvtx_("SetAuxillaryFields + ConsToPrim + PrimToCons [fused]");
  int rc = dispatch_geom_s(g.geom, [&](auto G) {
    constexpr int GG = decltype(G)::value;
    if (fluid == AB200_GAS) k_finish_stage<GG, AB200_GAS><<<grid, kThreads, 0, c->stream>>>(g, f, dt_min);
    else k_finish_stage<GG, AB200_DUST><<<grid, kThreads, 0, c->stream>>>(g, f, dt_min);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}

static unsigned grid_s(const GridDev &g) {
  const long long n = (long long)g.nb * (g.ke - g.ks + 1) * (g.je - g.js + 1) * (g.ie - g.is + 1);
  return (unsigned)((n + kThreads - 1) / kThreads);
}

static int two_fluids(ab200_ctx *c, TwoFluids &tf) {
  int any = 0;
  for (int f = 0; f < 2; ++f) {
    tf.on[f] = c->fl[f].bound ? 1 : 0;
    if (tf.on[f]) {
      AB_TRY(sync_prim_home(c, f, 0));  // the sources read the caller-visible primitives
      tf.f[f] = c->fl[f].d;
      any = 1;
    }
  }
  AB_REQUIRE(any, AB200_ESTATE, "source term: no fluid bound");
  return AB200_OK;
}

}  // namespace ab200

using namespace ab200;

#define AB_ENTER_S(c)                                                                    \
  AB_REQUIRE((c) != nullptr, AB200_EINVAL, "null context");                              \
  AB_REQUIRE((c)->grid_set, AB200_ESTATE, "no grid bound: call ab200_set_grid");         \
  AB_CUDA(cudaSetDevice((c)->device));

extern "C" {

static int gravity_impl(ab200_ctx *c, double dt, const double *dt_dev, double beta, double gx1,
                        double gx2, double gx3) {
  AB_ENTER_S(c)
  TwoFluids tf{};
  AB_TRY(two_fluids(c, tf));
  const GridDev &g = c->g;
  int rc = dispatch_geom_s(g.geom, [&](auto G) {
    k_uniform_gravity<decltype(G)::value><<<grid_s(g), kThreads, 0, c->stream>>>(g, tf, dt, dt_dev, beta, gx1, gx2, gx3);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}
int ab200_uniform_gravity(ab200_ctx *c, double dt, double gx1, double gx2, double gx3) {
  return gravity_impl(c, dt, nullptr, 0.0, gx1, gx2, gx3);
}

static int shearing_impl(ab200_ctx *c, double dt, const double *dt_dev, double beta, double omega,
                         double qshear) {
  AB_ENTER_S(c)
  // src/rotating_frame/rotating_frame.cpp:31-37
  AB_REQUIRE(omega != 0.0, AB200_EINVAL, "rotating_frame/omega cannot be zero!");
  AB_REQUIRE(c->g.geom == AB200_CARTESIAN, AB200_EINVAL,
             "ab200_shearing_box: the shearing box is Cartesian (curvilinear frames use "
             "RotatingFrameImpl, which stays on the reference path)");
  TwoFluids tf{};
  AB_TRY(two_fluids(c, tf));
  k_shearing_box<<<grid_s(c->g), kThreads, 0, c->stream>>>(c->g, tf, dt, dt_dev, beta, omega, qshear);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}
int ab200_shearing_box(ab200_ctx *c, double dt, double omega, double qshear) {
  return shearing_impl(c, dt, nullptr, 0.0, omega, qshear);
}

static int drag_impl(ab200_ctx *c, double dt, const double *dt_dev, double beta,
                     const ab200_drag_desc *d) {
  AB_ENTER_S(c)
  AB_REQUIRE(d != nullptr, AB200_EINVAL, "ab200_drag_source: null descriptor");
  AB_REQUIRE(d->coupling == AB200_DRAG_SIMPLE_DUST || d->coupling == AB200_DRAG_SELF, AB200_EINVAL,
             "Invalid drag model!");
  AB_REQUIRE(d->model == AB200_DRAG_CONSTANT || d->model == AB200_DRAG_STOKES, AB200_EINVAL,
             "bad type for stopping time model");
  const bool gas = c->fl[0].bound, dust = c->fl[1].bound;
  if (d->coupling == AB200_DRAG_SIMPLE_DUST) {
    // src/drag/drag.cpp:76-80, drag.hpp:387
    AB_REQUIRE(gas && dust, AB200_ESTATE, "drag type simple_dust requires do_gas = do_dust = true");
    AB_REQUIRE(c->fl[0].d.S == 1, AB200_EINVAL,
               "ab200_drag_source: simple_dust couples gas species 0 only (drag.hpp:387)");
    AB_REQUIRE(c->fl[1].d.S <= kMaxDustSpecies, AB200_EINVAL, "ab200_drag_source: too many dust species");
  }
  AB_REQUIRE(gas || dust, AB200_ESTATE, "source term: no fluid bound");
  for (int k = 0; k < 3; ++k) {  // drag.hpp:101-106
    AB_REQUIRE(d->g_irate[k] >= 0.0 && d->d_irate[k] >= 0.0, AB200_EINVAL,
               "The damping rate in the x1 direction must be >= 0");
    AB_REQUIRE(d->g_ix[k] <= d->g_ox[k] && d->d_ix[k] <= d->d_ox[k], AB200_EINVAL,
               "The damping bounds must have inner_x1 <= outer_x1");
  }
  AB_REQUIRE(!d->g_damp_to_visc || (c->has_diffusion && c->diffusion.visc_type != AB200_VISC_NONE),
             AB200_ESTATE, "The chosen viscosity model does not work with damping");
  TwoFluids tf{};
  for (int f = 0; f < 2; ++f) {
    tf.on[f] = c->fl[f].bound ? 1 : 0;
    if (tf.on[f]) tf.f[f] = c->fl[f].d;  // the conserved state only: primitives are not read
  }
  DragDev dp{};
  dp.coupling = d->coupling; dp.model = d->model; dp.damp_to_visc = d->g_damp_to_visc;
  dp.scale = d->scale; dp.grain_density = d->grain_density;
  for (int n = 0; n < kMaxDustSpecies; ++n) { dp.tau[n] = d->tau[n]; dp.sizes[n] = d->sizes[n]; }
  for (int k = 0; k < 3; ++k) {
    dp.g_ix[k] = d->g_ix[k]; dp.g_ox[k] = d->g_ox[k];
    dp.g_irate[k] = d->g_irate[k]; dp.g_orate[k] = d->g_orate[k];
    dp.d_ix[k] = d->d_ix[k]; dp.d_ox[k] = d->d_ox[k];
    dp.d_irate[k] = d->d_irate[k]; dp.d_orate[k] = d->d_orate[k];
    dp.xmin[k] = d->xmin[k]; dp.xmax[k] = d->xmax[k];
  }
  const DiffDev dd = c->has_diffusion ? diff_dev(c) : DiffDev{};
  const GridDev &g = c->g;
  int rc = dispatch_geom_s(g.geom, [&](auto G) {
    k_drag<decltype(G)::value><<<grid_s(g), kThreads, 0, c->stream>>>(g, tf, dt, dt_dev, beta, dp, dd);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}
int ab200_drag_source(ab200_ctx *c, double dt, const ab200_drag_desc *d) {
  return drag_impl(c, dt, nullptr, 0.0, d);
}
// <dust/stopping_time> type = constant, no damping zones (inputs/drag/simple_drag.in)
static ab200_drag_desc constant_drag(int ntau, const double *tau) {
  ab200_drag_desc d{};
  d.coupling = AB200_DRAG_SIMPLE_DUST; d.model = AB200_DRAG_CONSTANT; d.scale = 1.0;
  for (int n = 0; n < ntau && n < kMaxDustSpecies; ++n) d.tau[n] = tau[n];
  const double big = 1.79769313486231570815e+308;
  for (int k = 0; k < 3; ++k) { d.g_ix[k] = d.d_ix[k] = -big; d.g_ox[k] = d.d_ox[k] = big; }
  return d;
}
int ab200_drag_simple(ab200_ctx *c, double dt, int ntau, const double *tau) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  AB_REQUIRE(tau && c->fl[1].bound && ntau == c->fl[1].d.S, AB200_EINVAL,
             "ab200_drag_simple: one stopping time per dust species");
  AB_REQUIRE(ntau <= kMaxDustSpecies, AB200_EINVAL, "ab200_drag_simple: too many dust species");
  const ab200_drag_desc d = constant_drag(ntau, tau);
  return drag_impl(c, dt, nullptr, 0.0, &d);
}


static int point_mass_impl(ab200_ctx *c, double dt, const double *dt_dev, double beta,
                           const ab200_point_mass_desc *d) {
  AB_ENTER_S(c)
  AB_REQUIRE(d != nullptr, AB200_EINVAL, "ab200_point_mass_gravity: null descriptor");
  TwoFluids tf{};
  AB_TRY(two_fluids(c, tf));
  PointMass pm{d->gm, {d->x, d->y, d->z}, d->soft * d->soft, d->sink_rate, d->sink};
  const GridDev &g = c->g;
  int rc = dispatch_geom_s(g.geom, [&](auto G) {
    k_point_mass<decltype(G)::value><<<grid_s(g), kThreads, 0, c->stream>>>(g, tf, dt, dt_dev, beta, pm);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}
int ab200_point_mass_gravity(ab200_ctx *c, double dt, const ab200_point_mass_desc *d) {
  return point_mass_impl(c, dt, nullptr, 0.0, d);
}

static int rotating_frame_impl(ab200_ctx *c, double dt, const double *dt_dev, double beta,
                               double omega) {
  AB_ENTER_S(c)
  AB_REQUIRE(omega != 0.0, AB200_EINVAL, "rotating_frame/omega cannot be zero!");
  AB_REQUIRE(c->g.geom != AB200_CARTESIAN, AB200_EINVAL,
             "ab200_rotating_frame: the Cartesian rotating frame is ab200_shearing_box "
             "(rotating_frame.cpp:67-68)");
  TwoFluids tf{};
  TwoMassFlux mf{};
  // only the conserved state is touched; the primitives are not read
  int any = 0;
  for (int f = 0; f < 2; ++f) {
    tf.on[f] = c->fl[f].bound ? 1 : 0;
    if (!tf.on[f]) continue;
    const FluidHost &fh = c->fl[f];
    AB_REQUIRE(fh.dflux_src != 0, AB200_ESTATE,
               "ab200_rotating_frame: no mass fluxes of this stage (run ab200_fused_stage with "
               "AB200_STAGE_TAP_DFLUX or ab200_calculate_fluxes first)");
    tf.f[f] = fh.d;
    for (int d = 0; d < 3; ++d) mf.m[f].tab[d] = fh.dflux_src == 1 ? fh.d.dflux[d] : fh.d.flux[d];
    mf.m[f].stride = fh.dflux_src == 1 ? fh.d.S : fh.d.nvar;
    any = 1;
  }
  AB_REQUIRE(any, AB200_ESTATE, "source term: no fluid bound");
  const GridDev &g = c->g;
  int rc = dispatch_geom_s(g.geom, [&](auto G) {
    constexpr int GG = decltype(G)::value;
    if constexpr (GG != AB200_CARTESIAN)
      k_rotating_frame<GG><<<grid_s(g), kThreads, 0, c->stream>>>(g, tf, mf, dt, dt_dev, beta, omega);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}
int ab200_rotating_frame(ab200_ctx *c, double dt, double omega) {
  return rotating_frame_impl(c, dt, nullptr, 0.0, omega);
}

int ab200_configure_sources(ab200_ctx *c, const ab200_sources_desc *src) {
  AB_REQUIRE(c, AB200_EINVAL, "null context");
  if (!src) {
    c->sources = ab200_sources_desc{};
    c->has_sources = false;
    return AB200_OK;
  }
  AB_REQUIRE(!src->drag || (src->ntau >= 1 && src->ntau <= 16), AB200_EINVAL,
             "ab200_configure_sources: 1..16 stopping times");
  c->sources = *src;
  AB_REQUIRE(!(src->gravity && src->point_mass), AB200_EINVAL,
             "ab200_configure_sources: one gravity type (gravity.cpp:68-88)");
  AB_REQUIRE(!(src->shearing_box && src->rotating_frame), AB200_EINVAL,
             "ab200_configure_sources: shearing box (Cartesian) or rotating frame (curvilinear)");
  AB_REQUIRE(!(src->drag && src->drag_model), AB200_EINVAL,
             "ab200_configure_sources: drag (constant tau) or drag_model (full descriptor)");
  c->has_sources = src->gravity || src->shearing_box || src->drag || src->point_mass ||
                   src->rotating_frame || src->drag_model;
  return AB200_OK;
}

}  // extern "C"

namespace ab200 {
// ArtemisDriver::StepTasks for one stage on the device-resident path
// (src/artemis_driver.cpp:184-255): dt comes from the device scalar.
int run_stage(ab200_ctx *c, double g0, double g1, double beta, int pcm, int first, int last) {
  const int base = AB200_STAGE_DEVICE_DT | AB200_STAGE_PINGPONG;
  if (!c->has_sources && !c->has_diffusion)
    return ab200_fused_stage(c, g0, g1, beta, 0.0, pcm, first, base | (last ? AB200_STAGE_REDUCE_DT : 0));
  const ab200_sources_desc &s = c->sources;
  // diffusion fluxes come from the stage-start primitives, which the deferred stage leaves alone
  if (c->has_diffusion) AB_TRY(launch_diffusion_flux(c));
  AB_TRY(ab200_fused_stage(c, g0, g1, beta, 0.0, pcm, first,
                           base | AB200_STAGE_DEFER_C2P |
                               (s.rotating_frame ? AB200_STAGE_TAP_DFLUX : 0)));
  // beta * dt is formed on the device from the dt scalar: no host round trip
  if (c->has_diffusion) AB_TRY(launch_diffusion_update(c, 0.0, c->d_time, beta));
  if (s.gravity) AB_TRY(gravity_impl(c, 0.0, c->d_time, beta, s.g[0], s.g[1], s.g[2]));
  if (s.point_mass) AB_TRY(point_mass_impl(c, 0.0, c->d_time, beta, &s.pm));
  if (s.shearing_box) AB_TRY(shearing_impl(c, 0.0, c->d_time, beta, s.omega, s.qshear));
  if (s.rotating_frame) AB_TRY(rotating_frame_impl(c, 0.0, c->d_time, beta, s.rf_omega));
  if (s.drag_model) {
    AB_TRY(drag_impl(c, 0.0, c->d_time, beta, &s.drag_desc));
  } else if (s.drag) {
    const ab200_drag_desc d = constant_drag(s.ntau, s.tau);
    AB_TRY(drag_impl(c, 0.0, c->d_time, beta, &d));
  }
  return ab200_finish_stage(c, last ? AB200_STAGE_REDUCE_DT : 0);
}
}  // namespace ab200
