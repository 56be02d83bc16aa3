// diffcoef.cuh -- transport coefficients shared by the diffusion operators (diffusion.cu) and
// the damping zones of the drag source (sources.cu): Diffusion::DiffCoeffParams and
// DiffusionCoeff<viscosity_*>::Get (src/utils/diffusion/diffusion_coeff.hpp:59-268).
#pragma once
#include "tasks.cuh"

namespace ab200 {

struct DiffDev {  // Diffusion::DiffCoeffParams of <gas/viscosity> and <gas/conductivity>
  int visc_type, visc_avg, cond_type, cond_avg;
  double nu, eta, r0, r_exp, alpha, omega0;
  double cond, kappa, temp_exp, rho_exp, rho_ref, t_ref, cv;
  double *flx[3];  // [nb][4S][fnk][fnj][fni] per direction
};

// dynamic viscosity rho nu of one zone from its density and specific internal energy
// (DiffusionCoeff<viscosity_plaw / viscosity_alpha>::Get, diffusion_coeff.hpp:178-268)
template <int GEOM>
AB_D double visc_mu_val(const Coords<GEOM> &c, const DiffDev &dd, double gm1, double dens,
                        double sie) {
  if (dd.visc_type == AB200_VISC_PLAW) return dd.nu * dens * pow(c.cyl_radius() / dd.r0, dd.r_exp);
  const double Omk = dd.omega0 * pow(c.sph_radius() / dd.r0, -1.5);
  const double blk = dmax(0.0, (gm1 + 1) * gm1 * dens * sie);
  return dd.alpha * blk / Omk;
}

inline DiffDev diff_dev(const ab200_ctx *c) {
  const ab200_diffusion_desc &s = c->diffusion;
  DiffDev d{};
  d.visc_type = s.visc_type; d.visc_avg = s.visc_avg;
  d.cond_type = s.cond_type; d.cond_avg = s.cond_avg;
  d.nu = s.nu; d.eta = s.eta_bulk; d.r0 = s.r0; d.r_exp = s.r_exp;
  d.alpha = s.alpha; d.omega0 = s.omega0;
  d.cond = s.cond; d.kappa = s.kappa; d.temp_exp = s.temp_exp; d.rho_exp = s.rho_exp;
  d.rho_ref = s.rho_ref; d.t_ref = s.t_ref; d.cv = s.cv;
  for (int k = 0; k < 3; ++k) d.flx[k] = c->d_dflx[k];
  return d;
}

}  // namespace ab200
