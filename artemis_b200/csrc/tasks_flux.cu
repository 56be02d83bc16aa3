// tasks_flux.cu -- instantiations of k_calculate_fluxes for ONE coordinate system
// (compiled six times with -DAB_GEOM=0..5 so the translation units build in parallel).
#include <type_traits>

#include "tasks.cuh"

#ifndef AB_GEOM
#error "compile with -DAB_GEOM=<0..5>"
#endif

namespace ab200 {

template <int GEOM, int FLUID, int RS, int RC>
static int launch_dir(ab200_ctx *c, const FluidDev &f) {
  const GridDev &g = c->g;
  for (int dir = 1; dir <= g.ndim; ++dir) {
    const long long nir = g.ie - g.is + 1 + (dir == 1), njr = g.je - g.js + 1 + (dir == 2),
                    nkr = g.ke - g.ks + 1 + (dir == 3);
    const long long total = (long long)g.nb * nkr * njr * nir;
    const unsigned grid = (unsigned)((total + kThreads - 1) / kThreads);
    if (dir == 1) k_calculate_fluxes<GEOM, FLUID, RS, RC, 1><<<grid, kThreads, 0, c->stream>>>(g, f);
    if (dir == 2) k_calculate_fluxes<GEOM, FLUID, RS, RC, 2><<<grid, kThreads, 0, c->stream>>>(g, f);
    if (dir == 3) k_calculate_fluxes<GEOM, FLUID, RS, RC, 3><<<grid, kThreads, 0, c->stream>>>(g, f);
    c->launches++;
  }
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

template <int GEOM, int FLUID, int RS>
static int launch_rc(ab200_ctx *c, const FluidDev &f, int recon) {
  switch (recon) {
  case AB200_PCM: return launch_dir<GEOM, FLUID, RS, AB200_PCM>(c, f);
  case AB200_PLM: return launch_dir<GEOM, FLUID, RS, AB200_PLM>(c, f);
  case AB200_PPM: return launch_dir<GEOM, FLUID, RS, AB200_PPM>(c, f);
  }
  set_error("Reconstruction method not recognized!");
  return AB200_EINVAL;
}

template <int GEOM>
int launch_flux_geom(ab200_ctx *c, int fluid, int pcm);

template <>
int launch_flux_geom<AB_GEOM>(ab200_ctx *c, int fluid, int pcm) {
  const FluidDev &f = c->fl[fluid].d;
  // fluid_fluxes.hpp:225: pcm flag (vl2 stage 1) overrides the package's method
  const int recon = pcm ? AB200_PCM : f.recon;
  if (fluid == AB200_GAS) {
    switch (f.riemann) {
    case AB200_HLLC: return launch_rc<AB_GEOM, AB200_GAS, AB200_HLLC>(c, f, recon);
    case AB200_HLLE: return launch_rc<AB_GEOM, AB200_GAS, AB200_HLLE>(c, f, recon);
    case AB200_LLF: return launch_rc<AB_GEOM, AB200_GAS, AB200_LLF>(c, f, recon);
    }
  } else {
    switch (f.riemann) {
    case AB200_HLLE: return launch_rc<AB_GEOM, AB200_DUST, AB200_HLLE>(c, f, recon);
    case AB200_LLF: return launch_rc<AB_GEOM, AB200_DUST, AB200_LLF>(c, f, recon);
    }
  }
  set_error("Riemann solver not recognized!");
  return AB200_EINVAL;
}

}  // namespace ab200
