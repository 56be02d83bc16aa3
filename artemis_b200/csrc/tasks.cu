// tasks.cu -- launchers for the un-fused task kernels (everything except the flux kernels,
// which live in tasks_flux.cu, one translation unit per coordinate system).
#include <type_traits>

#include "tasks.cuh"

namespace ab200 {

template <int GEOM>
int launch_flux_geom(ab200_ctx *c, int fluid, int pcm);
template <> int launch_flux_geom<0>(ab200_ctx *, int, int);
template <> int launch_flux_geom<1>(ab200_ctx *, int, int);
template <> int launch_flux_geom<2>(ab200_ctx *, int, int);
template <> int launch_flux_geom<3>(ab200_ctx *, int, int);
template <> int launch_flux_geom<4>(ab200_ctx *, int, int);
template <> int launch_flux_geom<5>(ab200_ctx *, int, int);

template <typename F>
static int dispatch_geom(int geom, F &&fn) {
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

static unsigned grid_for(long long total) { return (unsigned)((total + kThreads - 1) / kThreads); }
static long long interior_cells(const GridDev &g) {
  return (long long)g.nb * (g.ke - g.ks + 1) * (g.je - g.js + 1) * (g.ie - g.is + 1);
}

int launch_calculate_fluxes(ab200_ctx *c, int fluid, int pcm) {
  NvtxRange nvtx_("CalculateFluxes::X1-Flux / X2-Flux / Hydro::X3-Flux");
  return dispatch_geom(c->g.geom, [&](auto G) {
    return launch_flux_geom<decltype(G)::value>(c, fluid, pcm);
  });
}

int launch_apply_update(ab200_ctx *c, int fluid, double gam0, double gam1, double beta_dt) {
  NvtxRange nvtx_("ApplyUpdate");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fluid].d;
  const unsigned grid = grid_for(interior_cells(g));
  int rc = dispatch_geom(g.geom, [&](auto G) {
    k_apply_update<decltype(G)::value><<<grid, kThreads, 0, c->stream>>>(g, f, gam0, gam1, beta_dt);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}

int launch_flux_source(ab200_ctx *c, int fluid, double dt) {
  NvtxRange nvtx_("GeometricSourceTerms");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fluid].d;
  const bool x1dep = g.geom != AB200_CARTESIAN;
  const bool x2dep = (g.geom == AB200_SPHERICAL2D || g.geom == AB200_SPHERICAL3D) && g.ndim >= 2;
  // Dust::FluxSource early-out, src/dust/dust.cpp:311-312
  if (fluid == AB200_DUST && !(x1dep || x2dep)) return AB200_OK;
  const unsigned grid = grid_for(interior_cells(g));
  const double omf = c->omf;
  int rc = dispatch_geom(g.geom, [&](auto G) {
    constexpr int GG = decltype(G)::value;
    if (fluid == AB200_GAS)
      k_flux_source<GG, AB200_GAS><<<grid, kThreads, 0, c->stream>>>(g, f, omf, dt);
    else
      k_flux_source<GG, AB200_DUST><<<grid, kThreads, 0, c->stream>>>(g, f, omf, dt);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}

int launch_set_aux(ab200_ctx *c) {
  NvtxRange nvtx_("SetAuxillaryFields");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[AB200_GAS].d;
  const unsigned grid = grid_for(interior_cells(g));
  int rc = dispatch_geom(g.geom, [&](auto G) {
    k_set_aux<decltype(G)::value><<<grid, kThreads, 0, c->stream>>>(g, f);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}

int launch_cons_to_prim(ab200_ctx *c, int fluid) {
  NvtxRange nvtx_("ConsToPrim");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fluid].d;
  const unsigned grid = grid_for(interior_cells(g));
  int rc = dispatch_geom(g.geom, [&](auto G) {
    constexpr int GG = decltype(G)::value;
    if (fluid == AB200_GAS)
      k_cons_to_prim<GG, AB200_GAS><<<grid, kThreads, 0, c->stream>>>(g, f);
    else
      k_cons_to_prim<GG, AB200_DUST><<<grid, kThreads, 0, c->stream>>>(g, f);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}

int launch_prim_to_cons(ab200_ctx *c, int fluid, int ghosts_only) {
  NvtxRange nvtx_("PrimToCons");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fluid].d;
  const unsigned grid = grid_for((long long)g.nb * g.nk * g.nj * g.ni);
  int rc = dispatch_geom(g.geom, [&](auto G) {
    constexpr int GG = decltype(G)::value;
    if (fluid == AB200_GAS)
      k_prim_to_cons<GG, AB200_GAS><<<grid, kThreads, 0, c->stream>>>(g, f, ghosts_only);
    else
      k_prim_to_cons<GG, AB200_DUST><<<grid, kThreads, 0, c->stream>>>(g, f, ghosts_only);
    return AB200_OK;
  });
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return rc;
}

__global__ void __launch_bounds__(kThreads) k_deep_copy(GridDev g, FluidDev f) {
  const size_t cells = (size_t)g.nk * g.nj * g.ni;
  const size_t total = cells * f.nvar * g.nb;
  for (size_t t = blockIdx.x * (size_t)blockDim.x + threadIdx.x; t < total;
       t += (size_t)gridDim.x * blockDim.x) {
    const size_t e = t / cells, o = t % cells;
    f.u1[e][o] = f.u0[e][o];
  }
}

int launch_deep_copy(ab200_ctx *c, int fluid) {
  NvtxRange nvtx_("DeepCopyConservedData");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fluid].d;
  k_deep_copy<<<c->sm_count * 8, kThreads, 0, c->stream>>>(g, f);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

// second stage of the dt reduction: min over per-CTA partials, times cfl
__global__ void k_finish_dt(const double *partial, int n, double cfl, double *out, int combine) {
  double v = 1.79769313486231570815e+308;
  for (int i = threadIdx.x; i < n; i += blockDim.x) v = dmin(v, partial[i]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, o));
  __shared__ double sm[32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x < 32) {
    v = threadIdx.x < (blockDim.x >> 5) ? sm[threadIdx.x] : 1.79769313486231570815e+308;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = dmin(v, __shfl_xor_sync(0xffffffffu, v, o));
    if (threadIdx.x == 0) {
      v = cfl * v;
      *out = combine ? dmin(*out, v) : v;
    }
  }
}

int launch_finish_dt(ab200_ctx *c, const double *partial, int n, double cfl, double *d_out,
                     int combine) {
  k_finish_dt<<<1, 256, 0, c->stream>>>(partial, n, cfl, d_out, combine);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

int launch_estimate_dt(ab200_ctx *c, int fluid, double *d_out, int combine) {
  NvtxRange nvtx_(fluid == AB200_GAS ? "Gas::EstimateTimestepMesh" : "Dust::EstimateTimestepMesh");
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fluid].d;
  long long total = interior_cells(g);
  int grid = (int)((total + kThreads - 1) / kThreads);
  const int maxgrid = 1024;  // <= 2048 partial slots in d_red
  if (grid > maxgrid) grid = maxgrid;
  double *partial = c->d_red;
  int rc = dispatch_geom(g.geom, [&](auto G) {
    constexpr int GG = decltype(G)::value;
    if (fluid == AB200_GAS)
      k_estimate_dt<GG, AB200_GAS><<<grid, kThreads, 0, c->stream>>>(g, f, partial);
    else
      k_estimate_dt<GG, AB200_DUST><<<grid, kThreads, 0, c->stream>>>(g, f, partial);
    return AB200_OK;
  });
  k_finish_dt<<<1, 256, 0, c->stream>>>(partial, grid, f.cfl, d_out, combine);
  c->launches += 2;
  AB_CUDA(cudaGetLastError());
  // Gas::EstimateTimestepMesh: cfl * min(hydro, viscous, conductive) (gas.cpp:437-467)
  if (rc == AB200_OK && fluid == AB200_GAS && c->has_diffusion)
    rc = launch_diffusion_dt(c, d_out, 1);
  return rc;
}

// EvolutionDriver::SetGlobalTimeStep on the device (P:driver/driver.cpp:210-269):
// t[0]=dt, t[1]=new block dt, t[2]=time, t[3]=ncycle
__global__ void k_set_global_dt(double *t, double tlim, int advance_time) {
  if (threadIdx.x || blockIdx.x) return;
  double dt = t[0];
  if (advance_time) {
    t[3] += 1.0;
    t[2] += dt;
  }
  if (dt < 0.1 * 1.79769313486231570815e+308) dt *= 2.0;
  dt = dmin(dt, t[1]);
  const double time = t[2];
  if (time < tlim && (tlim - time) < dt) dt = tlim - time;
  t[0] = dt;
}

int launch_set_global_dt(ab200_ctx *c, double tlim, int advance_time) {
  k_set_global_dt<<<1, 32, 0, c->stream>>>(c->d_time, tlim, advance_time);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

}  // namespace ab200
