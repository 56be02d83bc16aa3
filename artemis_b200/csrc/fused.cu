// fused.cu -- instantiations + launcher of the fused directional passes for ONE coordinate
// system (compiled six times with -DAB_GEOM=0..5).
#include <cstdlib>

#include "fused.cuh"
#include "march.cuh"
#include "xchunk.cuh"
#include "xtile.cuh"

#ifndef AB_GEOM
#error "compile with -DAB_GEOM=<0..5>"
#endif

namespace ab200 {

template <int GEOM, int FLUID, int RS, int RC, int DIR>
static int launch_pass(ab200_ctx *c, const FluidDev &f, FusedArgs a) {
  const GridDev &g = c->g;
  const FluidHost &fh = c->fl[FLUID];
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int L = DIR == 1 ? nir : (DIR == 2 ? njr : nkr);
  const int nL = DIR == 1 ? g.ni : (DIR == 2 ? g.nj : g.nk);
  AB_REQUIRE(L + 2 <= kFusedMaxThreads, AB200_EINVAL,
             "ab200_fused_stage: MeshBlock extent exceeds 510 zones; use the task-level path");
  constexpr int NV = FLUID == AB200_GAS ? 6 : 4, NF = FLUID == AB200_GAS ? 8 : 4;
  const int npencils = DIR == 1 ? nkr * njr : (DIR == 2 ? nkr * nir : njr * nir);  // per block
  if (fh.tma_ready && fh.tma_np[DIR - 1] > 0) {
    // ---- TMA-staged kernel ---------------------------------------------------------------
    const int np = fh.tma_np[DIR - 1];
    const int tiled = DIR == 1 ? njr : nir, rows = DIR == 3 ? njr : nkr;
    a.np = np;
    a.npencils = npencils;
    a.tiles_per_row = (tiled + np - 1) / np;
    a.tiles_per_block = a.tiles_per_row * rows;
    a.nwork = g.nb * a.tiles_per_block * f.S;
    a.maps = reinterpret_cast<const CUtensorMap *>(fh.tma_maps[DIR - 1]);
    const int nthreads = ((np * (L + 2) + 31) / 32) * 32;
    const bool need_u1 = a.first && !a.copy_u1;
    const size_t tile_stride = (((size_t)np * nL * 8 + 127) / 128) * 128;
    const size_t shmem = 128 + 2 * (need_u1 ? 3 : 2) * NV * tile_stride +
                         sizeof(double) * ((size_t)(NV + NF) * np * (L + 1) + 3 * (size_t)f.nvar);
    auto kern = k_fused_pass<GEOM, FLUID, RS, RC, DIR, true>;
    if (shmem > 48 * 1024)
      AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    // persistent CTAs: 2 per SM, each streams work items through a 2-stage TMA pipeline
    int ncta = 2 * c->sm_count;
    if (const char *env = getenv("AB200_FUSED_CTAS")) ncta = atoi(env) > 0 ? atoi(env) : ncta;
    if (ncta > a.nwork) ncta = a.nwork;
    dim3 grid((unsigned)ncta, 1);
    kern<<<grid, nthreads, shmem, c->stream>>>(g, f, a);
  } else {
    // ---- fallback: stencil straight from global memory through L1 -------------------------
    int np = kFusedMaxThreads / (L + 2);
    if (const char *env = getenv("AB200_FUSED_NP")) {  // tuning knob: pencils per CTA
      const int v = atoi(env);
      if (v >= 1 && v <= np) np = v;
    }
    if (np > npencils) np = npencils;
    a.np = np;
    a.npencils = npencils;
    a.tiles_per_row = 0;
    a.maps = nullptr;
    const int nthreads = ((np * (L + 2) + 31) / 32) * 32;
    const size_t shmem = sizeof(double) * ((size_t)(NV + NF) * np * (L + 1) + 3 * (size_t)f.nvar);
    auto kern = k_fused_pass<GEOM, FLUID, RS, RC, DIR, false>;
    if (shmem > 48 * 1024)
      AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    dim3 grid((unsigned)((npencils + np - 1) / np), (unsigned)g.nb);
    kern<<<grid, nthreads, shmem, c->stream>>>(g, f, a);
  }
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

// x2 / x3 passes as marching kernels (march.cuh): one thread per pencil, register window.
template <int GEOM, int FLUID, int RS, int RC, int DIR>
static int launch_march(ab200_ctx *c, const FluidDev &f, FusedArgs a) {
  NvtxRange nvtx_(DIR == 2 ? "CalculateFluxes::X2-Flux + ApplyUpdate + GeometricSourceTerms [fused]" : (a.last ? "Hydro::X3-Flux + ApplyUpdate + GeometricSourceTerms + SetAuxillaryFields + ConsToPrim + PrimToCons [fused]" : "Hydro::X3-Flux + ApplyUpdate + GeometricSourceTerms [fused]"));
  const GridDev &g = c->g;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int ncols = nir * (DIR == 2 ? nkr : njr);
  dim3 grid((unsigned)((ncols + kMarchThreads - 1) / kMarchThreads),
            (unsigned)(a.blist ? a.nbl : g.nb), (unsigned)f.S);
  if (a.last) k_march_pass<GEOM, FLUID, RS, RC, DIR, true><<<grid, kMarchThreads, 0, c->stream>>>(g, f, a);
  else k_march_pass<GEOM, FLUID, RS, RC, DIR, false><<<grid, kMarchThreads, 0, c->stream>>>(g, f, a);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

// x1 pass as the warp-autonomous streaming kernel (xchunk.cuh); first pass of a >= 2-D stage.
template <int GEOM, int FLUID, int RS, int RC>
static int launch_xchunk(ab200_ctx *c, const FluidDev &f, FusedArgs a) {
  NvtxRange nvtx_("CalculateFluxes::X1-Flux + ApplyUpdate + GeometricSourceTerms [fused]");
  const GridDev &g = c->g;
  const int njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  int nseg = njr >= 32 ? 2 : 1;
  if (const char *env = getenv("AB200_XC_NSEG")) nseg = atoi(env) > 0 ? atoi(env) : nseg;
  if (nseg > njr) nseg = njr;
  a.np = nseg;
  constexpr int NV = FLUID == AB200_GAS ? 6 : 4, NF = FLUID == AB200_GAS ? 8 : 4;
  const long long nwarps = (long long)(a.blist ? a.nbl : g.nb) * f.S * nkr * nseg;
  const size_t shmem = sizeof(double) * kXcWarps * (2 * NV + NF) * 64;
  const unsigned grid = (unsigned)((nwarps + kXcWarps - 1) / kXcWarps);
  k_xchunk_pass<GEOM, FLUID, RS, RC><<<grid, kXcWarps * 32, shmem, c->stream>>>(g, f, a);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

// x1 pass as the thread-per-cell tile kernel (xtile.cuh)
template <int GEOM, int FLUID, int RS, int RC>
static int launch_xtile(ab200_ctx *c, const FluidDev &f, FusedArgs a) {
  NvtxRange nvtx_("CalculateFluxes::X1-Flux + ApplyUpdate + GeometricSourceTerms [fused]");
  const GridDev &g = c->g;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  int R = kXtThreads / (nir + 2);
  if (R > njr) R = njr;
  a.np = R;
  constexpr int NV = FLUID == AB200_GAS ? 6 : 4, NF = FLUID == AB200_GAS ? 8 : 4;
  const size_t shmem = sizeof(double) * (size_t)(NV + NF) * R * (nir + 1);
  const long long ncta = (long long)(a.blist ? a.nbl : g.nb) * f.S * nkr * ((njr + R - 1) / R);
  auto kern = k_xtile_pass<GEOM, FLUID, RS, RC>;
  if (shmem > 48 * 1024)
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
  kern<<<(unsigned)ncta, kXtThreads, shmem, c->stream>>>(g, f, a);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

// which x1 kernel: AB200_X1=tile | chunk.  Default: chunk -- measured on B200 (r02, 256^3 blast)
// the thread-per-cell tile kernel (72 registers, 27 warps per SM) takes 1.2 ms per pass against
// 0.8 ms for the lane-pipelined streaming kernel (160 registers, 12 warps): on this path the
// fp64 dependency chains are hidden by per-thread ILP, not by occupancy.
static bool use_xtile(const ab200_ctx *c) {
  static int v = -1;
  if (v < 0) {
    const char *e = getenv("AB200_X1");
    v = (e && e[0] == 't') ? 1 : 0;
  }
  const int nir = c->g.ie - c->g.is + 1;
  return v == 1 && nir + 2 <= kXtThreads && nir >= 8;
}

static bool use_xchunk() {
  static int v = -1;
  if (v < 0) v = getenv("AB200_NO_XCHUNK") ? 0 : 1;
  return v == 1;
}

static bool use_march() {
  static int v = -1;
  if (v < 0) v = getenv("AB200_NO_MARCH") ? 0 : 1;
  return v == 1;
}

template <int GEOM, int FLUID, int RS, int RC>
static int launch_dirs(ab200_ctx *c, const FluidDev &f, FusedArgs a) {
  const int ndim = c->g.ndim;
  AB_REQUIRE(!a.blist || (ndim >= 2 && use_xchunk() && use_march()), AB200_ESTATE,
             "block subsets need the streaming / marching passes");
  // Block subsets (AB200_STAGE_SURFACE / _INTERIOR) split only the LAST directional pass: the
  // SURFACE call runs the earlier passes over every block and the last pass over the surface
  // blocks, the INTERIOR call runs the last pass over the remaining blocks.  Passes of one
  // block never read another block, so the order is free; the remote exchange then overlaps
  // the interior share of the last pass (enough to hide it) at the price of ONE extra launch.
  const int *blist = a.blist;
  const int nbl = a.nbl;
  const bool interior_call = blist && a.subset == 2;
  a.blist = nullptr;
  const int copy = a.copy_u1;
  unsigned long long *dt_min = a.dt_min;
  a.dt_min = nullptr;  // only the marching kernel of the last direction folds the dt reduction
  const int fin = a.defer_c2p ? 0 : 1;  // deferred: SetAux / C2P run in ab200_finish_stage
  a.first = 1; a.last = fin * (ndim == 1); a.copy_u1 = copy;
  if (!interior_call) {
    if (ndim >= 2 && use_xchunk() && use_xtile(c) && !a.tap) AB_TRY((launch_xtile<GEOM, FLUID, RS, RC>(c, f, a)));
    else if (ndim >= 2 && use_xchunk()) AB_TRY((launch_xchunk<GEOM, FLUID, RS, RC>(c, f, a)));
    else AB_TRY((launch_pass<GEOM, FLUID, RS, RC, 1>(c, f, a)));
  }
  if (ndim >= 2) {
    a.first = 0; a.last = fin * (ndim == 2); a.copy_u1 = 0;
    a.dt_min = (ndim == 2 && fin) ? dt_min : nullptr;
    if (ndim == 2) { a.blist = blist; a.nbl = nbl; }
    if (ndim == 2 || !interior_call) {
      if (use_march()) AB_TRY((launch_march<GEOM, FLUID, RS, RC, 2>(c, f, a)));
      else AB_TRY((launch_pass<GEOM, FLUID, RS, RC, 2>(c, f, a)));
    }
  }
  if (ndim >= 3) {
    a.first = 0; a.last = fin; a.copy_u1 = 0;
    a.dt_min = fin ? dt_min : nullptr;
    a.blist = blist; a.nbl = nbl;
    if (use_march()) AB_TRY((launch_march<GEOM, FLUID, RS, RC, 3>(c, f, a)));
    else AB_TRY((launch_pass<GEOM, FLUID, RS, RC, 3>(c, f, a)));
  }
  return AB200_OK;
}

template <int GEOM, int FLUID, int RS>
static int launch_rc(ab200_ctx *c, const FluidDev &f, int recon, const FusedArgs &a) {
  switch (recon) {
  case AB200_PCM: return launch_dirs<GEOM, FLUID, RS, AB200_PCM>(c, f, a);
  case AB200_PLM: return launch_dirs<GEOM, FLUID, RS, AB200_PLM>(c, f, a);
  case AB200_PPM: return launch_dirs<GEOM, FLUID, RS, AB200_PPM>(c, f, a);
  }
  set_error("Reconstruction method not recognized!");
  return AB200_EINVAL;
}

template <int GEOM>
int launch_fused_geom(ab200_ctx *c, int fluid, const FusedArgs &a, int pcm);

template <>
int launch_fused_geom<AB_GEOM>(ab200_ctx *c, int fluid, const FusedArgs &a_in, int pcm) {
  FusedArgs a = a_in;
  if (!(use_march() && c->g.ndim >= 2)) a.dt_min = nullptr;
  AB_TRY(ensure_tma(c, fluid, kTmaMaxThreads));
  const FluidDev &f = c->fl[fluid].d;
  const int recon = pcm ? AB200_PCM : f.recon;
  if (fluid == AB200_GAS) {
    switch (f.riemann) {
    case AB200_HLLC: return launch_rc<AB_GEOM, AB200_GAS, AB200_HLLC>(c, f, recon, a);
    case AB200_HLLE: return launch_rc<AB_GEOM, AB200_GAS, AB200_HLLE>(c, f, recon, a);
    case AB200_LLF: return launch_rc<AB_GEOM, AB200_GAS, AB200_LLF>(c, f, recon, a);
    }
  } else {
    switch (f.riemann) {
    case AB200_HLLE: return launch_rc<AB_GEOM, AB200_DUST, AB200_HLLE>(c, f, recon, a);
    case AB200_LLF: return launch_rc<AB_GEOM, AB200_DUST, AB200_LLF>(c, f, recon, a);
    }
  }
  set_error("Riemann solver not recognized!");
  return AB200_EINVAL;
}

}  // namespace ab200
