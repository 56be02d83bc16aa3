// sweep_host.cu -- host side of the single-pass stage kernel (sweep.cuh): eligibility, the
// library-owned alternate primitive set and its tensor maps, the Riemann-solver dispatch and
// the copy that brings the current primitives back into the caller's arrays.
#include <cuda.h>

#include <cstdlib>

#include "trio.cuh"

namespace ab200 {

template <int RS>
int launch_trio_rs(ab200_ctx *c, int fluid, int recon, const SweepArgs &a);
template <> int launch_trio_rs<0>(ab200_ctx *, int, int, const SweepArgs &);
template <> int launch_trio_rs<1>(ab200_ctx *, int, int, const SweepArgs &);
template <> int launch_trio_rs<2>(ab200_ctx *, int, int, const SweepArgs &);

template <int RS>
int launch_sweep_rs(ab200_ctx *c, int fluid, int recon, const SweepArgs &a);
template <> int launch_sweep_rs<0>(ab200_ctx *, int, int, const SweepArgs &);
template <> int launch_sweep_rs<1>(ab200_ctx *, int, int, const SweepArgs &);
template <> int launch_sweep_rs<2>(ab200_ctx *, int, int, const SweepArgs &);

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

void release_sweep(FluidHost &fh) {
  for (int w = 0; w < 2; ++w) {
    if (fh.sw_maps[w]) cudaFree(fh.sw_maps[w]);
    fh.sw_maps[w] = nullptr;
  }
  fh.sw_ready = false;
}

// Allocates the alternate primitive set and builds the plane-tile tensor maps of both sets.
// Leaves fh.sw_ready == false (the caller then uses the three-pass path) when TMA cannot be
// used: no driver entry point, odd ni, arrays not 16-byte aligned.
static int ensure_sweep(ab200_ctx *c, int fluid) {
  FluidHost &fh = c->fl[fluid];
  if (fh.sw_tried) return AB200_OK;
  fh.sw_tried = true;
  fh.sw_ready = false;
  EncodeTiledFn enc = (EncodeTiledFn)tma_encode_fn();
  if (!enc) return AB200_OK;
  const GridDev &g = c->g;
  if ((g.ni * 8) % 16 != 0) return AB200_OK;
  FluidDev &f = fh.d;
  const size_t nent = (size_t)g.nb * f.nvar;
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  if ((cells * 8) % 16 != 0) return AB200_OK;
  std::vector<double *> tabs[2];
  tabs[0].resize(nent);
  AB_CUDA(cudaMemcpy(tabs[0].data(), f.prim, nent * sizeof(double *), cudaMemcpyDeviceToHost));
  for (size_t e = 0; e < nent; ++e)
    if (((uintptr_t)tabs[0][e]) % 16 != 0) return AB200_OK;
  // alternate set: one slab, zero-initialised
  double *slab = nullptr;
  if (cudaMalloc((void **)&slab, sizeof(double) * cells * nent) != cudaSuccess) {
    cudaGetLastError();
    return AB200_OK;  // not enough memory for the second set: three-pass path
  }
  fh.owned_scratch.push_back(slab);
  AB_CUDA(cudaMemsetAsync(slab, 0, sizeof(double) * cells * nent, c->stream));
  tabs[1].resize(nent);
  for (size_t e = 0; e < nent; ++e) tabs[1][e] = slab + e * cells;
  double **dtab = nullptr;
  AB_CUDA(cudaMalloc((void **)&dtab, nent * sizeof(double *)));
  fh.owned_tables.push_back(dtab);
  AB_CUDA(cudaMemcpy(dtab, tabs[1].data(), nent * sizeof(double *), cudaMemcpyHostToDevice));
  fh.prim_tab[0] = f.prim;
  fh.prim_tab[1] = dtab;
  fh.prim_cur = 0;
  const cuuint64_t gdim[3] = {(cuuint64_t)g.ni, (cuuint64_t)g.nj, (cuuint64_t)g.nk};
  const cuuint64_t gstr[2] = {(cuuint64_t)g.ni * 8, (cuuint64_t)g.ni * g.nj * 8};
  const cuuint32_t box[3] = {(cuuint32_t)kSwPI, (cuuint32_t)kSwPJ, 1};
  const cuuint32_t estr[3] = {1, 1, 1};
  for (int w = 0; w < 2; ++w) {
    std::vector<CUtensorMap> maps(nent);
    for (size_t e = 0; e < nent; ++e) {
      CUresult r = enc(&maps[e], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, tabs[w][e], gdim, gstr, box,
                       estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                       CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
      if (r != CUDA_SUCCESS) {
        release_sweep(fh);
        return AB200_OK;
      }
    }
    void *d = nullptr;
    AB_CUDA(cudaMalloc(&d, maps.size() * sizeof(CUtensorMap)));
    AB_CUDA(cudaMemcpy(d, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    fh.sw_maps[w] = d;
  }
  fh.sw_ready = true;
  return AB200_OK;
}

// AB200_PATH_AUTO policy, from the B200 measurements in profiles/r01d_stage_matrix.json (256^3
// zones, 64^3 blocks).  The single-pass kernel moves 2.2x less HBM traffic but runs one CTA of
// 10 warps per SM with three barriers per plane; it is bound by instruction issue and smem /
// dependency latency, not by HBM.  It beats the barrier-free directional passes only where the
// arithmetic per zone is light: LLF fluxes with PCM / PLM reconstruction (gas 1.53 vs 1.80 ms,
// dust 1.04 vs 1.17 ms per stage); for everything else (PPM, HLLC, HLLE) the passes win.
static bool sweep_preferred(const FluidDev &f) {
  return f.riemann == AB200_LLF && f.recon != AB200_PPM;
}

// Which of the two single-pass kernels runs: the warp-specialised one (trio.cuh) when asked for
// explicitly or chosen by AUTO, the one-role plane sweep (sweep.cuh) for AB200_PATH_SINGLE_PASS.
bool sweep_uses_role_split(const ab200_ctx *c, int fluid) {
  (void)fluid;
  // AUTO keeps the one-role kernel where it picks the single-pass family (LLF with PCM / PLM):
  // measured on B200 (r02) the warp-specialised kernel's x3+update group is the critical path
  // and the other two groups idle ~60 % of the time at the hand-over barriers
  return c->stage_path == AB200_PATH_ROLE_SPLIT;
}

bool sweep_eligible(ab200_ctx *c, int fluid) {
  static int env = -2;  // -1: forced off, 1: forced on, 0: no override
  if (env == -2) env = getenv("AB200_NO_SWEEP") ? -1 : (getenv("AB200_SWEEP") ? 1 : 0);
  if (env < 0 || c->stage_path == AB200_PATH_THREE_PASS) return false;
  const GridDev &g = c->g;
  if (g.ndim != 3 || g.geom != AB200_CARTESIAN) return false;
  if (c->stage_path == AB200_PATH_AUTO && env == 0 && !sweep_preferred(c->fl[fluid].d))
    return false;
  if (ensure_sweep(c, fluid) != AB200_OK) return false;
  return c->fl[fluid].sw_ready;
}

int launch_sweep_stage(ab200_ctx *c, int fluid, double gam0, double gam1, double beta, double dt,
                       int pcm, int stage1_copy, int use_device_dt, unsigned long long *dt_min) {
  NvtxRange nvtx_("CalculateFluxes + ApplyUpdate + GeometricSourceTerms + SetAuxillaryFields + ConsToPrim + PrimToCons [single pass]");
  FluidHost &fh = c->fl[fluid];
  AB_REQUIRE(fh.sw_ready, AB200_ESTATE, "launch_sweep_stage: sweep path not available");
  const GridDev &g = c->g;
  SweepArgs a{};
  a.gam0 = gam0; a.gam1 = gam1; a.beta = beta; a.dt = dt;
  a.dt_dev = use_device_dt ? c->d_time : nullptr;
  a.copy_u1 = stage1_copy;
  a.tiles_x = (g.ie - g.is + 1 + kSwTI - 1) / kSwTI;
  a.tiles_y = (g.je - g.js + 1 + kSwTJ - 1) / kSwTJ;
  const int in = fh.prim_cur, out = in ^ 1;
  a.maps = reinterpret_cast<const CUtensorMap *>(fh.sw_maps[in]);
  a.prim_out = fh.prim_tab[out];
  a.dt_min = dt_min;
  const int recon = pcm ? AB200_PCM : fh.d.recon;
  int rc = AB200_EINVAL;
  if (sweep_uses_role_split(c, fluid)) {  // warp-specialised kernel (trio.cuh)
    switch (fh.d.riemann) {
    case AB200_HLLC: rc = launch_trio_rs<0>(c, fluid, recon, a); break;
    case AB200_HLLE: rc = launch_trio_rs<1>(c, fluid, recon, a); break;
    case AB200_LLF: rc = launch_trio_rs<2>(c, fluid, recon, a); break;
    default: set_error("Riemann solver not recognized!");
    }
  } else {
    switch (fh.d.riemann) {
    case AB200_HLLC: rc = launch_sweep_rs<0>(c, fluid, recon, a); break;
    case AB200_HLLE: rc = launch_sweep_rs<1>(c, fluid, recon, a); break;
    case AB200_LLF: rc = launch_sweep_rs<2>(c, fluid, recon, a); break;
    default: set_error("Riemann solver not recognized!");
    }
  }
  AB_TRY(rc);
  fh.prim_cur = out;  // the new primitives live in the other set now
  fh.d.prim = fh.prim_tab[out];
  return AB200_OK;
}

// dst[e] <- src[e] for every (block, variable) array; interior zones only or whole arrays
static __global__ void __launch_bounds__(256)
k_copy_prim(GridDev g, double *const *dst, double *const *src, int interior_only) {
  const int e = blockIdx.y;
  const long long cells = (long long)g.ni * g.nj * g.nk;
  double *__restrict__ d = dst[e];
  const double *__restrict__ s = src[e];
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < cells;
       t += (long long)gridDim.x * blockDim.x) {
    if (interior_only) {
      const int i = (int)(t % g.ni), j = (int)((t / g.ni) % g.nj), k = (int)(t / ((long long)g.ni * g.nj));
      if (i < g.is || i > g.ie || j < g.js || j > g.je || k < g.ks || k > g.ke) continue;
    }
    d[t] = s[t];
  }
}

int sync_prim_home(ab200_ctx *c, int fluid, int interior_only) {
  FluidHost &fh = c->fl[fluid];
  if (!fh.bound || !fh.sw_ready || fh.prim_cur == 0) return AB200_OK;
  const GridDev &g = c->g;
  const long long cells = (long long)g.ni * g.nj * g.nk;
  int gx = (int)((cells + 256 * 8 - 1) / (256 * 8));
  if (gx < 1) gx = 1;
  // (block, variable) arrays in chunks of 65535: grid.y is a 16-bit dimension
  const long long nent = (long long)g.nb * fh.d.nvar;
  for (long long e0 = 0; e0 < nent; e0 += 65535) {
    const unsigned ny = (unsigned)((nent - e0) < 65535 ? (nent - e0) : 65535);
    dim3 grid((unsigned)gx, ny);
    k_copy_prim<<<grid, 256, 0, c->stream>>>(g, fh.prim_tab[0] + e0, fh.prim_tab[1] + e0,
                                             interior_only);
    c->launches++;
  }
  AB_CUDA(cudaGetLastError());
  fh.prim_cur = 0;
  fh.d.prim = fh.prim_tab[0];
  return AB200_OK;
}

}  // namespace ab200
