// tma_maps.cu -- host side of the TMA staging used by the fused passes: one CUtensorMap per
// (array kind, direction, block, pack entry) describing the 3-D array [nk][nj][ni] and the
// pencil-tile box the fused kernel pulls into shared memory with cp.async.bulk.tensor.
//
// cuTensorMapEncodeTiled is fetched through cudaGetDriverEntryPoint so the library does not
// link against libcuda (it must load on a machine without a driver for the ABI tests).
#include <cuda.h>

#include <cstdlib>
#include <cstring>

#include "ab200_ctx.cuh"

namespace ab200 {

typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  static bool tried = false;
  if (!tried) {
    tried = true;
    void *p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)p;
    else
      cudaGetLastError();
  }
  return fn;
}

void *tma_encode_fn() { return (void *)get_encode(); }

// pencils per CTA for direction `dir` (1..3) given the interior extent L along it
int tma_pencils(const ab200_ctx *c, int dir, int L, int max_threads) {
  int np = max_threads / (L + 2);
  if (const char *env = getenv("AB200_FUSED_NP")) {
    const int v = atoi(env);
    if (v >= 1 && v < np) np = v;
  }
  if (dir != 1) np &= ~1;  // inner box extent NP*8 B must be a multiple of 16 B
  (void)c;
  return np;
}

void release_tma(FluidHost &fh) {
  for (int d = 0; d < 3; ++d) {
    if (fh.tma_maps[d]) cudaFree(fh.tma_maps[d]);
    fh.tma_maps[d] = nullptr;
    fh.tma_np[d] = 0;
  }
  fh.tma_ready = false;
}

// Builds (once per binding) the tensor maps of prim / u0 / u1 for every direction.
// Returns AB200_OK with fh.tma_ready == false when TMA cannot be used (the launcher then
// falls back to the L1-staged kernel): no driver entry point, odd ni, unaligned arrays.
int ensure_tma(ab200_ctx *c, int fluid, int max_threads) {
  FluidHost &fh = c->fl[fluid];
  if (fh.tma_tried) return AB200_OK;
  fh.tma_tried = true;
  fh.tma_ready = false;
  if (getenv("AB200_NO_TMA")) return AB200_OK;
  EncodeTiledFn enc = get_encode();
  if (!enc) return AB200_OK;
  const GridDev &g = c->g;
  if ((g.ni * 8) % 16 != 0) return AB200_OK;  // global strides must be multiples of 16 B
  const FluidDev &f = fh.d;
  if (!f.prim || !f.u0 || !f.u1) return AB200_OK;
  const size_t nent = (size_t)g.nb * f.nvar;
  std::vector<double *> tabs[3];
  double *const *src[3] = {f.prim, f.u0, f.u1};
  for (int w = 0; w < 3; ++w) {
    tabs[w].resize(nent);
    AB_CUDA(cudaMemcpy(tabs[w].data(), src[w], nent * sizeof(double *), cudaMemcpyDeviceToHost));
    for (size_t e = 0; e < nent; ++e)
      if (((uintptr_t)tabs[w][e]) % 16 != 0) return AB200_OK;
  }
  const int ext[3] = {g.ie - g.is + 1, g.je - g.js + 1, g.ke - g.ks + 1};
  const int nall[3] = {g.ni, g.nj, g.nk};
  for (int dir = 1; dir <= g.ndim; ++dir) {
    const int np = tma_pencils(c, dir, ext[dir - 1], max_threads);
    if (np < 1 || (dir != 1 && np < 2)) {
      release_tma(fh);
      return AB200_OK;
    }
    std::vector<CUtensorMap> maps(3 * nent);
    const cuuint64_t gdim[3] = {(cuuint64_t)g.ni, (cuuint64_t)g.nj, (cuuint64_t)g.nk};
    const cuuint64_t gstr[2] = {(cuuint64_t)g.ni * 8, (cuuint64_t)g.ni * g.nj * 8};
    cuuint32_t box[3];
    if (dir == 1) { box[0] = g.ni; box[1] = np; box[2] = 1; }
    if (dir == 2) { box[0] = np; box[1] = g.nj; box[2] = 1; }
    if (dir == 3) { box[0] = np; box[1] = 1; box[2] = g.nk; }
    if (box[0] > 256 || box[1] > 256 || box[2] > 256) {
      release_tma(fh);
      return AB200_OK;
    }
    const cuuint32_t estr[3] = {1, 1, 1};
    for (int w = 0; w < 3; ++w)
      for (size_t e = 0; e < nent; ++e) {
        CUresult r = enc(&maps[w * nent + e], CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, tabs[w][e],
                         gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                         CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                         CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) {
          release_tma(fh);
          return AB200_OK;
        }
      }
    void *d = nullptr;
    AB_CUDA(cudaMalloc(&d, maps.size() * sizeof(CUtensorMap)));
    AB_CUDA(cudaMemcpy(d, maps.data(), maps.size() * sizeof(CUtensorMap), cudaMemcpyHostToDevice));
    fh.tma_maps[dir - 1] = d;
    fh.tma_np[dir - 1] = np;
    (void)nall;
  }
  fh.tma_ready = true;
  return AB200_OK;
}

}  // namespace ab200
