// tma_probe.cu -- which {box, array} shapes does cp.async.bulk.tensor.3d accept for fp64 planes?
// usage: tma_probe ni nj nk bx by c0 c1 c2   (one variant per process: a fault kills the context)
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cstdint>
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *,
                                  const cuuint64_t *, const cuuint64_t *, const cuuint32_t *,
                                  const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void k(const CUtensorMap *map, double *out, int n, int c0, int c1, int c2) {
  extern __shared__ __align__(128) unsigned char raw[];
  uint64_t *bar = reinterpret_cast<uint64_t *>(raw);
  double *dst = reinterpret_cast<double *>(raw + 256);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(1));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(n * 8) : "memory");
    asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];"
                 ::"r"(smem_u32(dst)), "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2) : "memory");
  }
  asm volatile("{\n.reg .pred p;\nW:\nmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(0) : "memory");
  for (int t = threadIdx.x; t < n; t += blockDim.x) out[t] = dst[t];
}
int main(int argc, char **argv) {
  int ni = atoi(argv[1]), nj = atoi(argv[2]), nk = atoi(argv[3]), bx = atoi(argv[4]), by = atoi(argv[5]);
  int c0 = atoi(argv[6]), c1 = atoi(argv[7]), c2 = atoi(argv[8]);
  void *p = nullptr; cudaDriverEntryPointQueryResult q;
  cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q);
  EncodeTiledFn enc = (EncodeTiledFn)p;
  size_t cells = (size_t)ni * nj * nk;
  std::vector<double> h(cells);
  for (size_t t = 0; t < cells; ++t) h[t] = (double)t;
  double *d, *out; cudaMalloc(&d, cells * 8); cudaMemcpy(d, h.data(), cells * 8, cudaMemcpyHostToDevice);
  int n = bx * by; cudaMalloc(&out, n * 8);
  CUtensorMap m;
  cuuint64_t gdim[3] = {(cuuint64_t)ni, (cuuint64_t)nj, (cuuint64_t)nk}, gstr[2] = {(cuuint64_t)ni * 8, (cuuint64_t)ni * nj * 8};
  cuuint32_t box[3] = {(cuuint32_t)bx, (cuuint32_t)by, 1}, es[3] = {1, 1, 1};
  CUresult r = enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) { printf("encode failed %d\n", (int)r); return 2; }
  CUtensorMap *dm; cudaMalloc(&dm, sizeof(m)); cudaMemcpy(dm, &m, sizeof(m), cudaMemcpyHostToDevice);
  k<<<1, 128, 256 + n * 8 + 128>>>(dm, out, n, c0, c1, c2);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) { printf("ni=%d nj=%d nk=%d box=%dx%d at (%d,%d,%d): FAULT %s\n", ni, nj, nk, bx, by, c0, c1, c2, cudaGetErrorString(e)); return 1; }
  std::vector<double> o(n); cudaMemcpy(o.data(), out, n * 8, cudaMemcpyDeviceToHost);
  int bad = 0;
  for (int y = 0; y < by; ++y) for (int x = 0; x < bx; ++x) {
    int gi = c0 + x, gj = c1 + y;
    double want = (gi < 0 || gi >= ni || gj < 0 || gj >= nj) ? 0.0 : (double)(((size_t)c2 * nj + gj) * ni + gi);
    if (o[y * bx + x] != want) ++bad;
  }
  printf("ni=%d nj=%d nk=%d box=%dx%d at (%d,%d,%d): ok, %d mismatches\n", ni, nj, nk, bx, by, c0, c1, c2, bad);
  return 0;
}
