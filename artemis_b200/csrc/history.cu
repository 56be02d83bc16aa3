// history.cu -- device-side history reductions: the volume integrals the
// reference writes to its .hst file every output cadence,
//   ArtemisUtils::ReduceSpeciesVolumeIntegral<GEOM, VAR>        src/utils/history.hpp:24-57
//   ArtemisUtils::ReduceSpeciesVectorVolumeIntegral<GEOM, DIR>  src/utils/history.hpp:59-95
// registered as gas_mass, gas_momentum_x{1,2,3}, gas_energy, gas_internal_energy (and the dust
// ones) in src/gas/gas.cpp:650-675.  The reference launches one par_reduce PER SPECIES AND
// VARIABLE, each followed by a device -> host copy of one scalar; here ONE launch integrates
// every conserved pack entry of a fluid (sum over interior zones of u * Volume) and one small
// kernel finishes the sum in a fixed order (deterministic, unlike an atomic or Kokkos::Sum
// tree), so an output cadence costs two launches and one copy.
#include <type_traits>

#include "tasks.cuh"

namespace ab200 {

constexpr int kHistBlocks = 592;  // 4 CTAs per SM on 148 SMs

template <int GEOM>
__global__ void __launch_bounds__(kThreads)
k_history_partial(GridDev g, FluidDev f, double *partial) {
  const int v = blockIdx.y;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const long long total = (long long)g.nb * nkr * njr * nir;
  double acc = 0.0;
  for (long long t = blockIdx.x * (long long)blockDim.x + threadIdx.x; t < total;
       t += (long long)gridDim.x * blockDim.x) {
    const CellIdx c = decode(t, nir, njr, nkr, g.is, g.js, g.ks);
    Coords<GEOM> cc(g, c.b, c.k, c.j, c.i);
    const size_t off = ((size_t)c.k * g.nj + c.j) * g.ni + c.i;
    acc += f.u0[(size_t)c.b * f.nvar + v][off] * cc.volume();
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  __shared__ double sm[kThreads / 32];
  if ((threadIdx.x & 31) == 0) sm[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int w = 0; w < kThreads / 32; ++w) s += sm[w];
    partial[(size_t)v * gridDim.x + blockIdx.x] = s;
  }
}

__global__ void k_history_final(const double *partial, int nblocks, double *out) {
  const int v = blockIdx.x;
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int b = 0; b < nblocks; ++b) s += partial[(size_t)v * nblocks + b];
    out[v] = s;
  }
}

}  // namespace ab200

using namespace ab200;

extern "C" int ab200_history_volume_integrals(ab200_ctx *c, int fluid, double *out_host, int nout) {
  AB_REQUIRE(c && c->grid_set, AB200_ESTATE, "ab200_history_volume_integrals: no grid bound");
  AB_REQUIRE(fluid == 0 || fluid == 1, AB200_EINVAL, "Fluid type not recognized!");
  AB_REQUIRE(c->fl[fluid].bound, AB200_ESTATE, "fluid pack not bound: call ab200_bind_pack");
  const FluidDev &f = c->fl[fluid].d;
  AB_REQUIRE(out_host && nout == f.nvar, AB200_EINVAL,
             "ab200_history_volume_integrals: one output per conserved pack entry");
  AB_CUDA(cudaSetDevice(c->device));
  const GridDev &g = c->g;
  double *scratch = nullptr;
  AB_CUDA(cudaMalloc((void **)&scratch, sizeof(double) * ((size_t)f.nvar * kHistBlocks + f.nvar)));
  dim3 grid(kHistBlocks, (unsigned)f.nvar);
  NvtxRange nvtx_("ReduceSpeciesVolumeIntegral / ReduceSpeciesVectorVolumeIntegral [all pack entries]");
  switch (g.geom) {
#define AB_H(G) case G: k_history_partial<G><<<grid, kThreads, 0, c->stream>>>(g, f, scratch); break;
    AB_H(0) AB_H(1) AB_H(2) AB_H(3) AB_H(4) AB_H(5)
#undef AB_H
  }
  double *d_out = scratch + (size_t)f.nvar * kHistBlocks;
  k_history_final<<<f.nvar, 32, 0, c->stream>>>(scratch, kHistBlocks, d_out);
  c->launches += 2;
  AB_CUDA(cudaGetLastError());
  AB_CUDA(cudaMemcpyAsync(out_host, d_out, sizeof(double) * f.nvar, cudaMemcpyDeviceToHost, c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));
  AB_CUDA(cudaFree(scratch));
  return AB200_OK;
}
