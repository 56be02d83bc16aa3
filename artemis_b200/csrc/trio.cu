// trio.cu -- instantiations + launcher of the warp-specialised single-pass stage kernel
// (trio.cuh) for ONE Riemann solver (compiled three times with -DAB_RS=0..2 so the builds run
// in parallel).
#include "trio.cuh"

#ifndef AB_RS
#error "compile with -DAB_RS=<0..2>"
#endif

namespace ab200 {

template <int FLUID, int RS, int RC, int MODE>
static int launch_mode(ab200_ctx *c, const FluidDev &f, const SweepArgs &a) {
  constexpr int NV = FLUID == AB200_GAS ? 6 : 4, NF = FLUID == AB200_GAS ? 8 : 4;
  const size_t shmem = TrSmem<NV, NF>::bytes;
  auto kern = k_trio_stage<FLUID, RS, RC, MODE>;
  static bool attr_set = false;  // per instantiation
  if (!attr_set) {
    AB_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)shmem));
    attr_set = true;
  }
  dim3 grid((unsigned)(a.tiles_x * a.tiles_y), (unsigned)c->g.nb, (unsigned)f.S);
  kern<<<grid, kTrThreads, shmem, c->stream>>>(c->g, f, a);
  c->launches++;
  AB_CUDA(cudaGetLastError());
  return AB200_OK;
}

template <int FLUID, int RS, int RC>
static int launch_one(ab200_ctx *c, const FluidDev &f, const SweepArgs &a) {
  if (a.copy_u1) return launch_mode<FLUID, RS, RC, 0>(c, f, a);
  if (a.gam0 == 0.0) return launch_mode<FLUID, RS, RC, 1>(c, f, a);
  return launch_mode<FLUID, RS, RC, 2>(c, f, a);
}

template <int FLUID, int RS>
static int launch_rc(ab200_ctx *c, const FluidDev &f, int recon, const SweepArgs &a) {
  switch (recon) {
  case AB200_PCM: return launch_one<FLUID, RS, AB200_PCM>(c, f, a);
  case AB200_PLM: return launch_one<FLUID, RS, AB200_PLM>(c, f, a);
  case AB200_PPM: return launch_one<FLUID, RS, AB200_PPM>(c, f, a);
  }
  set_error("Reconstruction method not recognized!");
  return AB200_EINVAL;
}

template <int RS>
int launch_trio_rs(ab200_ctx *c, int fluid, int recon, const SweepArgs &a);

template <>
int launch_trio_rs<AB_RS>(ab200_ctx *c, int fluid, int recon, const SweepArgs &a) {
  const FluidDev &f = c->fl[fluid].d;
  if (fluid == AB200_GAS) return launch_rc<AB200_GAS, AB_RS>(c, f, recon, a);
#if AB_RS != 0  // HLLC is gas-only (src/dust/dust.cpp:76-85)
  return launch_rc<AB200_DUST, AB_RS>(c, f, recon, a);
#else
  set_error("Riemann solver (dust) not recognized.");
  return AB200_EINVAL;
#endif
}

}  // namespace ab200
