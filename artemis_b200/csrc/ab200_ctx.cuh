// ab200_ctx.cuh -- host-side context of libartemis_b200 (internal).
#pragma once
#include <cuda_runtime.h>

#include <cstring>
#include <string>
#include <vector>

#include "ab200_dev.cuh"

// NVTX ranges carrying the reference's own kernel / task labels (Parthenon wraps every par_for in
// a Kokkos profiling region named by its label, P:utils/instrument.hpp:22-49), so a timeline of
// this library lines up with one of the reference.  Header-only NVTX3: a no-op unless a profiler
// injects itself.
#include <nvtx3/nvToolsExt.h>

namespace ab200 {

struct NvtxRange {
  explicit NvtxRange(const char *label) { nvtxRangePushA(label); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange &) = delete;
};

void set_error(const std::string &msg);
int cuda_fail(cudaError_t e, const char *what, const char *file, int line);

#define AB_CUDA(call)                                                                    \
  do {                                                                                   \
    cudaError_t e__ = (call);                                                            \
    if (e__ != cudaSuccess) return ab200::cuda_fail(e__, #call, __FILE__, __LINE__);     \
  } while (0)
#define AB_REQUIRE(cond, code, msg)                                                      \
  do {                                                                                   \
    if (!(cond)) {                                                                       \
      ab200::set_error(msg);                                                             \
      return (code);                                                                     \
    }                                                                                    \
  } while (0)
#define AB_TRY(call)                                                                     \
  do {                                                                                   \
    int rc__ = (call);                                                                   \
    if (rc__ != AB200_OK) return rc__;                                                   \
  } while (0)

struct FluidHost {
  bool bound = false;
  FluidDev d{};                       // device-visible descriptor (pointer tables on device)
  std::vector<void *> owned_tables;   // device pointer tables we allocated
  std::vector<void *> owned_scratch;  // flux/pflux/vface/u1 scratch we allocated
  // ghost-exchange variable list (FillGhost fields): pack indices + vector component
  int n_ghost = 0;
  int *ghost_vars = nullptr;  // device
  int *ghost_vdir = nullptr;  // device
  // TMA staging of the fused passes: CUtensorMap[3 kinds][nb*nvar] per direction (device)
  bool tma_tried = false, tma_ready = false;
  void *tma_maps[3] = {nullptr, nullptr, nullptr};
  int tma_np[3] = {0, 0, 0};
  // single-pass sweep kernel (sweep.cuh): the stage reads one primitive set and writes the
  // other.  prim_tab[0] = the caller's arrays ("home"), prim_tab[1] = library-owned alternate
  // set; d.prim always points at prim_tab[prim_cur], the set that holds the current state.
  bool sw_tried = false, sw_ready = false;
  double *const *prim_tab[2] = {nullptr, nullptr};
  void *sw_maps[2] = {nullptr, nullptr};  // CUtensorMap[nb*nvar] per set, box {PI, PJ, 1}
  int prim_cur = 0;
  // lazy ghost cons (ab200_set_ghost_cons_lazy): the ghost fills wrote primitives only
  bool ghost_cons_stale = false;
  // where the mass fluxes of the latest stage live (ab200_rotating_frame reads them):
  // 0 nowhere, 1 the tap tables d.dflux (fused stage with AB200_STAGE_TAP_DFLUX), 2 the full
  // flux arrays d.flux (ab200_calculate_fluxes)
  int dflux_src = 0;
};
int ensure_tma(ab200_ctx *c, int fluid, int max_threads);
void release_tma(FluidHost &fh);
void release_sweep(FluidHost &fh);
void *tma_encode_fn();  // cuTensorMapEncodeTiled through the runtime, or nullptr

struct Topology {
  bool set = false;
  int nbx = 1, nby = 1, nbz = 1;
  int bc[6] = {0, 0, 0, 0, 0, 0};
};

}  // namespace ab200

struct ab200_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  cudaStream_t halo_stream = nullptr;  // ab200_set_halo_stream (pack / unpack kernels)
  bool halo_stream_set = false;
  bool grid_set = false;
  ab200::GridDev g{};
  ab200::GridDev gc{};          // coarse buffers of the multilevel operators (refine.cu)
  bool coarse_ready = false;
  // ab200_set_shear_bc_params: q and Omega_0 of the shearing-box `inflow` user condition
  double shear_q = 0.0, shear_om0 = 0.0;
  bool shear_bc_set = false;
  std::vector<void *> grid_allocs;
  std::vector<double> h_xmin, h_dx;
  ab200::FluidHost fl[2];
  ab200::Topology topo;
  double omf = 0.0;
  int stage_path = 0;  // AB200_PATH_*
  bool ghost_cons_lazy = false;  // ghost fills skip PrimToCons until ab200_sync_ghost_cons
  int host_transfer = 0;         // ab200_set_host_transfer flags (ab200_cycles_host)
  bool graph_replay = false;     // ab200_run_cycles captures one cycle into a CUDA graph (opt-in)
  long long graph_replays = 0;   // cycles executed as cudaGraphLaunch so far
  double *d_time = nullptr;     // device double[4]: dt, new_dt, time, ncycle
  double *d_red = nullptr;      // reduction scratch
  double *h_pinned = nullptr;   // pinned host scratch (8 doubles)
  long long launches = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  int sm_count = 148;
  // host-buffer path (ab200_cycles_host) owned device state
  std::vector<void *> host_path_allocs;
  bool host_path_ready = false;
  // device copies of halo descriptor lists, keyed by content (ab200_halo_pack / unpack)
  struct HaloCacheEntry {
    size_t bytes = 0;
    std::vector<unsigned char> host;
    void *dev = nullptr;
  };
  std::vector<HaloCacheEntry> halo_cache;
  // block subsets for compute / communication overlap: [0] surface blocks (touch a face owned
  // by another rank, AB200_BC_NONE), [1] all other blocks; built by ab200_set_topology
  int *d_blist[2] = {nullptr, nullptr};
  int n_blist[2] = {0, 0};
  // source terms applied by the device-resident drivers (ab200_configure_sources)
  ab200_sources_desc sources{};
  bool has_sources = false;
  // gas diffusion (ab200_configure_diffusion): parameters + library-owned face flux arrays
  ab200_diffusion_desc diffusion{};
  bool has_diffusion = false;
  double *d_dflx[3] = {nullptr, nullptr, nullptr};
  size_t dflx_elems = 0;
  double *d_dcoef = nullptr;  // [3][nb][S][cells]: viscosity, conductivity, div(u) per zone
  size_t dcoef_elems = 0;
  // multi-rank transport (comm.cu): NCCL communicator, comm stream, planned exchange
  void *comm_state = nullptr;
};

namespace ab200 {
// Device copy of a host descriptor list, cached by content.  `host` must have been zero-filled
// before its members were assigned (struct padding takes part in the memcmp key).  The cache is
// bounded: past kDescCacheMax entries the oldest is released (remeshes change every list), and
// ab200_set_grid / ab200_unbind clear it.  grid.y of the descriptor kernels is the list length:
// lists longer than 65535 are rejected here instead of failing at launch.
constexpr size_t kDescCacheMax = 64;
inline int cached_descriptors(ab200_ctx *c, const void *host, size_t bytes, int n, void **dev_out) {
  AB_REQUIRE(n > 0 && n <= 65535, AB200_EINVAL,
             "descriptor list: 1..65535 entries per call (split longer lists)");
  for (auto &e : c->halo_cache)
    if (e.bytes == bytes && memcmp(e.host.data(), host, bytes) == 0) {
      *dev_out = e.dev;
      return AB200_OK;
    }
  if (c->halo_cache.size() >= kDescCacheMax) {
    AB_CUDA(cudaStreamSynchronize(c->stream));  // a kernel may still read the oldest copy
    if (c->halo_stream_set) AB_CUDA(cudaStreamSynchronize(c->halo_stream));
    cudaFree(c->halo_cache.front().dev);
    c->halo_cache.erase(c->halo_cache.begin());
  }
  void *d = nullptr;
  AB_CUDA(cudaMalloc(&d, bytes));
  AB_CUDA(cudaMemcpyAsync(d, host, bytes, cudaMemcpyHostToDevice, c->stream));
  AB_CUDA(cudaStreamSynchronize(c->stream));  // the host list goes out of scope
  ab200_ctx::HaloCacheEntry e;
  e.bytes = bytes;
  e.host.assign((const unsigned char *)host, (const unsigned char *)host + bytes);
  e.dev = d;
  c->halo_cache.push_back(std::move(e));
  *dev_out = d;
  return AB200_OK;
}
inline void clear_descriptor_cache(ab200_ctx *c) {
  for (auto &e : c->halo_cache) cudaFree(e.dev);
  c->halo_cache.clear();
}
// kernel launch entry points implemented in tasks_*.cu / fused.cu / halo.cu
int launch_calculate_fluxes(ab200_ctx *c, int fluid, int pcm);
int launch_apply_update(ab200_ctx *c, int fluid, double gam0, double gam1, double beta_dt);
int launch_flux_source(ab200_ctx *c, int fluid, double dt);
int launch_set_aux(ab200_ctx *c);
int launch_cons_to_prim(ab200_ctx *c, int fluid);
int launch_prim_to_cons(ab200_ctx *c, int fluid, int ghosts_only);
int launch_deep_copy(ab200_ctx *c, int fluid);
int launch_estimate_dt(ab200_ctx *c, int fluid, double *d_out, int combine);
int launch_fused_stage(ab200_ctx *c, int fluid, double gam0, double gam1, double beta,
                       double dt, int pcm, int stage1_copy, int use_device_dt,
                       unsigned long long *dt_min, int defer_c2p = 0, int subset = 0,
                       int tap = 0);
bool fused_supports_subsets(const ab200_ctx *c);
bool fused_folds_dt(const ab200_ctx *c);
// single-pass stage (sweep.cuh / sweep_host.cu)
bool sweep_eligible(ab200_ctx *c, int fluid);
bool sweep_uses_role_split(const ab200_ctx *c, int fluid);
int launch_sweep_stage(ab200_ctx *c, int fluid, double gam0, double gam1, double beta, double dt,
                       int pcm, int stage1_copy, int use_device_dt, unsigned long long *dt_min);
// bring the current primitives back into the caller's arrays (no-op when already there)
int sync_prim_home(ab200_ctx *c, int fluid, int interior_only);
int launch_finish_dt(ab200_ctx *c, const double *partial, int n, double cfl, double *d_out,
                     int combine);
int launch_exchange(ab200_ctx *c, int fluid);
int launch_physical_bcs(ab200_ctx *c, int fluid);
int launch_fill_ghosts(ab200_ctx *c, int fluid, int remote_pass);
bool topology_is_local(const ab200_ctx *c);
int launch_finish_stage(ab200_ctx *c, int fluid, unsigned long long *dt_min);
// one stage of the device-resident drivers: fused stage (+ sources + finish when configured)
int run_stage(ab200_ctx *c, double g0, double g1, double beta, int pcm, int first, int last);
int launch_halo(ab200_ctx *c, const ab200_bnd_desc *bnd, int n, int unpack);
int launch_set_global_dt(ab200_ctx *c, double tlim, int advance_time);
int ensure_scratch(ab200_ctx *c, int fluid, bool need_flux, bool need_u1);
// refine.cu: geometry (GridDev gc + metric tables) of the blocks' coarse buffers, built once
int ensure_coarse_grid(ab200_ctx *c);
// diffusion.cu
int launch_diffusion_flux(ab200_ctx *c);
int launch_diffusion_update(ab200_ctx *c, double dt, const double *dt_dev, double beta);
int launch_diffusion_dt(ab200_ctx *c, double *d_out, int combine);
// density-flux tap tables of the fused passes (FluidDev::dflux), allocated on first use
int ensure_dflux(ab200_ctx *c, int fluid);
int build_geom_tables_for(ab200_ctx *c, GeomTab &t, int geom, int nb, int ni, int nj, int nk,
                          const double *xmin_all, const double *dx_all);
}  // namespace ab200
