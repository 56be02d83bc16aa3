// host_path.cu -- ab200_cycles_host: the entry point for callers whose state lives in HOST
// memory (a Kokkos-OpenMP Parthenon build; bench.py's end-to-end leg).  Uploads primitives,
// runs full integrator cycles on the device with the fused path, downloads the result.
#include <algorithm>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "ab200_ctx.cuh"

using namespace ab200;

namespace {
// (gam0, gam1, beta) per stage: P:time_integration/low_storage_integrator.cpp:50-130
struct Stage { double g0, g1, b; };
const Stage kRK1[] = {{0.0, 1.0, 1.0}};
const Stage kRK2[] = {{0.0, 1.0, 1.0}, {0.5, 0.5, 0.5}};
const Stage kVL2[] = {{0.0, 1.0, 0.5}, {0.0, 1.0, 1.0}};
const Stage kRK3[] = {{0.0, 1.0, 1.0}, {0.25, 0.75, 0.25}, {2.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0}};

// ---- interior-only transfers (ab200_set_host_transfer) ---------------------------------------
// Zero-copy variant: one warp moves one interior row [is..ie] of one (block, entry) array
// between the pinned host array (addressed in place over PCIe) and the device array.  Rows are
// 512 B at BASELINE's 64^3 MeshBlocks; every lane keeps kRowsPerWarp independent 8-byte
// accesses in flight so the PCIe round trip is covered by memory-level parallelism, not by
// occupancy.  `skip0..skip1` = pack entries that do not travel (gas pressure on the way in).
constexpr int kXferThreads = 256;
constexpr int kRowsPerWarp = 4;
template <bool TO_DEVICE, typename T>
__global__ void __launch_bounds__(kXferThreads)
k_host_rows(GridDev g, double *const *__restrict__ tab, int nvar, int e0, double *host,
            int skip0, int skip1) {
  constexpr int W = sizeof(T) / sizeof(double);  // doubles per access (1, or 2 when rows are 16-byte aligned)
  const int nir = (g.ie - g.is + 1) / W, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  const int e = blockIdx.y;  // (block, entry) relative to the launch's first entry e0
  const int n = (e0 + e) % nvar;
  if (n >= skip0 && n < skip1) return;
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  T *dev = reinterpret_cast<T *>(tab[e]);
  T *hst = reinterpret_cast<T *>(host + (size_t)e * cells);
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarps = (gridDim.x * blockDim.x) >> 5;
  const int nrows = njr * nkr;
  const int pitch = g.ni / W, i_s = g.is / W;
  for (int r0 = warp * kRowsPerWarp; r0 < nrows; r0 += nwarps * kRowsPerWarp) {
    for (int i0 = lane; i0 < nir; i0 += 32) {
      T v[kRowsPerWarp];
#pragma unroll
      for (int q = 0; q < kRowsPerWarp; ++q) {
        const int r = r0 + q;
        if (r < nrows) {
          const size_t off = ((size_t)(g.ks + r / njr) * g.nj + (g.js + r % njr)) * pitch + i_s + i0;
          v[q] = TO_DEVICE ? __ldcs(hst + off) : __ldcs(dev + off);
        }
      }
#pragma unroll
      for (int q = 0; q < kRowsPerWarp; ++q) {
        const int r = r0 + q;
        if (r < nrows) {
          const size_t off = ((size_t)(g.ks + r / njr) * g.nj + (g.js + r % njr)) * pitch + i_s + i0;
          if (TO_DEVICE) dev[off] = v[q];
          else __stcs(hst + off, v[q]);
        }
      }
    }
  }
}

// device-visible alias of a pinned host array (UVA: normally the same address)
int host_alias(double *h, double **out) {
  cudaPointerAttributes at;
  std::memset(&at, 0, sizeof at);
  cudaError_t e = cudaPointerGetAttributes(&at, h);
  if (e != cudaSuccess) (void)cudaGetLastError();
  AB_REQUIRE(e == cudaSuccess && at.type == cudaMemoryTypeHost && at.devicePointer, AB200_EINVAL,
             "ab200_cycles_host: AB200_HOST_ZERO_COPY needs pinned host arrays "
             "(cudaHostAlloc / cudaHostRegister)");
  *out = (double *)at.devicePointer;
  return AB200_OK;
}

// interior zones of every (block, entry) array of one fluid, host <-> device
int interior_transfer(ab200_ctx *c, int fl, double *host, bool to_device,
                      const std::vector<double *> &tab) {
  const GridDev &g = c->g;
  const FluidDev &f = c->fl[fl].d;
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  const int skip0 = (to_device && fl == AB200_GAS) ? 4 * f.S : 0;
  const int skip1 = (to_device && fl == AB200_GAS) ? 5 * f.S : 0;
  const int nir = g.ie - g.is + 1, njr = g.je - g.js + 1, nkr = g.ke - g.ks + 1;
  if (c->host_transfer & (to_device ? AB200_HOST_ZERO_COPY_IN : AB200_HOST_ZERO_COPY_OUT)) {
    double *alias = nullptr;
    AB_TRY(host_alias(host, &alias));
    const int nrows = njr * nkr;
    const int warps = (nrows + kRowsPerWarp - 1) / kRowsPerWarp;
    const unsigned gx = (unsigned)std::max(1, std::min((warps * 32 + kXferThreads - 1) / kXferThreads, 64));
    const size_t nent = tab.size();
    // 16-byte accesses when every interior row starts and ends on a 16-byte boundary
    bool wide = (nir % 2 == 0) && (g.is % 2 == 0) && (g.ni % 2 == 0) && (cells % 2 == 0) &&
                ((uintptr_t)alias % 16 == 0) && !getenv("AB200_XFER_SCALAR");
    for (size_t e = 0; e < tab.size() && wide; ++e) wide = ((uintptr_t)tab[e] % 16 == 0);
    for (size_t e0 = 0; e0 < nent; e0 += 65535) {  // grid.y limit
      const unsigned ny = (unsigned)std::min<size_t>(65535, nent - e0);
      dim3 grid(gx, ny);
#define AB_XFER(TD, T)                                                                            \
  k_host_rows<TD, T><<<grid, kXferThreads, 0, c->stream>>>(g, f.prim + e0, f.nvar, (int)(e0 % f.nvar), \
                                                          alias + e0 * cells, skip0, skip1)
      if (to_device) { if (wide) AB_XFER(true, double2); else AB_XFER(true, double); }
      else { if (wide) AB_XFER(false, double2); else AB_XFER(false, double); }
#undef AB_XFER
      c->launches++;
    }
    AB_CUDA(cudaGetLastError());
    return AB200_OK;
  }
  // strided DMA: one 3-D copy per array (rows of nir doubles, pitch ni)
  for (size_t e = 0; e < tab.size(); ++e) {
    const int n = (int)(e % f.nvar);
    if (n >= skip0 && n < skip1) continue;
    cudaMemcpy3DParms p;
    std::memset(&p, 0, sizeof p);
    double *h = host + e * cells;
    const cudaPitchedPtr hp = make_cudaPitchedPtr(h, (size_t)g.ni * sizeof(double), g.ni, g.nj);
    const cudaPitchedPtr dp = make_cudaPitchedPtr(tab[e], (size_t)g.ni * sizeof(double), g.ni, g.nj);
    p.srcPtr = to_device ? hp : dp;
    p.dstPtr = to_device ? dp : hp;
    p.srcPos = p.dstPos = make_cudaPos((size_t)g.is * sizeof(double), g.js, g.ks);
    p.extent = make_cudaExtent((size_t)nir * sizeof(double), njr, nkr);
    p.kind = to_device ? cudaMemcpyHostToDevice : cudaMemcpyDeviceToHost;
    AB_CUDA(cudaMemcpy3DAsync(&p, c->stream));
  }
  return AB200_OK;
}
}  // namespace

extern "C" int ab200_set_graph_replay(ab200_ctx *c, int on) {
  AB_REQUIRE(c, AB200_EINVAL, "ab200_set_graph_replay: null context");
  c->graph_replay = on != 0;
  return AB200_OK;
}

extern "C" int ab200_graph_replay_count(ab200_ctx *c, long long *count) {
  AB_REQUIRE(c && count, AB200_EINVAL, "ab200_graph_replay_count: null argument");
  *count = c->graph_replays;
  return AB200_OK;
}

extern "C" int ab200_set_host_transfer(ab200_ctx *c, int flags) {
  AB_REQUIRE(c, AB200_EINVAL, "ab200_set_host_transfer: null context");
  AB_REQUIRE((flags & ~(AB200_HOST_INTERIOR_IN | AB200_HOST_INTERIOR_OUT | AB200_HOST_ZERO_COPY)) == 0,
             AB200_EINVAL, "ab200_set_host_transfer: unknown flag");
  c->host_transfer = flags;
  return AB200_OK;
}

extern "C" int ab200_run_cycles(ab200_ctx *c, int integrator, int ncycles, double tlim) {
  AB_REQUIRE(c && c->grid_set, AB200_ESTATE, "ab200_run_cycles: no grid bound");
  AB_REQUIRE(c->topo.set, AB200_ESTATE, "ab200_run_cycles: call ab200_set_topology first");
  AB_REQUIRE(integrator >= 0 && integrator <= 3, AB200_EINVAL, "unknown integrator");
  // single-rank entry point: faces owned by another rank (AB200_BC_NONE) need the remote
  // exchange between the stages, which only the caller's transport can do
  AB_REQUIRE(topology_is_local(c), AB200_ESTATE,
             "ab200_run_cycles: topology has AB200_BC_NONE (remote) faces; drive the stages "
             "with ab200_fused_stage + the remote ghost exchange instead");
  AB_CUDA(cudaSetDevice(c->device));
  // a finite tlim is honoured like EvolutionDriver::Execute does (P:driver/driver.cpp:99):
  // the loop stops once time >= tlim, which costs one 32-byte read-back per cycle; with
  // tlim = DBL_MAX the cycles are queued without any host round trip
  const bool finite_tlim = tlim < 1.0e300;
  // the fused stages never read conserved ghost zones: inside the loop the ghost fills write
  // primitives only; PrimToCons on the ghosts (ghost part of fill_derived.cpp:217-274) runs once
  // after the last cycle so the caller's arrays are complete again
  const bool lazy_before = c->ghost_cons_lazy;
  c->ghost_cons_lazy = true;
  struct Restore {
    ab200_ctx *c; bool v;
    ~Restore() { c->ghost_cons_lazy = v; }
  } restore{c, lazy_before};
  const Stage *st = integrator == 0 ? kRK1 : integrator == 1 ? kRK2 : integrator == 2 ? kVL2 : kRK3;
  const int nst = integrator == 0 ? 1 : integrator == 3 ? 3 : 2;
  auto one_cycle = [&]() -> int {
    for (int s = 0; s < nst; ++s) {
      const int pcm = (s == 0 && integrator == 2);  // vl2 stage 1: artemis_driver.cpp:182
      AB_TRY(run_stage(c, st[s].g0, st[s].g1, st[s].b, pcm, s == 0, s == nst - 1));
      AB_TRY(ab200_fill_ghosts(c));
    }
    return ab200_set_global_timestep_device(c, tlim, 1);
  };
  int cyc = 0;
  // CUDA-graph replay (opt-in, ab200_set_graph_replay).  With no finite tlim the cycle is a
  // fixed launch sequence whose every input (dt, time, the CFL slots) lives on the device, so
  // cycle 2 is captured once and replayed: one cudaGraphLaunch per cycle instead of 8-20 kernel
  // launches.  Measured: 6-8 % on launch-bound meshes (config 1's 8 MeshBlocks of 32^2: 137 ->
  // 127 us per cycle), nothing at 256^3 -- the eager loop is already asynchronous.  The
  // first cycle runs eagerly so that first-use allocations (tensor maps, descriptor caches) are
  // not captured.  Host-side state must be periodic over one cycle for the captured pointers to
  // stay valid: the single-pass kernels alternate primitive sets per STAGE, so an odd stage
  // count (rk1 / rk3) is replayed only when no fluid runs them.
  bool periodic = (nst % 2 == 0);
  if (!periodic) {
    periodic = true;
    for (int f = 0; f < 2; ++f)
      if (c->fl[f].bound && sweep_eligible(c, f)) periodic = false;
  }
  static const bool env_on = getenv("AB200_GRAPH") != nullptr;
  if ((c->graph_replay || env_on) && !finite_tlim && periodic && ncycles >= 3) {
    AB_TRY(one_cycle());
    cyc = 1;
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    const long long l0 = c->launches;
    // the legacy default stream cannot be captured: record on a private stream (nothing runs
    // during a capture, so ordering against the real stream is moot) and replay on the real one
    cudaStream_t real = c->stream, cap = c->stream;
    const bool own = (real == nullptr || real == cudaStreamLegacy || real == cudaStreamPerThread);
    if (own && cudaStreamCreateWithFlags(&cap, cudaStreamNonBlocking) != cudaSuccess) cap = nullptr;
    if (cap != nullptr &&
        cudaStreamBeginCapture(cap, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      c->stream = cap;
      const int rc_cap = one_cycle();
      c->stream = real;
      const cudaError_t e_end = cudaStreamEndCapture(cap, &graph);
      if (own) cudaStreamDestroy(cap);
      const long long per_cycle = c->launches - l0;
      c->launches = l0;  // nothing ran during the capture
      if (rc_cap != AB200_OK) {
        if (graph) cudaGraphDestroy(graph);
        (void)cudaGetLastError();
        return rc_cap;  // the message of the failing entry point stands
      }
      bool ok = (e_end == cudaSuccess) && graph &&
                cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess;
      for (; ok && cyc < ncycles; ++cyc) {
        ok = cudaGraphLaunch(exec, c->stream) == cudaSuccess;
        if (ok) {
          c->launches += per_cycle;
          c->graph_replays++;
        }
      }
      if (exec) cudaGraphExecDestroy(exec);
      if (graph) cudaGraphDestroy(graph);
      if (!ok) {
        (void)cudaGetLastError();
        if (cyc > 1) {  // a replay failed after some had run: do not guess, report
          set_error("ab200_run_cycles: cudaGraphLaunch failed");
          return AB200_ECUDA;
        }
      }
    } else {
      if (own && cap) cudaStreamDestroy(cap);
      (void)cudaGetLastError();
    }
  }
  for (; cyc < ncycles; ++cyc) {
    AB_TRY(one_cycle());
    if (finite_tlim) {
      double ts[4];
      AB_TRY(ab200_read_time_state(c, ts));
      if (ts[2] >= tlim) break;
    }
  }
  // odd number of single-pass stages in total (rk1 / rk3): primitives back to the caller
  AB_TRY(ab200_sync_prim(c));
  if (!lazy_before) AB_TRY(ab200_sync_ghost_cons(c));
  return AB200_OK;
}

extern "C" int ab200_cycles_host(ab200_ctx *c, int integrator, int ncycles, double *dt_io,
                                 double *gas_prim_host, double *gas_cons_host,
                                 double *dust_prim_host, double *dust_cons_host) {
  AB_REQUIRE(c && c->grid_set, AB200_ESTATE, "ab200_cycles_host: no grid bound");
  AB_REQUIRE(c->topo.set, AB200_ESTATE, "ab200_cycles_host: call ab200_set_topology first");
  // faces owned by another rank (AB200_BC_NONE): every rank calls this collectively and the
  // cycles run through ab200_run_cycles_mr (needs ab200_comm_init + ab200_comm_set_layout)
  const bool local = topology_is_local(c);
  AB_REQUIRE(dt_io, AB200_EINVAL, "ab200_cycles_host: null dt");
  AB_CUDA(cudaSetDevice(c->device));
  const GridDev &g = c->g;
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  double *hp[2] = {gas_prim_host, dust_prim_host};
  double *hc[2] = {gas_cons_host, dust_cons_host};
  // everything that can be refused is refused before the first byte moves
  AB_REQUIRE(!(c->host_transfer & AB200_HOST_INTERIOR_OUT) || (!hc[0] && !hc[1]), AB200_EINVAL,
             "ab200_cycles_host: AB200_HOST_INTERIOR_OUT returns primitives only "
             "(pass NULL for the conserved arrays)");
  for (int fl = 0; fl < 2; ++fl) {
    if (!c->fl[fl].bound) continue;
    AB_REQUIRE(hp[fl], AB200_EINVAL, "ab200_cycles_host: null host array for a bound fluid");
    const int ht = c->host_transfer;
    if (((ht & AB200_HOST_INTERIOR_IN) && (ht & AB200_HOST_ZERO_COPY_IN)) ||
        ((ht & AB200_HOST_INTERIOR_OUT) && (ht & AB200_HOST_ZERO_COPY_OUT))) {
      double *alias = nullptr;
      AB_TRY(host_alias(hp[fl], &alias));
    }
  }
  // Upload the primitives, [nblocks][nvar][nk][nj][ni].  The gas pressure entries are NOT
  // uploaded: PrimToCons recomputes P = EOS(rho, sie) over the entire domain
  // (fill_derived.cpp:247) before anything reads it, so those bytes never need to cross PCIe.
  for (int fl = 0; fl < 2; ++fl) {
    if (!c->fl[fl].bound) continue;
    AB_REQUIRE(hp[fl], AB200_EINVAL, "ab200_cycles_host: null host array for a bound fluid");
    const FluidDev &f = c->fl[fl].d;
    AB_TRY(sync_prim_home(c, fl, 0));
    std::vector<double *> tab((size_t)g.nb * f.nvar);
    AB_CUDA(cudaMemcpyAsync(tab.data(), f.prim, tab.size() * sizeof(double *),
                            cudaMemcpyDeviceToHost, c->stream));
    AB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->host_transfer & AB200_HOST_INTERIOR_IN) {
      AB_TRY(interior_transfer(c, fl, hp[fl], true, tab));
      continue;
    }
    for (size_t e = 0; e < tab.size(); ++e) {
      const int n = (int)(e % f.nvar);
      if (fl == AB200_GAS && n >= 4 * f.S && n < 5 * f.S) continue;  // pressure: derived
      AB_CUDA(cudaMemcpyAsync(tab[e], hp[fl] + e * cells, cells * sizeof(double),
                              cudaMemcpyHostToDevice, c->stream));
    }
  }
  if (c->host_transfer & AB200_HOST_INTERIOR_IN) {
    // ghost primitives from the just-uploaded interior zones: same-GPU exchange + physical
    // boundaries (what the end of the previous stage did on the caller's side); the pressure
    // entries they copy are stale, PrimToCons below recomputes P over the entire domain
    const bool lazy = c->ghost_cons_lazy;
    c->ghost_cons_lazy = true;
    int rc_fill;
    if (local) {
      rc_fill = ab200_fill_ghosts(c);
    } else {  // the same fill with the remote round between its two halves
      rc_fill = ab200_comm_exchange_begin(c);
      if (rc_fill == AB200_OK) rc_fill = ab200_fill_ghosts_local(c);
      if (rc_fill == AB200_OK) rc_fill = ab200_comm_exchange_end(c);
      if (rc_fill == AB200_OK) rc_fill = ab200_finish_remote_ghosts(c);
    }
    c->ghost_cons_lazy = lazy;
    AB_TRY(rc_fill);
    for (int fl = 0; fl < 2; ++fl) c->fl[fl].ghost_cons_stale = false;  // full PrimToCons follows
  }
  AB_TRY(ab200_prim_to_cons(c));  // cons == PrimToCons(prim) at every cycle boundary
  double ts[4] = {0, 0, 0, 0};
  if (*dt_io > 0.0) {
    ts[0] = *dt_io;
    AB_TRY(ab200_write_time_state(c, ts));
  } else {
    ts[0] = 1.79769313486231570815e+308;
    AB_TRY(ab200_write_time_state(c, ts));
    AB_TRY(ab200_estimate_timestep_device(c));
    if (!local) AB_TRY(ab200_allreduce_min(c, c->d_time + 1));
    AB_TRY(ab200_set_global_timestep_device(c, 1.79769313486231570815e+308, 0));
  }
  // Conserved arrays are optional on the way back (cons is a pure function of prim: the host
  // can ask for primitives only and halve the download); when they are not wanted the ghost
  // PrimToCons is skipped as well.
  const bool want_cons = (hc[0] != nullptr) || (hc[1] != nullptr);
  const bool lazy_before = c->ghost_cons_lazy;
  if (!want_cons) c->ghost_cons_lazy = true;
  const int rc_run = local ? ab200_run_cycles(c, integrator, ncycles, 1.79769313486231570815e+308)
                           : ab200_run_cycles_mr(c, integrator, ncycles, 1.79769313486231570815e+308);
  c->ghost_cons_lazy = lazy_before;
  AB_TRY(rc_run);
  for (int fl = 0; fl < 2; ++fl) {
    if (!c->fl[fl].bound) continue;
    const FluidDev &f = c->fl[fl].d;
    std::vector<double *> tp((size_t)g.nb * f.nvar), tc((size_t)g.nb * f.nvar);
    AB_CUDA(cudaMemcpyAsync(tp.data(), f.prim, tp.size() * sizeof(double *),
                            cudaMemcpyDeviceToHost, c->stream));
    AB_CUDA(cudaMemcpyAsync(tc.data(), f.u0, tc.size() * sizeof(double *),
                            cudaMemcpyDeviceToHost, c->stream));
    AB_CUDA(cudaStreamSynchronize(c->stream));
    if (c->host_transfer & AB200_HOST_INTERIOR_OUT) {
      AB_TRY(interior_transfer(c, fl, hp[fl], false, tp));
      continue;
    }
    for (size_t e = 0; e < tp.size(); ++e) {
      AB_CUDA(cudaMemcpyAsync(hp[fl] + e * cells, tp[e], cells * sizeof(double),
                              cudaMemcpyDeviceToHost, c->stream));
      if (hc[fl])
        AB_CUDA(cudaMemcpyAsync(hc[fl] + e * cells, tc[e], cells * sizeof(double),
                                cudaMemcpyDeviceToHost, c->stream));
    }
  }
  AB_TRY(ab200_read_time_state(c, ts));
  *dt_io = ts[0];
  return AB200_OK;
}
