// host_path.cu -- ab200_cycles_host: the entry point for callers whose state lives in HOST
// memory (a Kokkos-OpenMP Parthenon build; bench.py's end-to-end leg).  Uploads primitives,
// runs full integrator cycles on the device with the fused path, downloads the result.
#include <cstring>

#include "ab200_ctx.cuh"

using namespace ab200;

namespace {
// (gam0, gam1, beta) per stage: P:time_integration/low_storage_integrator.cpp:50-130
struct Stage { double g0, g1, b; };
const Stage kRK1[] = {{0.0, 1.0, 1.0}};
const Stage kRK2[] = {{0.0, 1.0, 1.0}, {0.5, 0.5, 0.5}};
const Stage kVL2[] = {{0.0, 1.0, 0.5}, {0.0, 1.0, 1.0}};
const Stage kRK3[] = {{0.0, 1.0, 1.0}, {0.25, 0.75, 0.25}, {2.0 / 3.0, 1.0 / 3.0, 2.0 / 3.0}};
}  // namespace

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
  for (int cyc = 0; cyc < ncycles; ++cyc) {
    for (int s = 0; s < nst; ++s) {
      const int pcm = (s == 0 && integrator == 2);  // vl2 stage 1: artemis_driver.cpp:182
      AB_TRY(run_stage(c, st[s].g0, st[s].g1, st[s].b, pcm, s == 0, s == nst - 1));
      AB_TRY(ab200_fill_ghosts(c));
    }
    AB_TRY(ab200_set_global_timestep_device(c, tlim, 1));
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
  AB_REQUIRE(topology_is_local(c), AB200_ESTATE,
             "ab200_cycles_host: single-rank entry point (topology has AB200_BC_NONE faces)");
  AB_REQUIRE(dt_io, AB200_EINVAL, "ab200_cycles_host: null dt");
  AB_CUDA(cudaSetDevice(c->device));
  const GridDev &g = c->g;
  const size_t cells = (size_t)g.ni * g.nj * g.nk;
  double *hp[2] = {gas_prim_host, dust_prim_host};
  double *hc[2] = {gas_cons_host, dust_cons_host};
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
    for (size_t e = 0; e < tab.size(); ++e) {
      const int n = (int)(e % f.nvar);
      if (fl == AB200_GAS && n >= 4 * f.S && n < 5 * f.S) continue;  // pressure: derived
      AB_CUDA(cudaMemcpyAsync(tab[e], hp[fl] + e * cells, cells * sizeof(double),
                              cudaMemcpyHostToDevice, c->stream));
    }
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
    AB_TRY(ab200_set_global_timestep_device(c, 1.79769313486231570815e+308, 0));
  }
  // Conserved arrays are optional on the way back (cons is a pure function of prim: the host
  // can ask for primitives only and halve the download); when they are not wanted the ghost
  // PrimToCons is skipped as well.
  const bool want_cons = (hc[0] != nullptr) || (hc[1] != nullptr);
  const bool lazy_before = c->ghost_cons_lazy;
  if (!want_cons) c->ghost_cons_lazy = true;
  const int rc_run = ab200_run_cycles(c, integrator, ncycles, 1.79769313486231570815e+308);
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
