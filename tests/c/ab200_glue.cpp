// ab200_glue.cpp -- the ONE translation unit a maintainer adds to lanl/artemis
// (as src/utils/ab200_glue.cpp) to route the hot path through libartemis_b200.  It uses public
// Parthenon accessors only -- MeshData::GetBlockData(b) (P:interface/mesh_data.hpp:282),
// MeshBlockData::Get(label) (P:interface/meshblock_data.hpp:262), Variable::data
// (P:interface/variable.hpp:138), ParArrayND::data() (P:parthenon_array_generic.hpp:209),
// MeshBlock::cellbounds / coords (P:mesh/meshblock.hpp:122) -- so no Parthenon patch is needed.
// (SparsePack's host mirror `pack_h_` is protected, P:interface/sparse_pack_base.hpp:93; its
// device accessor operator()(b, idx) P:interface/sparse_pack.hpp:303-307 is what the library's
// pointer tables replace.)
//
// Compiled on every test run against tests/c/parthenon_glue_mock.hpp (tests/test_c_abi.py);
// inside Artemis the first include is "artemis.hpp" instead.
#ifdef AB200_GLUE_SYNTAX_CHECK
#include "parthenon_glue_mock.hpp"
#else
#include "artemis.hpp"
#endif
#include "ab200.h"

namespace AB200Glue {
using parthenon::IndexDomain;
using parthenon::MeshData;
using parthenon::TaskStatus;

struct Binding {  // one per MeshData partition
  ab200_ctx *ctx = nullptr;
  std::vector<double *> prim[2], u0[2], u1[2];  // [fluid][block * nvar + pack index]
  std::vector<double> xmin, dx;                 // [block][3]
  int nblocks = -1;
  const void *first_array = nullptr;  // re-bind when Parthenon re-allocates (remesh)
};

inline void Check(int rc) {
  if (rc != AB200_OK) PARTHENON_FAIL(ab200_last_error());
}
inline TaskStatus Status(int rc) { return rc == AB200_OK ? TaskStatus::complete : TaskStatus::fail; }

// Components of one field of block b appended to a pointer table: a Variable's data is
// [ncomp][nk][nj][ni] (P:interface/variable.cpp:112-128), component c at data() + c*nk*nj*ni.
inline void Append(std::vector<double *> &tab, const parthenon::Variable<Real> &v, size_t cells) {
  const int ncomp = v.data.GetDim(4);
  for (int c = 0; c < ncomp; ++c) tab.push_back(v.data.data() + (size_t)c * cells);
}

// Pack order of ab200_pack_desc == the order CalculateFluxes builds its packs in
// (src/gas/gas.cpp:473-494: density, velocity, pressure, sie; hllc.hpp:66-73 IDN, IVX, IPR, ISE)
Binding &Bind(MeshData<Real> *u0, MeshData<Real> *u1, int cuda_device, void *cuda_stream) {
  static std::map<MeshData<Real> *, Binding> cache;
  Binding &B = cache[u0];
  auto *pm = u0->GetParentPointer();
  const int nb = u0->NumBlocks();
  const void *first = u0->GetBlockData(0)->Get("gas.prim.density").data.data();
  if (B.ctx && B.nblocks == nb && B.first_array == first) return B;  // still valid
  if (!B.ctx) Check(ab200_create(&B.ctx, cuda_device, cuda_stream));
  auto *pmb0 = u0->GetBlockData(0)->GetBlockPointer();
  const auto &cb = pmb0->cellbounds;
  const size_t cells = (size_t)cb.ncellsi(IndexDomain::entire) * cb.ncellsj(IndexDomain::entire) *
                       cb.ncellsk(IndexDomain::entire);
  const bool do_dust = false;  // = artemis_pkg->Param<bool>("do_dust"), src/artemis.cpp:60-70
  for (int f = 0; f < 2; ++f) { B.prim[f].clear(); B.u0[f].clear(); B.u1[f].clear(); }
  B.xmin.assign(3 * (size_t)nb, 0.0);
  B.dx.assign(3 * (size_t)nb, 0.0);
  for (int b = 0; b < nb; ++b) {
    auto &d0 = *u0->GetBlockData(b);
    auto &d1 = *u1->GetBlockData(b);
    for (const char *name : {"gas.prim.density", "gas.prim.velocity", "gas.prim.pressure", "gas.prim.sie"})
      Append(B.prim[0], d0.Get(name), cells);
    for (const char *name : {"gas.cons.density", "gas.cons.momentum", "gas.cons.total_energy",
                             "gas.cons.internal_energy"}) {
      Append(B.u0[0], d0.Get(name), cells);
      Append(B.u1[0], d1.Get(name), cells);
    }
    if (do_dust) {
      for (const char *name : {"dust.prim.density", "dust.prim.velocity"}) Append(B.prim[1], d0.Get(name), cells);
      for (const char *name : {"dust.cons.density", "dust.cons.momentum"}) {
        Append(B.u0[1], d0.Get(name), cells);
        Append(B.u1[1], d1.Get(name), cells);
      }
    }
    const auto &co = d0.GetBlockPointer()->coords;  // UniformCartesian: Xf(idx) = xmin_ + idx*dx_
    B.xmin[3 * b + 0] = co.Xf<1, 1>(0); B.dx[3 * b + 0] = co.Dxf<1>();
    B.xmin[3 * b + 1] = co.Xf<2, 2>(0); B.dx[3 * b + 1] = co.Dxf<2>();
    B.xmin[3 * b + 2] = co.Xf<3, 3>(0); B.dx[3 * b + 2] = co.Dxf<3>();
  }
  auto gas_pkg = pm->packages.Get("gas");
  ab200_grid_desc g{};
  g.geom = static_cast<int>(gas_pkg->Param<Coordinates>("coords"));  // same enum order, artemis.hpp:78
  g.ndim = pm->ndim;
  g.nghost = parthenon::Globals::nghost;
  g.nblocks = nb;
  g.ni = cb.ncellsi(IndexDomain::entire); g.nj = cb.ncellsj(IndexDomain::entire);
  g.nk = cb.ncellsk(IndexDomain::entire);
  g.is = cb.is(IndexDomain::interior); g.ie = cb.ie(IndexDomain::interior);
  g.js = cb.js(IndexDomain::interior); g.je = cb.je(IndexDomain::interior);
  g.ks = cb.ks(IndexDomain::interior); g.ke = cb.ke(IndexDomain::interior);
  g.fni = g.ni + 1; g.fnj = g.nj + (pm->ndim > 1); g.fnk = g.nk + (pm->ndim > 2);
  g.xmin = B.xmin.data(); g.dx = B.dx.data();
  Check(ab200_set_grid(B.ctx, &g));
  ab200_fluid_desc fd{};
  fd.fluid = AB200_GAS;
  fd.nspecies = gas_pkg->Param<int>("nspecies");                                 // gas.cpp:201
  fd.recon = static_cast<int>(gas_pkg->Param<ReconstructionMethod>("recon"));    // gas.cpp:81
  fd.riemann = static_cast<int>(gas_pkg->Param<RSolver>("rsolver"));             // gas.cpp:95
  fd.gm1 = gas_pkg->Param<Real>("adiabatic_index") - 1.0;                        // gas.cpp:121
  fd.dfloor = gas_pkg->Param<Real>("dfloor");                                    // gas.cpp:171
  fd.siefloor = gas_pkg->Param<Real>("siefloor");                                // gas.cpp:172
  fd.de_switch = gas_pkg->Param<Real>("de_switch");                              // gas.cpp:178
  fd.cfl = gas_pkg->Param<Real>("cfl");                                          // gas.cpp:99
  ab200_pack_desc pk{};
  pk.prim = B.prim[0].data(); pk.cons0 = B.u0[0].data(); pk.cons1 = B.u1[0].data();
  // flux / pflux / vface tables: fill the same way from "gas.cons.*" flux fields when the
  // per-task entry points are used; NULL is fine for the fused path
  Check(ab200_bind_pack(B.ctx, &fd, &pk));
  B.nblocks = nb;
  B.first_array = first;
  return B;
}

// ---- task bodies: signatures of the reference, bodies are one C-ABI call each ---------------
// Gas::CalculateFluxes, src/gas/gas.cpp:473
TaskStatus CalculateFluxes(MeshData<Real> *md, const bool pcm, int dev, void *stream) {
  return Status(ab200_calculate_fluxes(Bind(md, md, dev, stream).ctx, AB200_GAS, pcm));
}
// Gas::FluxSource, src/gas/gas.cpp:499
TaskStatus FluxSource(MeshData<Real> *md, const Real dt, int dev, void *stream) {
  return Status(ab200_flux_source(Bind(md, md, dev, stream).ctx, AB200_GAS, dt));
}
// ArtemisUtils::ApplyUpdate<GEOM>, src/utils/integrators/artemis_integrator.hpp:56
TaskStatus ApplyUpdate(MeshData<Real> *u0, MeshData<Real> *u1, const int stage,
                       parthenon::LowStorageIntegrator *integrator, int dev, void *stream) {
  return Status(ab200_apply_update(Bind(u0, u1, dev, stream).ctx, integrator->gam0[stage - 1],
                                   integrator->gam1[stage - 1],
                                   integrator->beta[stage - 1] * integrator->dt));
}
// the fused stage that replaces the chain src/artemis_driver.cpp:184-255 when no out-of-scope
// source term sits in between
TaskStatus FusedStage(MeshData<Real> *u0, MeshData<Real> *u1, const int stage,
                      parthenon::LowStorageIntegrator *integrator, const bool do_pcm, int dev,
                      void *stream) {
  return Status(ab200_fused_stage(Bind(u0, u1, dev, stream).ctx, integrator->gam0[stage - 1],
                                  integrator->gam1[stage - 1], integrator->beta[stage - 1],
                                  integrator->dt, do_pcm, stage == 1, 0));
}
// Gas::EstimateTimestepMesh<GEOM>, src/gas/gas.cpp:391
Real EstimateTimestepMesh(MeshData<Real> *md, int dev, void *stream) {
  Real dt = 0.0;
  Check(ab200_estimate_timestep(Bind(md, md, dev, stream).ctx, AB200_GAS, &dt));
  return dt;
}

// ---- user boundary conditions of the strat / ssheet generators ---------------------------------
// Registered in place of strat::ExtrapInnerX1 ... strat::ExtrapOuterX3
// (src/pgen/problem_modifier.hpp:114-127), e.g.
//   pman->app_input->RegisterBoundaryCondition(BF::inner_x1, "extrap",
//                                              AB200Glue::UserBc<0, AB200_BC_EXTRAP>);
// Parthenon calls the functions block by block and face by face
// (ApplyBoundaryConditionsOnCoarseOrFine, P:bvals/boundary_conditions.cpp); each call only APPENDS
// a descriptor, FlushUserBcs hands the list to the library, which applies the x1 faces of the
// whole list first, then x2, then x3 -- Parthenon's order within every block.
struct UserBcQueue {
  ab200_ctx *ctx = nullptr;
  int nspecies = 1, dust_species = 0;  // gas / dust nspecies of the bound packs (0: no dust)
  std::map<const parthenon::MeshBlock *, int> block_of;  // position of a block in the bound pack
  std::vector<ab200_block_bc_desc> pending;
  // per pending descriptor: the block's per-entry coarse arrays [6 S] (empty = fine arrays);
  // FlushUserBcs uploads them as DEVICE tables (ab200_malloc + ab200_memcpy_h2d)
  std::vector<std::vector<double *>> coarse_entries_host;
  std::vector<void *> device_tables;
};
inline UserBcQueue &Queue() { static UserBcQueue q; return q; }

template <int FACE, int TYPE>
void UserBc(std::shared_ptr<parthenon::MeshBlockData<Real>> &mbd, bool coarse) {
  UserBcQueue &Q = Queue();
  // the reference's functions set the gas and, when do_dust, every dust species in one call
  // (strat.hpp:183-218); here that is one descriptor per bound fluid
  const char *const gas_fields[] = {"gas.prim.density", "gas.prim.velocity", "gas.prim.pressure",
                                    "gas.prim.sie"};
  const char *const dust_fields[] = {"dust.prim.density", "dust.prim.velocity"};
  for (int fluid = AB200_GAS; fluid <= (Q.dust_species > 0 ? AB200_DUST : AB200_GAS); ++fluid) {
    ab200_block_bc_desc d{};
    d.fluid = fluid; d.block = Q.block_of[mbd->GetBlockPointer()];
    d.var0 = 0; d.ncomp = fluid == AB200_GAS ? 6 * Q.nspecies : 4 * Q.dust_species;
    d.face = FACE; d.type = TYPE;
    d.coarse = nullptr; d.coarse_entries = nullptr;
    std::vector<double *> ent;
    if (coarse) {
      // one coarse buffer per Variable (P:interface/variable.hpp:139): the pack entries of the
      // fluid's primitives in pack order (hllc.hpp:66-73), component c of a Variable at
      // coarse_s.data() + c * (cells of the coarse index space)
      const char *const *names = fluid == AB200_GAS ? gas_fields : dust_fields;
      const int nnames = fluid == AB200_GAS ? 4 : 2;
      for (int f = 0; f < nnames; ++f) {
        const auto &cs = mbd->Get(names[f]).coarse_s;
        const size_t ccells = (size_t)cs.GetDim(1) * cs.GetDim(2) * cs.GetDim(3);
        for (int c = 0; c < cs.GetDim(4); ++c) ent.push_back(cs.data() + c * ccells);
      }
    }
    Q.coarse_entries_host.push_back(ent);
    Q.pending.push_back(d);
  }
}

// after ApplyBoundaryConditionsOnCoarseOrFineMD has walked the blocks of the partition
inline TaskStatus FlushUserBcs(Real qshear, Real omega) {
  UserBcQueue &Q = Queue();
  if (Q.pending.empty()) return TaskStatus::complete;
  Check(ab200_set_shear_bc_params(Q.ctx, qshear, omega));  // StratParams, strat.hpp:36-66
  for (size_t q = 0; q < Q.pending.size(); ++q) {
    const auto &ent = Q.coarse_entries_host[q];
    if (ent.empty()) continue;
    ab200_block_bc_desc &d = Q.pending[q];
    void *dev = nullptr;
    Check(ab200_malloc(Q.ctx, &dev, ent.size() * sizeof(double *)));
    Check(ab200_memcpy_h2d(Q.ctx, dev, ent.data(), ent.size() * sizeof(double *)));
    Q.device_tables.push_back(dev);
    d.coarse_entries = static_cast<double *const *>(dev);
  }
  const int rc = ab200_block_bcs(Q.ctx, Q.pending.data(), (int)Q.pending.size());
  Check(ab200_synchronize(Q.ctx));
  for (void *dev : Q.device_tables) Check(ab200_free(Q.ctx, dev));
  Q.device_tables.clear();
  Q.coarse_entries_host.clear();
  Q.pending.clear();
  return Status(rc);
}
// the six registrations of the strat problem, as the reference spells them
inline void RegisterStratBoundaries(
    void (*reg)(int face, const char *name,
                void (*fn)(std::shared_ptr<parthenon::MeshBlockData<Real>> &, bool))) {
  reg(0, "extrap", UserBc<0, AB200_BC_EXTRAP>);
  reg(1, "extrap", UserBc<1, AB200_BC_EXTRAP>);
  reg(2, "inflow", UserBc<2, AB200_BC_INFLOW>);
  reg(3, "inflow", UserBc<3, AB200_BC_INFLOW>);
  reg(4, "extrap", UserBc<4, AB200_BC_EXTRAP>);
  reg(5, "extrap", UserBc<5, AB200_BC_EXTRAP>);
}
}  // namespace AB200Glue
