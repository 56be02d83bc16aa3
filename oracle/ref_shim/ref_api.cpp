// ref_api.cpp -- C entry points of oracle/_ref/libartemis_ref.so: the reference's OWN hot-path
// sources (compiled unmodified from /root/reference/src against the mock Parthenon of
// include/parthenon_shim.hpp) behind the same C signatures as oracle/artemis_oracle.h, so the
// tests can put reference, restatement and CUDA side by side on the same arrays.
// TEST INFRASTRUCTURE ONLY.
//
// What runs here is reference code: ArtemisUtils::CalculateFluxes (fluid_fluxes.hpp:76-290,
// Reconstruction<>, RiemannSolver<>, ScaleMomentumFlux), ArtemisUtils::FluxSource
// (fluid_fluxes.hpp:298-480), ArtemisUtils::ApplyUpdate / DeepCopyConservedData
// (artemis_integrator.hpp), ArtemisDerived::SetAuxillaryFields / ConsToPrim / PrimToCons
// (derived/fill_derived.cpp), geometry::Coords<GEOM> (geometry/*.hpp).  What is restated in
// this file: the pack construction of Gas/Dust::CalculateFluxes and ::FluxSource
// (src/gas/gas.cpp:473-519, src/dust/dust.cpp:281-326 -- those .cpp files pull in all of
// Artemis) and the timestep loop body of Gas/Dust::EstimateTimestepMesh
// (src/gas/gas.cpp:411-433, src/dust/dust.cpp:255-272).
#include "derived/fill_derived.cpp"
#include "utils/fluxes/fluid_fluxes.hpp"
#include "utils/integrators/artemis_integrator.hpp"
// source terms between FluxSource and SetAuxillaryFields
#include "rotating_frame/rotating_frame_impl.hpp"
#include "gravity/uniform.cpp"
#include "gravity/point_mass.cpp"
// diffusion operators
#include "utils/diffusion/diffusion.hpp"
#include "utils/diffusion/momentum_diffusion.hpp"
#include "utils/diffusion/thermal_diffusion.hpp"
// Drag::Initialize + Drag::DragSource<GEOM> with SelfDragSourceImpl / SimpleDragSourceImpl
#include "drag/drag.cpp"

#include "../artemis_oracle.h"

namespace {
using ArtemisUtils::EOS;

struct Ctx {
  Mesh mesh;
  MeshData<Real> md;
};
double g_cv = 1.0;  // specific heat handed to the EOS (only the diffusion operators read it)

void SetGrid(Ctx &c, const ao_grid *g) {
  c.mesh.ndim = g->ndim;
  parthenon::Globals::nghost = g->ng;
  auto &md = c.md;
  md.pm = &c.mesh;
  md.nb = g->nb; md.ni = g->ni; md.nj = g->nj; md.nk = g->nk;
  md.fni = g->fni; md.fnj = g->fnj; md.fnk = g->fnk;
  md.ib = {g->is, g->ie}; md.jb = {g->js, g->je}; md.kb = {g->ks, g->ke};
  md.coords.resize(g->nb);
  for (int b = 0; b < g->nb; ++b)
    for (int d = 0; d < 3; ++d) {
      md.coords[b].xmin_[d] = g->xmin[3 * b + d];
      md.coords[b].dx_[d] = g->dx[3 * b + d];
    }
  auto art = std::make_shared<StateDescriptor>();
  art->AddParam<bool>("do_gas", false);
  art->AddParam<bool>("do_dust", false);
  art->AddParam<bool>("do_rotating_frame", false);
  c.mesh.packages.pkgs["artemis"] = art;
}

void SetFluidPkg(Ctx &c, const ao_grid *g, const ao_fluid *f, double omf = 0.0) {
  auto pkg = std::make_shared<StateDescriptor>();
  pkg->AddParam<int>("nspecies", f->nspecies);
  pkg->AddParam<int>("scr_level", 0);
  pkg->AddParam<ReconstructionMethod>("recon", static_cast<ReconstructionMethod>(f->recon));
  pkg->AddParam<RSolver>("rsolver", static_cast<RSolver>(f->riemann));
  pkg->AddParam<Coordinates>("coords", static_cast<Coordinates>(g->geom));
  pkg->AddParam<Fluid>("fluid_type", static_cast<Fluid>(f->fluid));
  pkg->AddParam<Real>("dfloor", f->dfloor);
  pkg->AddParam<Real>("cfl", f->cfl);
  if (f->fluid == AO_GAS) {
    pkg->AddParam<Real>("siefloor", f->siefloor);
    pkg->AddParam<Real>("de_switch", f->de_switch);
    pkg->AddParam<EOS>("eos_d", EOS(f->gm1, g_cv));  // src/gas/gas.cpp:104-117
  }
  const bool gas = f->fluid == AO_GAS;
  c.mesh.packages.pkgs[gas ? "gas" : "dust"] = pkg;
  auto art = std::make_shared<StateDescriptor>(*c.mesh.packages.Get("artemis"));
  art->AddParam<bool>(gas ? "do_gas" : "do_dust", true);
  if (omf != 0.0) {
    art->AddParam<bool>("do_rotating_frame", true);
    auto rf = std::make_shared<StateDescriptor>();
    rf->AddParam<Real>("omega", omf);
    c.mesh.packages.pkgs["rotating_frame"] = rf;
  }
  c.mesh.packages.pkgs["artemis"] = art;
}

// slices of the [nb][nvar][cells] slabs of artemis_oracle.h as named Parthenon fields
void AddSlab(Ctx &c, const ao_grid *g, const ao_fluid *f, bool prim, double *a, double *fl[3],
             double *pfl[3], double *vf[3]) {
  const int S = f->nspecies;
  const bool gas = f->fluid == AO_GAS;
  const int nvar = (gas ? 6 : 4) * S;
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  const size_t fcells = (size_t)g->fni * g->fnj * g->fnk;
  const std::string base = std::string(gas ? "gas." : "dust.") + (prim ? "prim." : "cons.");
  auto add = [&](const std::string &nm, int off, int ncomp, bool with_flux) {
    Field fd;
    fd.name = base + nm;
    fd.ncomp = ncomp;
    fd.conserved = !prim;
    fd.data = a ? a + off * cells : nullptr;
    fd.bstride = (size_t)nvar * cells;
    if (with_flux && fl)
      for (int d = 0; d < 3; ++d) fd.flux[d] = fl[d] ? fl[d] + off * cells : nullptr;
    fd.flux_bstride = (size_t)nvar * cells;
    c.md.fields.push_back(fd);
  };
  add("density", 0, S, true);
  add(prim ? "velocity" : "momentum", S, 3 * S, true);
  if (gas) {
    if (prim) {
      Field fd;  // gas.prim.pressure: its flux slot is the interface pressure
      fd.name = "gas.prim.pressure";
      fd.ncomp = S;
      fd.data = a ? a + 4 * S * cells : nullptr;
      fd.bstride = (size_t)nvar * cells;
      if (pfl)
        for (int d = 0; d < 3; ++d) fd.flux[d] = pfl[d];
      fd.flux_bstride = (size_t)S * cells;
      c.md.fields.push_back(fd);
      add("sie", 5 * S, S, false);
      if (vf) {
        Field fv;
        fv.name = "gas.face.velocity";
        fv.ncomp = S;
        for (int d = 0; d < 3; ++d) fv.face[d] = vf[d];
        fv.face_bstride = (size_t)S * fcells;
        c.md.fields.push_back(fv);
      }
    } else {
      add("total_energy", 4 * S, S, true);
      add("internal_energy", 5 * S, S, true);
    }
  }
}

template <typename F>
void GeomDispatch(int geom, F &&fn) {
  switch (geom) {
  case AO_CARTESIAN: fn(std::integral_constant<Coordinates, Coordinates::cartesian>{}); break;
  case AO_CYLINDRICAL: fn(std::integral_constant<Coordinates, Coordinates::cylindrical>{}); break;
  case AO_SPHERICAL1D: fn(std::integral_constant<Coordinates, Coordinates::spherical1D>{}); break;
  case AO_SPHERICAL2D: fn(std::integral_constant<Coordinates, Coordinates::spherical2D>{}); break;
  case AO_SPHERICAL3D: fn(std::integral_constant<Coordinates, Coordinates::spherical3D>{}); break;
  case AO_AXISYMMETRIC: fn(std::integral_constant<Coordinates, Coordinates::axisymmetric>{}); break;
  default: PARTHENON_FAIL("Coordinate type not recognized!");
  }
}
}  // namespace

extern "C" {

void ar_calculate_fluxes(const ao_grid *g, const ao_fluid *f, int pcm, double *prim,
                         double *flux1, double *flux2, double *flux3, double *pflux1,
                         double *pflux2, double *pflux3, double *vface1, double *vface2,
                         double *vface3) {
  Ctx c;
  SetGrid(c, g);
  SetFluidPkg(c, g, f);
  double *fl[3] = {flux1, flux2, flux3}, *pfl[3] = {pflux1, pflux2, pflux3},
         *vf[3] = {vface1, vface2, vface3};
  AddSlab(c, g, f, true, prim, nullptr, pfl, vf);
  AddSlab(c, g, f, false, nullptr, fl, nullptr, nullptr);
  MeshData<Real> *md = &c.md;
  auto pm = md->GetParentPointer();
  auto &resolved_pkgs = pm->resolved_packages;
  if (f->fluid == AO_GAS) {  // pack construction of Gas::CalculateFluxes, src/gas/gas.cpp:473-494
    auto &pkg = pm->packages.Get("gas");
    auto desc_prim = parthenon::MakePackDescriptor<gas::prim::density, gas::prim::velocity,
                                                   gas::prim::pressure, gas::prim::sie>(
        resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
    auto desc_flux = parthenon::MakePackDescriptor<gas::cons::density, gas::cons::momentum,
                                                   gas::cons::total_energy,
                                                   gas::cons::internal_energy>(
        resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
    auto desc_face = parthenon::MakePackDescriptor<gas::face::velocity>(resolved_pkgs.get());
    auto vprim = desc_prim.GetPack(md);
    auto vflux = desc_flux.GetPack(md);
    auto vface = desc_face.GetPack(md);
    ArtemisUtils::CalculateFluxes<Fluid::gas>(md, pkg, vprim, vflux, vface, pcm != 0);
  } else {  // Dust::CalculateFluxes, src/dust/dust.cpp:281-298
    auto &pkg = pm->packages.Get("dust");
    auto desc_prim = parthenon::MakePackDescriptor<dust::prim::density, dust::prim::velocity>(
        resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
    auto desc_flux = parthenon::MakePackDescriptor<dust::cons::density, dust::cons::momentum>(
        resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
    auto vprim = desc_prim.GetPack(md);
    auto vflux = desc_flux.GetPack(md);
    parthenon::SparsePackShim vface;
    ArtemisUtils::CalculateFluxes<Fluid::dust>(md, pkg, vprim, vflux, vface, pcm != 0);
  }
}

void ar_apply_update(const ao_grid *g, int nvar, double *u0, const double *u1,
                     const double *flux1, const double *flux2, const double *flux3,
                     double gam0, double gam1, double beta_dt) {
  // two MeshData (u0, u1) sharing the flux arrays; one "Conserved" field of nvar components
  Ctx c0, c1;
  SetGrid(c0, g);
  SetGrid(c1, g);
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  Field fd;
  fd.name = "conserved";
  fd.ncomp = nvar;
  fd.conserved = true;
  fd.bstride = fd.flux_bstride = (size_t)nvar * cells;
  fd.flux[0] = const_cast<double *>(flux1);
  fd.flux[1] = const_cast<double *>(flux2);
  fd.flux[2] = const_cast<double *>(flux3);
  fd.data = u0;
  c0.md.fields.push_back(fd);
  fd.data = const_cast<double *>(u1);
  c1.md.fields.push_back(fd);
  parthenon::LowStorageIntegrator integ;
  integ.dt = 1.0;  // beta_dt = beta[stage-1] * dt, artemis_integrator.hpp:66
  integ.gam0 = {gam0}; integ.gam1 = {gam1}; integ.beta = {beta_dt};
  GeomDispatch(g->geom, [&](auto G) {
    ArtemisUtils::ApplyUpdate<decltype(G)::value>(&c0.md, &c1.md, 1, &integ);
  });
}

void ar_flux_source(const ao_grid *g, const ao_fluid *f, double *prim, double *cons,
                    double *pflux1, double *pflux2, double *pflux3, double *vface1,
                    double *vface2, double *vface3, double omf, double dt) {
  Ctx c;
  SetGrid(c, g);
  SetFluidPkg(c, g, f, omf);
  double *pfl[3] = {pflux1, pflux2, pflux3}, *vf[3] = {vface1, vface2, vface3};
  AddSlab(c, g, f, true, prim, nullptr, pfl, vf);
  AddSlab(c, g, f, false, cons, nullptr, nullptr, nullptr);
  MeshData<Real> *md = &c.md;
  auto pm = md->GetParentPointer();
  auto &resolved_pkgs = pm->resolved_packages;
  if (f->fluid == AO_GAS) {  // Gas::FluxSource, src/gas/gas.cpp:499-519
    auto &pkg = pm->packages.Get("gas");
    auto desc_prim = parthenon::MakePackDescriptor<gas::prim::density, gas::prim::velocity,
                                                   gas::prim::pressure>(
        resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
    auto desc_cons =
        parthenon::MakePackDescriptor<gas::cons::momentum, gas::cons::internal_energy>(
            resolved_pkgs.get());
    auto desc_face = parthenon::MakePackDescriptor<gas::face::velocity>(resolved_pkgs.get());
    auto vprim = desc_prim.GetPack(md);
    auto vcons = desc_cons.GetPack(md);
    auto vface = desc_face.GetPack(md);
    ArtemisUtils::FluxSource(md, pkg, vprim, vcons, vface, dt);
  } else {  // Dust::FluxSource, src/dust/dust.cpp:303-326
    auto &pkg = pm->packages.Get("dust");
    auto sys = pkg->Param<Coordinates>("coords");
    if (geometry::x1dep(sys) || ((geometry::x2dep(sys)) && (pm->ndim >= 2)) ||
        ((geometry::x3dep(sys)) && (pm->ndim == 3))) {
      auto desc_prim = parthenon::MakePackDescriptor<dust::prim::density, dust::prim::velocity>(
          resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
      auto desc_cons = parthenon::MakePackDescriptor<dust::cons::momentum>(resolved_pkgs.get());
      auto vprim = desc_prim.GetPack(md);
      auto vcons = desc_cons.GetPack(md);
      parthenon::SparsePackShim vface;
      ArtemisUtils::FluxSource(md, pkg, vprim, vcons, vface, dt);
    }
  }
}

void ar_set_aux(const ao_grid *g, const ao_fluid *f, double *cons) {
  if (f->fluid != AO_GAS) return;
  Ctx c;
  SetGrid(c, g);
  SetFluidPkg(c, g, f);
  AddSlab(c, g, f, false, cons, nullptr, nullptr, nullptr);
  GeomDispatch(g->geom, [&](auto G) {
    ArtemisDerived::SetAuxillaryFields<decltype(G)::value>(&c.md);
  });
}

void ar_cons_to_prim(const ao_grid *g, const ao_fluid *f, double *cons, double *prim) {
  Ctx c;
  SetGrid(c, g);
  SetFluidPkg(c, g, f);
  AddSlab(c, g, f, false, cons, nullptr, nullptr, nullptr);
  AddSlab(c, g, f, true, prim, nullptr, nullptr, nullptr);
  GeomDispatch(g->geom,
               [&](auto G) { ArtemisDerived::ConsToPrim<decltype(G)::value>(&c.md); });
}

void ar_prim_to_cons(const ao_grid *g, const ao_fluid *f, double *prim, double *cons) {
  Ctx c;
  SetGrid(c, g);
  SetFluidPkg(c, g, f);
  AddSlab(c, g, f, false, cons, nullptr, nullptr, nullptr);
  AddSlab(c, g, f, true, prim, nullptr, nullptr, nullptr);
  GeomDispatch(g->geom, [&](auto G) {
    ArtemisDerived::PrimToCons<MeshData<Real>, decltype(G)::value>(&c.md);
  });
}

void ar_deep_copy(const ao_grid *g, int nvar, double *to, const double *from) {
  Ctx c0, c1;
  SetGrid(c0, g);
  SetGrid(c1, g);
  const size_t cells = (size_t)g->ni * g->nj * g->nk;
  Field fd;
  fd.name = "conserved";
  fd.ncomp = nvar;
  fd.conserved = true;
  fd.bstride = (size_t)nvar * cells;
  fd.data = to;
  c0.md.fields.push_back(fd);
  fd.data = const_cast<double *>(from);
  c1.md.fields.push_back(fd);
  ArtemisUtils::DeepCopyConservedData(&c0.md, &c1.md);
}

// Gas/Dust::EstimateTimestepMesh: gas.cpp / dust.cpp cannot be compiled here; the loop body is
// restated with the reference's geometry::Coords<GEOM>::GetCellWidths and EOS
// (src/gas/gas.cpp:411-433, :465-467; src/dust/dust.cpp:255-275)
double ar_estimate_dt(const ao_grid *g, const ao_fluid *f, const double *prim) {
  Ctx c;
  SetGrid(c, g);
  SetFluidPkg(c, g, f);
  AddSlab(c, g, f, true, const_cast<double *>(prim), nullptr, nullptr, nullptr);
  const int S = f->nspecies, ndim = g->ndim;
  const bool gas = f->fluid == AO_GAS;
  Real min_dt = Big<Real>();
  GeomDispatch(g->geom, [&](auto G) {
    constexpr Coordinates GEOM = decltype(G)::value;
    const size_t cells = (size_t)g->ni * g->nj * g->nk;
    const int nvar = (gas ? 6 : 4) * S;
    EOS eos_d(f->gm1, 1.0);
#pragma omp parallel for collapse(3) reduction(min : min_dt) schedule(static)
    for (int b = 0; b < g->nb; ++b)
      for (int k = g->ks; k <= g->ke; ++k)
        for (int j = g->js; j <= g->je; ++j)
          for (int i = g->is; i <= g->ie; ++i) {
            geometry::Coords<GEOM> coords(c.md.coords[b], k, j, i);
            const auto &dx = coords.GetCellWidths();
            const size_t off = ((size_t)k * g->nj + j) * g->ni + i;
            auto at = [&](int n) { return prim[((size_t)b * nvar + n) * cells + off]; };
            for (int n = 0; n < S; ++n) {
              Real cs = 0.0;
              if (gas) {
                const Real dens = at(n), sie = at(5 * S + n);
                const Real bulk = eos_d.BulkModulusFromDensityInternalEnergy(dens, sie);
                cs = std::sqrt(bulk / dens);
              }
              Real denom = 0.0;
              for (int d = 0; d < ndim; d++) {
                const Real ss = std::abs(at(S + ArtemisUtils::VI(n, d))) + cs;
                denom += ss / dx[d];
              }
              min_dt = std::min(min_dt, 1.0 / denom);
            }
          }
  });
  return f->cfl * min_dt;
}

int ar_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}

// ---- pointwise source terms: the reference's own kernels over BOTH fluids of one MeshData ---
static void SetBoth(Ctx &c, const ao_grid *g, const ao_fluid *gas, double *gprim, double *gcons,
                    const ao_fluid *dust, double *dprim, double *dcons) {
  SetGrid(c, g);
  if (gas) {
    SetFluidPkg(c, g, gas);
    AddSlab(c, g, gas, true, gprim, nullptr, nullptr, nullptr);
    AddSlab(c, g, gas, false, gcons, nullptr, nullptr, nullptr);
  }
  if (dust) {
    SetFluidPkg(c, g, dust);
    AddSlab(c, g, dust, true, dprim, nullptr, nullptr, nullptr);
    AddSlab(c, g, dust, false, dcons, nullptr, nullptr, nullptr);
  }
}

void ar_shearing_box(const ao_grid *g, const ao_fluid *gas, double *gprim, double *gcons,
                     const ao_fluid *dust, double *dprim, double *dcons, double dt, double om0,
                     double qshear) {
  Ctx c;
  SetBoth(c, g, gas, gprim, gcons, dust, dprim, dcons);
  RotatingFrame::ShearingBoxImpl(&c.md, om0, qshear, gas != nullptr, dust != nullptr, dt);
}

void ar_uniform_gravity(const ao_grid *g, const ao_fluid *gas, double *gprim, double *gcons,
                        const ao_fluid *dust, double *dprim, double *dcons, double dt, double gx1,
                        double gx2, double gx3) {
  Ctx c;
  SetBoth(c, g, gas, gprim, gcons, dust, dprim, dcons);
  auto grav = std::make_shared<StateDescriptor>();
  grav->AddParam<Real>("gx1", gx1);
  grav->AddParam<Real>("gx2", gx2);
  grav->AddParam<Real>("gx3", gx3);
  c.mesh.packages.pkgs["gravity"] = grav;
  GeomDispatch(g->geom, [&](auto G) {
    Gravity::UniformGravity<decltype(G)::value>(&c.md, 0.0, dt);
  });
}

// Gravity::PointMassGravity<GEOM> (src/gravity/point_mass.cpp:26-196); pm = {gm, x, y, z, soft,
// sink_rate, sink}: the "gravity" package parameters it reads.
void ar_point_mass_gravity(const ao_grid *g, const ao_fluid *gas, double *gprim, double *gcons,
                           const ao_fluid *dust, double *dprim, double *dcons, double dt,
                           const double *pm) {
  Ctx c;
  SetBoth(c, g, gas, gprim, gcons, dust, dprim, dcons);
  auto grav = std::make_shared<StateDescriptor>();
  grav->AddParam<Real>("gm", pm[0]);
  grav->AddParam<Real>("x", pm[1]);
  grav->AddParam<Real>("y", pm[2]);
  grav->AddParam<Real>("z", pm[3]);
  grav->AddParam<Real>("soft", pm[4]);
  grav->AddParam<Real>("sink_rate", pm[5]);
  grav->AddParam<Real>("sink", pm[6]);
  c.mesh.packages.pkgs["gravity"] = grav;
  GeomDispatch(g->geom, [&](auto G) {
    Gravity::PointMassGravity<decltype(G)::value>(&c.md, 0.0, dt);
  });
}

// RotatingFrame::RotatingFrameImpl<GEOM> (rotating_frame_impl.hpp:96-199), the curvilinear
// branch of RotatingFrameForce (rotating_frame.cpp:54-86): reads the DENSITY fluxes of the
// stage.  gflux / dflux: [3] slabs [nb][nvar][cells] as written by ar_calculate_fluxes.
void ar_rotating_frame(const ao_grid *g, const ao_fluid *gas, double *gcons, double *gflux1,
                       double *gflux2, double *gflux3, const ao_fluid *dust, double *dcons,
                       double *dflux1, double *dflux2, double *dflux3, double dt, double om0) {
  Ctx c;
  SetGrid(c, g);
  double *gf[3] = {gflux1, gflux2, gflux3}, *df[3] = {dflux1, dflux2, dflux3};
  if (gas) {
    SetFluidPkg(c, g, gas);
    AddSlab(c, g, gas, false, gcons, gf, nullptr, nullptr);
  }
  if (dust) {
    SetFluidPkg(c, g, dust);
    AddSlab(c, g, dust, false, dcons, df, nullptr, nullptr);
  }
  GeomDispatch(g->geom, [&](auto G) {
    constexpr Coordinates GG = decltype(G)::value;
    if constexpr (GG != Coordinates::cartesian)
      RotatingFrame::RotatingFrameImpl<GG>(&c.md, om0, gas != nullptr, dust != nullptr, dt);
  });
}
// ---- diffusion: the reference's own MomentumFluxImpl / ThermalFluxImpl / DiffusionUpdateImpl /
// EstimateTimestep / ZeroDiffusionImpl; the type dispatch of Gas::ViscousFlux / ThermalFlux /
// ZeroDiffusionFlux / DiffusionUpdate (src/gas/gas.cpp:524-642, that .cpp pulls in all of
// Artemis) and of the timestep (src/gas/gas.cpp:437-464) is restated here.
static void SetDiffusion(Ctx &c, const ao_grid *g, const ao_fluid *gas, double *gprim,
                         double *gcons, const ao_diffusion *dd, double *dflx[3],
                         Diffusion::DiffCoeffParams &vp, Diffusion::DiffCoeffParams &cp) {
  g_cv = dd->cv;
  SetGrid(c, g);
  SetFluidPkg(c, g, gas);
  g_cv = 1.0;
  AddSlab(c, g, gas, true, gprim, nullptr, nullptr, nullptr);
  if (gcons) AddSlab(c, g, gas, false, gcons, nullptr, nullptr, nullptr);
  const int S = gas->nspecies;
  const size_t fcells = (size_t)g->fni * g->fnj * g->fnk;
  if (dflx) {
    Field fm, fe;
    fm.name = "gas.diff.momentum"; fm.ncomp = 3 * S;
    fe.name = "gas.diff.energy";   fe.ncomp = S;
    for (int d = 0; d < 3; ++d) {
      fm.face[d] = dflx[d];
      fe.face[d] = dflx[d] ? dflx[d] + (size_t)3 * S * fcells : nullptr;
    }
    fm.face_bstride = fe.face_bstride = (size_t)4 * S * fcells;
    c.md.fields.push_back(fm);
    c.md.fields.push_back(fe);
  }
  using Diffusion::DiffType;
  using Diffusion::DiffAvg;
  vp = Diffusion::DiffCoeffParams();
  cp = Diffusion::DiffCoeffParams();
  vp.type = dd->visc_type == AO_VISC_PLAW ? DiffType::viscosity_plaw
            : dd->visc_type == AO_VISC_ALPHA ? DiffType::viscosity_alpha : DiffType::null;
  vp.avg = dd->visc_avg == AO_AVG_HARMONIC ? DiffAvg::harmonic : DiffAvg::arithmetic;
  vp.nu_s = dd->nu; vp.eta = dd->eta; vp.R0 = dd->r0; vp.r_exp = dd->r_exp;
  vp.alpha = dd->alpha; vp.Omega0 = dd->omega0;
  cp.type = dd->cond_type == AO_COND_CONDUCTIVITY ? DiffType::conductivity_plaw
            : dd->cond_type == AO_COND_DIFFUSIVITY ? DiffType::thermaldiff_plaw : DiffType::null;
  cp.avg = dd->cond_avg == AO_AVG_HARMONIC ? DiffAvg::harmonic : DiffAvg::arithmetic;
  cp.hcond_0 = dd->cond; cp.kappa_0 = dd->kappa; cp.temp_exp = dd->temp_exp;
  cp.rho_exp = dd->rho_exp; cp.d0 = dd->rho_ref; cp.T0 = dd->t_ref;
}

void ar_diffusion_flux(const ao_grid *g, const ao_fluid *gas, double *gprim,
                       const ao_diffusion *dd, double *dflx1, double *dflx2, double *dflx3) {
  Ctx c;
  double *dflx[3] = {dflx1, dflx2, dflx3};
  Diffusion::DiffCoeffParams vp, cp;
  SetDiffusion(c, g, gas, gprim, nullptr, dd, dflx, vp, cp);
  MeshData<Real> *md = &c.md;
  auto pm = md->GetParentPointer();
  auto &pkg = pm->packages.Get("gas");
  auto &resolved_pkgs = pm->resolved_packages;
  using Diffusion::DiffType;
  GeomDispatch(g->geom, [&](auto G) {
    constexpr Coordinates GG = decltype(G)::value;
    {  // Gas::ZeroDiffusionFlux, src/gas/gas.cpp:592-603
      auto desc_flux = parthenon::MakePackDescriptor<gas::diff::momentum, gas::diff::energy>(
          resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
      auto vf = desc_flux.GetPack(md);
      Diffusion::ZeroDiffusionImpl(md, vf);
    }
    if (vp.type != DiffType::null) {  // Gas::ViscousFlux, src/gas/gas.cpp:524-556
      auto desc_prim = parthenon::MakePackDescriptor<gas::prim::density, gas::prim::velocity,
                                                     gas::prim::sie>(resolved_pkgs.get());
      auto desc_flux = parthenon::MakePackDescriptor<gas::diff::momentum, gas::diff::energy>(
          resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
      auto vprim = desc_prim.GetPack(md);
      auto vf = desc_flux.GetPack(md);
      if (vp.type == DiffType::viscosity_plaw)
        Diffusion::MomentumFluxImpl<GG, Fluid::gas, DiffType::viscosity_plaw>(md, vp, pkg, vprim, vf);
      else
        Diffusion::MomentumFluxImpl<GG, Fluid::gas, DiffType::viscosity_alpha>(md, vp, pkg, vprim, vf);
    }
    if (cp.type != DiffType::null) {  // Gas::ThermalFlux, src/gas/gas.cpp:561-587
      auto desc_prim = parthenon::MakePackDescriptor<gas::prim::density, gas::prim::sie>(
          resolved_pkgs.get());
      auto desc_flux = parthenon::MakePackDescriptor<gas::diff::energy>(
          resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
      auto vprim = desc_prim.GetPack(md);
      auto vf = desc_flux.GetPack(md);
      if (cp.type == DiffType::conductivity_plaw)
        Diffusion::ThermalFluxImpl<GG, Fluid::gas, DiffType::conductivity_plaw>(md, cp, pkg, vprim, vf);
      else
        Diffusion::ThermalFluxImpl<GG, Fluid::gas, DiffType::thermaldiff_plaw>(md, cp, pkg, vprim, vf);
    }
  });
}

void ar_diffusion_update(const ao_grid *g, const ao_fluid *gas, double *gprim, double *gcons,
                         const ao_diffusion *dd, double *dflx1, double *dflx2, double *dflx3,
                         double dt) {  // Gas::DiffusionUpdate, src/gas/gas.cpp:608-642
  Ctx c;
  double *dflx[3] = {dflx1, dflx2, dflx3};
  Diffusion::DiffCoeffParams vp, cp;
  SetDiffusion(c, g, gas, gprim, gcons, dd, dflx, vp, cp);
  MeshData<Real> *md = &c.md;
  auto pm = md->GetParentPointer();
  auto &pkg = pm->packages.Get("gas");
  auto &resolved_pkgs = pm->resolved_packages;
  const bool do_viscosity = vp.type != Diffusion::DiffType::null;
  GeomDispatch(g->geom, [&](auto G) {
    constexpr Coordinates GG = decltype(G)::value;
    auto desc_cons = parthenon::MakePackDescriptor<gas::cons::momentum, gas::cons::total_energy,
                                                   gas::cons::internal_energy>(resolved_pkgs.get());
    auto desc_prim = parthenon::MakePackDescriptor<gas::prim::velocity>(resolved_pkgs.get());
    auto desc_flux = parthenon::MakePackDescriptor<gas::diff::momentum, gas::diff::energy>(
        resolved_pkgs.get(), {}, {parthenon::PDOpt::WithFluxes});
    auto vcons = desc_cons.GetPack(md);
    auto vprim = desc_prim.GetPack(md);
    auto vf = desc_flux.GetPack(md);
    Diffusion::DiffusionUpdateImpl<GG, Fluid::gas>(md, pkg, vcons, vprim, vf, do_viscosity, dt);
  });
}

double ar_diffusion_dt(const ao_grid *g, const ao_fluid *gas, double *gprim,
                       const ao_diffusion *dd) {  // src/gas/gas.cpp:437-464
  Ctx c;
  Diffusion::DiffCoeffParams vp, cp;
  SetDiffusion(c, g, gas, gprim, nullptr, dd, nullptr, vp, cp);
  MeshData<Real> *md = &c.md;
  auto pm = md->GetParentPointer();
  auto &gas_pkg = pm->packages.Get("gas");
  auto &resolved_pkgs = pm->resolved_packages;
  auto eos_d = gas_pkg->Param<EOS>("eos_d");
  using Diffusion::DiffType;
  Real visc_dt = Big<Real>(), cond_dt = Big<Real>();
  GeomDispatch(g->geom, [&](auto G) {
    constexpr Coordinates GG = decltype(G)::value;
    auto desc = parthenon::MakePackDescriptor<gas::prim::density, gas::prim::velocity,
                                              gas::prim::sie>(resolved_pkgs.get());
    auto vmesh = desc.GetPack(md);
    if (vp.type == DiffType::viscosity_plaw)
      visc_dt = Diffusion::EstimateTimestep<GG, Fluid::gas, DiffType::viscosity_plaw>(md, vp, gas_pkg, eos_d, vmesh);
    else if (vp.type == DiffType::viscosity_alpha)
      visc_dt = Diffusion::EstimateTimestep<GG, Fluid::gas, DiffType::viscosity_alpha>(md, vp, gas_pkg, eos_d, vmesh);
    if (cp.type == DiffType::conductivity_plaw)
      cond_dt = Diffusion::EstimateTimestep<GG, Fluid::gas, DiffType::conductivity_plaw>(md, cp, gas_pkg, eos_d, vmesh);
    else if (cp.type == DiffType::thermaldiff_plaw)
      cond_dt = Diffusion::EstimateTimestep<GG, Fluid::gas, DiffType::thermaldiff_plaw>(md, cp, gas_pkg, eos_d, vmesh);
  });
  return std::min(visc_dt, cond_dt);
}
// ---- drag: the reference's own Drag::Initialize (parameter parsing from an input deck) and
// Drag::DragSource<GEOM> (dispatch + SelfDragSourceImpl / SimpleDragSourceImpl), src/drag/drag.cpp
// and drag.hpp, driven from a ParameterInput the harness fills like the deck would.
void ar_drag_source(const ao_grid *g, const ao_fluid *gas, double *gcons, const ao_fluid *dust,
                    double *dcons, const ao_drag *dp, const ao_diffusion *dd, double dt) {
  Ctx c;
  SetGrid(c, g);
  if (dd) g_cv = dd->cv;
  if (gas) {
    SetFluidPkg(c, g, gas);
    AddSlab(c, g, gas, false, gcons, nullptr, nullptr, nullptr);
  }
  g_cv = 1.0;
  if (dust) {
    SetFluidPkg(c, g, dust);
    AddSlab(c, g, dust, false, dcons, nullptr, nullptr, nullptr);
  }
  parthenon::ParameterInput pin;
  pin.Set("drag", "type", dp->coupling == AO_DRAG_SELF ? "self" : "simple_dust");
  const char *xn[3] = {"x1", "x2", "x3"};
  for (int d = 0; d < 3; ++d) {
    pin.Set("parthenon/mesh", std::string(xn[d]) + "min", dp->xmin[d]);
    pin.Set("parthenon/mesh", std::string(xn[d]) + "max", dp->xmax[d]);
  }
  pin.Set("physics", "gas", gas ? "true" : "false");
  pin.Set("physics", "dust", dust ? "true" : "false");
  auto damping = [&](const char *blk, const double *ix, const double *ox, const double *ir,
                     const double *orr, int to_visc) {
    for (int d = 0; d < 3; ++d) {
      pin.Set(blk, "inner_" + std::string(xn[d]), ix[d]);
      pin.Set(blk, "outer_" + std::string(xn[d]), ox[d]);
      pin.Set(blk, "inner_" + std::string(xn[d]) + "_rate", ir[d]);
      pin.Set(blk, "outer_" + std::string(xn[d]) + "_rate", orr[d]);
    }
    pin.Set(blk, "damp_to_visc", to_visc ? "true" : "false");
  };
  if (gas) damping("gas/damping", dp->g_ix, dp->g_ox, dp->g_irate, dp->g_orate, dp->g_damp_to_visc);
  if (dust) damping("dust/damping", dp->d_ix, dp->d_ox, dp->d_irate, dp->d_orate, 0);
  if (dust) {
    const int nd = dust->nspecies;
    pin.Set("dust", "nspecies", (Real)nd);
    pin.Set("dust/stopping_time", "type", dp->model == AO_DRAG_STOKES ? "stokes" : "constant");
    pin.Set("dust/stopping_time", "scale", dp->scale);
    std::ostringstream o;
    o.precision(17);
    for (int n = 0; n < nd; ++n) o << (n ? "," : "") << dp->tau[n];
    pin.Set("dust/stopping_time", "tau", o.str());
    // <dust> sizes / grain_density as Dust::Initialize stores them (src/dust/dust.cpp:102-140)
    auto dpk = std::make_shared<StateDescriptor>(*c.mesh.packages.Get("dust"));
    parthenon::ParArray1D<Real> sizes("sizes", nd);
    for (int n = 0; n < nd; ++n) sizes(n) = dp->sizes[n];
    dpk->AddParam<parthenon::ParArray1D<Real>>("sizes", sizes);
    dpk->AddParam<Real>("grain_density", dp->grain_density);
    c.mesh.packages.pkgs["dust"] = dpk;
  }
  if (gas && dd) {
    Diffusion::DiffCoeffParams vp, cp;
    double *none[3] = {nullptr, nullptr, nullptr};
    (void)none;
    using Diffusion::DiffType;
    vp.type = dd->visc_type == AO_VISC_PLAW ? DiffType::viscosity_plaw
              : dd->visc_type == AO_VISC_ALPHA ? DiffType::viscosity_alpha : DiffType::null;
    vp.avg = Diffusion::DiffAvg::arithmetic;
    vp.nu_s = dd->nu; vp.eta = dd->eta; vp.R0 = dd->r0; vp.r_exp = dd->r_exp;
    vp.alpha = dd->alpha; vp.Omega0 = dd->omega0;
    auto gpk = std::make_shared<StateDescriptor>(*c.mesh.packages.Get("gas"));
    gpk->AddParam<Diffusion::DiffCoeffParams>("visc_params", vp);
    c.mesh.packages.pkgs["gas"] = gpk;
  }
  c.mesh.packages.pkgs["drag"] = Drag::Initialize(&pin);
  auto art = std::make_shared<StateDescriptor>(*c.mesh.packages.Get("artemis"));
  art->AddParam<Coordinates>("coords", static_cast<Coordinates>(g->geom));
  c.mesh.packages.pkgs["artemis"] = art;
  GeomDispatch(g->geom, [&](auto G) {
    Drag::DragSource<decltype(G)::value>(&c.md, 0.0, dt);
  });
}
}  // extern "C"
