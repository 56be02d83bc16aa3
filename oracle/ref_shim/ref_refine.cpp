// ref_refine.cpp -- oracle/_ref entry points for the multilevel operators: the reference's OWN
// ArtemisUtils::RestrictAverage<GEOM> and ArtemisUtils::ProlongateSharedMinMod<GEOM>
// (src/utils/refinement/restriction.hpp:41-114, prolongation.hpp:82-184), compiled unmodified
// from /root/reference/src against the mock Parthenon (the two real headers are named by the
// build recipe through AR_RESTRICTION_HPP / AR_PROLONGATION_HPP because the include path
// resolves "utils/refinement/*.hpp" to the mocks used by the uniform-mesh translation unit).
// Restated here: the loop over the coarse index box (Parthenon's
// refinement::loops::ProlongationRestrictionLoop, P:prolong_restrict/prolong_restrict.hpp)
// and Parthenon's coarse-coordinates constructor (P:coordinates/uniform_cartesian.hpp:41-55).
// TEST INFRASTRUCTURE ONLY.
#include AR_PROLONGATION_HPP
#include AR_RESTRICTION_HPP

#include "../artemis_oracle.h"

namespace {
struct Setup {
  Coordinates_t fine, coarse;
  IndexRange ib, jb, kb, cib, cjb, ckb;
};

Setup MakeSetup(const ao_refine_geom *r) {
  Setup s;
  const int act[3] = {1, r->ndim > 1, r->ndim > 2};
  for (int d = 0; d < 3; ++d) {
    s.fine.xmin_[d] = r->xmin[d];
    s.fine.dx_[d] = r->dx[d];
    // UniformCartesian(const UniformCartesian &src, int coarsen), coarsen = 2
    const int istart = act[d] ? r->ng : 0;
    const int coarsen = 2;
    s.coarse.dx_[d] = r->dx[d];
    s.coarse.xmin_[d] = r->xmin[d];
    s.coarse.xmin_[d] += istart * s.coarse.dx_[d] * (1 - coarsen);
    s.coarse.dx_[d] *= (d == 0 ? coarsen : (istart > 0 ? coarsen : 1));
  }
  s.ib.s = r->ib_s; s.jb.s = r->jb_s; s.kb.s = r->kb_s;
  s.cib.s = r->cib_s; s.cjb.s = r->cjb_s; s.ckb.s = r->ckb_s;
  return s;
}

template <typename F>
void GeomDispatchR(int geom, F &&fn) {
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

template <template <Coordinates> class Op, int DIM, Coordinates G>
void Loop(const ao_refine_geom *r, int nvar, double *coarse, double *fine, const int *box) {
  Setup s = MakeSetup(r);
  ParArrayND<Real, VariableState> pc{coarse, nvar, r->cnk, r->cnj, r->cni};
  ParArrayND<Real, VariableState> pf{fine, nvar, r->nk, r->nj, r->ni};
  for (int n = 0; n < nvar; ++n)
    for (int k = box[4]; k <= box[5]; ++k)
      for (int j = box[2]; j <= box[3]; ++j)
        for (int i = box[0]; i <= box[1]; ++i)
          Op<G>::template Do<DIM, TE::CC, TE::CC>(0, 0, n, k, j, i, s.ckb, s.cjb, s.cib, s.kb,
                                                  s.jb, s.ib, s.fine, s.coarse, &pc, &pf);
}

// restriction of a face-centred (flux) field: RestrictAverage<GEOM>::Do<DIM, F1|F2|F3>
template <int DIM, Coordinates G, TE EL>
void LoopFace(const ao_refine_geom *r, int nvar, double *coarse, double *fine, const int *box) {
  Setup s = MakeSetup(r);
  ParArrayND<Real, VariableState> pc{coarse, nvar, r->cnk, r->cnj, r->cni};
  ParArrayND<Real, VariableState> pf{fine, nvar, r->nk, r->nj, r->ni};
  for (int n = 0; n < nvar; ++n)
    for (int k = box[4]; k <= box[5]; ++k)
      for (int j = box[2]; j <= box[3]; ++j)
        for (int i = box[0]; i <= box[1]; ++i)
          ArtemisUtils::RestrictAverage<G>::template Do<DIM, EL, EL>(
              0, 0, n, k, j, i, s.ckb, s.cjb, s.cib, s.kb, s.jb, s.ib, s.fine, s.coarse, &pc, &pf);
}
template <int DIM, Coordinates G>
void RunFace(const ao_refine_geom *r, int nvar, double *coarse, double *fine, const int *box,
             int el) {
  if (el == 1) LoopFace<DIM, G, TE::F1>(r, nvar, coarse, fine, box);
  else if (el == 2) LoopFace<DIM, G, TE::F2>(r, nvar, coarse, fine, box);
  else LoopFace<DIM, G, TE::F3>(r, nvar, coarse, fine, box);
}

template <template <Coordinates> class Op>
void Run(const ao_refine_geom *r, int nvar, double *coarse, double *fine, const int *box) {
  GeomDispatchR(r->geom, [&](auto G) {
    constexpr Coordinates C = decltype(G)::value;
    if (r->ndim == 1) Loop<Op, 1, C>(r, nvar, coarse, fine, box);
    else if (r->ndim == 2) Loop<Op, 2, C>(r, nvar, coarse, fine, box);
    else Loop<Op, 3, C>(r, nvar, coarse, fine, box);
  });
}
}  // namespace

extern "C" {
void ar_restrict_average(const ao_refine_geom *r, int nvar, const double *fine, double *coarse,
                         const int *box) {
  Run<ArtemisUtils::RestrictAverage>(r, nvar, coarse, const_cast<double *>(fine), box);
}
// el = 1..3: the field lives on x1 / x2 / x3 faces (flux correction, SetFluxCorrections)
void ar_restrict_average_face(const ao_refine_geom *r, int nvar, const double *fine,
                              double *coarse, const int *box, int el) {
  GeomDispatchR(r->geom, [&](auto G) {
    constexpr Coordinates C = decltype(G)::value;
    double *f = const_cast<double *>(fine);
    if (r->ndim == 1) RunFace<1, C>(r, nvar, coarse, f, box, el);
    else if (r->ndim == 2) RunFace<2, C>(r, nvar, coarse, f, box, el);
    else RunFace<3, C>(r, nvar, coarse, f, box, el);
  });
}
void ar_prolongate_minmod(const ao_refine_geom *r, int nvar, const double *coarse, double *fine,
                          const int *box) {
  Run<ArtemisUtils::ProlongateSharedMinMod>(r, nvar, const_cast<double *>(coarse), fine, box);
}
}
