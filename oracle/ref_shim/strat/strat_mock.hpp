// strat_mock.hpp -- the few Parthenon / Artemis names the six shearing-box boundary functions of
// src/pgen/strat.hpp:154-666 (strat::ExtrapInnerX1 ... strat::ExtrapOuterX3) touch, so that the
// functions can be sliced out of the reference tree at build time and compiled as they are
// (oracle/ref_shim/strat/build_strat_ref.py).  TEST INFRASTRUCTURE: pins the oracle's
// ao_strat_bc (oracle/artemis_oracle.c) to the reference's own code.  Nothing here is reference
// text.  parthenon::IndexShape comes from the same generated slice of mesh/domain.hpp that the
// GenericBC pin uses (oracle/ref_shim/bc/build_bc_ref.py); geometry::Coords is mocked for the
// Cartesian system only (cell centre = mean of the two faces, src/geometry/geometry.hpp:165-167
// -- the shearing box is Cartesian; the curvilinear centroids are pinned by test_ref_vs_oracle).
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cmath>
#include <cstddef>
#include <map>
#include <memory>
#include <string>
#include <unordered_map>
#include <vector>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]

namespace parthenon {
using Real = double;
constexpr int NDIM = 3;
enum CoordinateDirection { NODIR = -1, X0DIR = 0, X1DIR = 1, X2DIR = 2, X3DIR = 3 };
enum class TopologicalElement : std::size_t { CC = 0, F1 = 3, F2 = 4, F3 = 5, E1 = 6, E2 = 7, E3 = 8, NN = 9 };
using TE = TopologicalElement;
inline int TopologicalOffsetI(TE el) { return el == TE::F1; }
inline int TopologicalOffsetJ(TE el) { return el == TE::F2; }
inline int TopologicalOffsetK(TE el) { return el == TE::F3; }
struct IndexRange { int s = 0, e = 0; };

#include "bc_indexshape_generated.inc"

struct MakePackDescriptor {};  // only named in a using-declaration

// UniformCartesian::Xf (P:coordinates/uniform_cartesian.hpp): xmin + idx * dx
struct Coordinates_t {
  Real xmin[3] = {0, 0, 0}, dx[3] = {1, 1, 1};
  template <int dir>
  Real Xf(int idx) const { return xmin[dir - 1] + idx * dx[dir - 1]; }
};
}  // namespace parthenon

using parthenon::IndexDomain;
using parthenon::IndexRange;
using parthenon::Real;
enum class Coordinates { cartesian };

// variable-name tags: gas::prim::density(n) etc. carry the species / component index
struct VarTag { int fluid, kind, idx; };  // kind 0 density, 1 velocity, 2 sie
#define STRAT_VAR(ns_fluid, name, kind_)                                        \
  struct name : VarTag { name(int n = -1) : VarTag{ns_fluid, kind_, n} {} }
namespace gas { namespace prim { STRAT_VAR(0, density, 0); STRAT_VAR(0, velocity, 1); STRAT_VAR(0, sie, 2); } }
namespace dust { namespace prim { STRAT_VAR(1, density, 0); STRAT_VAR(1, velocity, 1); } }
#undef STRAT_VAR

namespace strat { struct StratParams; }

// one block, fine arrays or coarse buffer: gas prim [6 S][nk][nj][ni] (density n | velocity
// S + 3 n + d | pressure 4 S + n | sie 5 S + n), dust prim [4 S][nk][nj][ni]
struct StratPack {
  Real *gasp = nullptr, *dustp = nullptr;
  int Sg = 0, Sd = 0, nk = 1, nj = 1, ni = 1;
  Real &operator()(int, const VarTag &v, int k, int j, int i) const {
    const int S = v.fluid == 0 ? Sg : Sd;
    const int l = v.kind == 0 ? v.idx : (v.kind == 1 ? S + v.idx : 5 * S + v.idx);
    Real *base = v.fluid == 0 ? gasp : dustp;
    return base[(((std::size_t)l * nk + k) * nj + j) * ni + i];
  }
  int GetSize(int, const VarTag &v) const { return v.fluid == 0 ? Sg : Sd; }
};

struct StratPkg {
  bool do_dust = false;
  const strat::StratParams *pars = nullptr;
  template <class T>
  const T &Param(const std::string &name) const;
};
struct StratPackages {
  StratPkg pkg;
  const StratPkg *Get(const std::string &) const { return &pkg; }
};
struct MeshRefinementMock {
  parthenon::Coordinates_t coarse;
  const parthenon::Coordinates_t &GetCoarseCoords() const { return coarse; }
};
struct MeshBlock {
  parthenon::IndexShape cellbounds, c_cellbounds, f_cellbounds;
  parthenon::Coordinates_t coords;
  MeshRefinementMock mr, *pmr = &mr;
  StratPackages packages;
  // P:mesh/meshblock.hpp:257-267
  template <typename F>
  void par_for_bndry(const std::string &, const IndexRange &nb, const IndexDomain &domain,
                     parthenon::TE el, const bool coarse, const bool fine, const F &f) {
    auto &bounds = fine ? (coarse ? cellbounds : f_cellbounds) : (coarse ? c_cellbounds : cellbounds);
    auto ib = bounds.GetBoundsI(domain, el);
    auto jb = bounds.GetBoundsJ(domain, el);
    auto kb = bounds.GetBoundsK(domain, el);
    for (int l = nb.s; l <= nb.e; ++l)
      for (int k = kb.s; k <= kb.e; ++k)
        for (int j = jb.s; j <= jb.e; ++j)
          for (int i = ib.s; i <= ib.e; ++i) f(l, k, j, i);
  }
};
template <class T>
struct MeshBlockData {
  MeshBlock *pmb = nullptr;
  StratPack fine, coarse;
  MeshBlock *GetBlockPointer() const { return pmb; }
};

namespace ArtemisUtils {
inline int VI(const int n, const int d) { return n * 3 + d; }  // artemis_utils.hpp:26
struct StratDescriptor {
  bool coarse;
  StratPack GetPack(MeshBlockData<Real> *rc) const { return coarse ? rc->coarse : rc->fine; }
};
struct StratDescriptorMap {
  StratDescriptor operator[](bool coarse) const { return StratDescriptor{coarse}; }
};
template <class... var_ts>
StratDescriptorMap GetBoundaryPackDescriptorMap(std::shared_ptr<MeshBlockData<Real>> &) {
  return StratDescriptorMap{};
}
}  // namespace ArtemisUtils
using ArtemisUtils::VI;

namespace geometry {
struct BBox {
  Real x1[2], x2[2], x3[2];
};
template <Coordinates GEOM>
struct Coords {
  BBox bnds;
  Coords(const parthenon::Coordinates_t &c, int k, int j, int i) {
    bnds.x1[0] = c.Xf<1>(i); bnds.x1[1] = c.Xf<1>(i + 1);
    bnds.x2[0] = c.Xf<2>(j); bnds.x2[1] = c.Xf<2>(j + 1);
    bnds.x3[0] = c.Xf<3>(k); bnds.x3[1] = c.Xf<3>(k + 1);
  }
  Real x1v() const { return 0.5 * (bnds.x1[0] + bnds.x1[1]); }
  Real x2v() const { return 0.5 * (bnds.x2[0] + bnds.x2[1]); }
  Real x3v() const { return 0.5 * (bnds.x3[0] + bnds.x3[1]); }
};
}  // namespace geometry
