// bc_mock.hpp -- the few Parthenon names parthenon::BoundaryFunction::GenericBC
// (external/parthenon/src/bvals/boundary_conditions_generic.hpp:172-246) and
// parthenon::IndexShape (external/parthenon/src/mesh/domain.hpp:83-...) touch, so that BOTH can
// be sliced out of the reference tree at build time and compiled as they are
// (oracle/ref_shim/bc/build_bc_ref.py).  TEST INFRASTRUCTURE: pins the oracle's outflow /
// reflecting boundary fill (oracle/artemis_oracle.c, ao_exchange_ghosts_phase, phase 2) to the
// reference's own code.  Nothing here is reference text.
#pragma once
#include <algorithm>
#include <array>
#include <cassert>
#include <cstddef>
#include <map>
#include <memory>
#include <string>
#include <tuple>
#include <vector>

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define PARTHENON_AUTO_LABEL std::string("bc")

namespace parthenon {
using Real = double;
constexpr int NDIM = 3;
enum CoordinateDirection { NODIR = -1, X0DIR = 0, X1DIR = 1, X2DIR = 2, X3DIR = 3 };
enum class TopologicalElement : std::size_t { CC = 0, F1 = 3, F2 = 4, F3 = 5, E1 = 6, E2 = 7, E3 = 8, NN = 9 };
using TE = TopologicalElement;
enum class TopologicalType { Cell, Face, Edge, Node };
inline TopologicalType GetTopologicalType(TopologicalElement) { return TopologicalType::Cell; }
// P:basic_types.hpp: only cell-centred fields are exchanged on this path
inline int TopologicalOffsetI(TE el) { return el == TE::F1; }
inline int TopologicalOffsetJ(TE el) { return el == TE::F2; }
inline int TopologicalOffsetK(TE el) { return el == TE::F3; }
struct IndexRange { int s = 0, e = 0; };

// ---- IndexDomain + class IndexShape: sliced from domain.hpp by the build script -------------
#include "bc_indexshape_generated.inc"

// one MeshBlock's FillGhost fields, dense [nvar][nk][nj][ni]
struct VarInfo { int vector_component = NODIR; };
struct BcPack {
  Real *data = nullptr;
  int nvar = 0, nk = 1, nj = 1, ni = 1;
  const int *vcomp = nullptr;
  bool empty = false;
  int GetLowerBoundHost(int) const { return 0; }
  int GetUpperBoundHost(int) const { return empty ? -1 : nvar - 1; }
  VarInfo operator()(int, TE, int l) const { return VarInfo{vcomp[l]}; }
  Real &operator()(int, TE, int l, int k, int j, int i) const {
    return data[(((std::size_t)l * nk + k) * nj + j) * ni + i];
  }
};

struct MeshBlock {
  IndexShape cellbounds, c_cellbounds, f_cellbounds;
  // P:mesh/meshblock.hpp:257-267
  template <typename F>
  void par_for_bndry(const std::string &, const IndexRange &nb, const IndexDomain &domain, TE el,
                     const bool coarse, const bool fine, const F &f) {
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
  BcPack pack;
  MeshBlock *GetBlockPointer() const { return pmb; }
};

namespace variable_names {
struct any {};
}  // namespace variable_names

namespace BoundaryFunction {
namespace impl {
using desc_key_t = std::tuple<bool, bool, TopologicalType>;
struct BcDescriptor {
  bool fine = false;
  BcPack GetPack(MeshBlockData<Real> *rc) const {
    BcPack p = rc->pack;
    p.empty = fine;  // no Metadata::Fine fields on this path: the `fine` pass finds an empty pack
    return p;
  }
};
struct BcDescriptorMap {
  BcDescriptor operator[](const desc_key_t &k) const { return BcDescriptor{std::get<1>(k)}; }
};
template <class... var_ts>
BcDescriptorMap GetPackDescriptorMap(std::shared_ptr<MeshBlockData<Real>> &) {
  return BcDescriptorMap{};
}
}  // namespace impl
}  // namespace BoundaryFunction
}  // namespace parthenon
