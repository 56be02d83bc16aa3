// bnd_mock.hpp -- the handful of Parthenon types that parthenon::CalcIndices
// (external/parthenon/src/bvals/comms/bnd_info.cpp:105-252) touches, mocked so that the
// function can be compiled ON ITS OWN from the reference tree: build_bnd_ref.py slices its text
// out of bnd_info.cpp at build time into oracle/_ref/ (never committed) and wraps it with the C
// entry point at the bottom of the generated file.  TEST INFRASTRUCTURE ONLY.
#pragma once
#include <array>
#include <cstddef>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <memory>

#define PARTHENON_FAIL(msg) (std::fprintf(stderr, "%s\n", msg), std::abort())

namespace parthenon {
using Real = double;
enum CoordinateDirection { NODIR = -1, X0DIR = 0, X1DIR = 1, X2DIR = 2, X3DIR = 3 };
enum class TopologicalElement : std::size_t { CC = 0, F1 = 3, F2 = 4, F3 = 5 };
// P:basic_types.hpp: a face element adds one index along its own direction
inline int TopologicalOffsetI(TopologicalElement el) { return el == TopologicalElement::F1; }
inline int TopologicalOffsetJ(TopologicalElement el) { return el == TopologicalElement::F2; }
inline int TopologicalOffsetK(TopologicalElement el) { return el == TopologicalElement::F3; }
enum class IndexDomain { entire, interior };
enum class IndexRangeType { BoundaryInteriorSend, BoundaryExteriorRecv, InteriorSend, InteriorRecv };
struct IndexRange { int s = 0, e = 0; };
namespace Globals { inline int nghost = 2; }

// P:mesh/domain.hpp: interior [ng, ng + nx - 1] in directions with nx > 1, [0, 0] otherwise
class IndexShape {
 public:
  IndexShape() = default;
  IndexShape(int nx3, int nx2, int nx1, int ng) : n_{nx1, nx2, nx3}, ng_(ng) {}
  IndexRange GetBoundsI(IndexDomain, TopologicalElement el) const { return b(0, TopologicalOffsetI(el)); }
  IndexRange GetBoundsJ(IndexDomain, TopologicalElement el) const { return b(1, TopologicalOffsetJ(el)); }
  IndexRange GetBoundsK(IndexDomain, TopologicalElement el) const { return b(2, TopologicalOffsetK(el)); }

 private:
  IndexRange b(int d, int top) const {  // P:mesh/domain.hpp:296-318
    const int g = n_[d] > 1 ? ng_ : 0;
    return IndexRange{g, g + n_[d] - 1 + top};
  }
  int n_[3] = {1, 1, 1}, ng_ = 0;
};

class LogicalLocation {
 public:
  LogicalLocation() = default;
  LogicalLocation(int level, std::int64_t l1, std::int64_t l2, std::int64_t l3)
      : level_(level), l_{l1, l2, l3} {}
  int level() const { return level_; }
  std::int64_t l(int d) const { return l_[d]; }

 private:
  int level_ = 0;
  std::int64_t l_[3] = {0, 0, 0};
};

struct RegionSize {
  int n[3] = {1, 1, 1};
  bool sym[3] = {false, false, false};
  int nx(CoordinateDirection d) const { return n[d - 1]; }
  bool symmetry(CoordinateDirection d) const { return sym[d - 1]; }
};

struct block_ownership_t {
  explicit block_ownership_t(bool) {}
  block_ownership_t() = default;
};
inline block_ownership_t GetIndexRangeMaskFromOwnership(TopologicalElement, const block_ownership_t &,
                                                        int, int, int) {
  return block_ownership_t(true);  // cell-centred data is always owned
}

namespace forest {
struct LogicalCoordinateTransformation {
  std::array<int, 3> Transform(const std::array<int, 3> &a) const { return a; }  // one tree
};
}  // namespace forest

struct NeighborBlock {
  LogicalLocation loc, origin_loc;
  RegionSize block_size;
  std::array<int, 3> offsets{0, 0, 0};
  block_ownership_t ownership;
};

struct MeshBlock {
  LogicalLocation loc;
  IndexShape cellbounds, c_cellbounds, f_cellbounds;
  RegionSize block_size;
};

struct Metadata {
  enum Flag { Flux, Fine };
};
template <class T>
struct Variable {
  bool flux = false;
  int GetDim(int) const { return 1; }
  bool IsSet(Metadata::Flag f) const { return f == Metadata::Flux && flux; }
};

struct SpatiallyMaskedIndexer6D {
  int s[3], e[3];  // i, j, k
  SpatiallyMaskedIndexer6D(block_ownership_t, IndexRange, IndexRange, IndexRange, IndexRange k,
                           IndexRange j, IndexRange i)
      : s{i.s, j.s, k.s}, e{i.e, j.e, k.e} {}
};
}  // namespace parthenon
