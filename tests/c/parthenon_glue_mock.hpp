// parthenon_glue_mock.hpp -- declarations (no bodies beyond trivial ones) of EXACTLY the
// Parthenon / Artemis names tests/c/ab200_glue.cpp uses, each with the signature it has in the
// reference and the file:line it is declared at (lanl/artemis @ 6c2a7a8, P: =
// external/parthenon/src/).  It exists so the glue TU is compiled (g++ -fsyntax-only) on every
// test run instead of rotting as pseudo-code; it is test infrastructure, not a Parthenon port.
#pragma once
#include <map>
#include <memory>
#include <string>
#include <vector>

using Real = double;  // P:basic_types.hpp:29-35

namespace parthenon {
enum class TaskStatus { complete, incomplete, iterate, fail };   // P:basic_types.hpp:59
enum class IndexDomain { entire, interior };                     // P:mesh/domain.hpp
struct IndexRange { int s = 0, e = 0; };                         // P:mesh/domain.hpp:38-48
constexpr int X1DIR = 1, X2DIR = 2, X3DIR = 3;                   // P:defs.hpp

// P:parthenon_array_generic.hpp:209 (data), :150-160 (GetDim)
template <class T>
struct ParArrayND {
  T *ptr = nullptr;
  int dim[7] = {1, 1, 1, 1, 1, 1, 1};
  T *data() const { return ptr; }
  int GetDim(int i) const { return dim[i - 1]; }
};
// P:interface/variable.hpp:138: `ParArrayND<T, VariableState> data;`
// :139: `ParArrayND<T, VariableState> coarse_s;` (the coarse buffer of the multilevel exchange)
template <class T>
struct Variable {
  ParArrayND<T> data;
  ParArrayND<T> coarse_s;
};
// P:coordinates/uniform_cartesian.hpp:84-88 (Dxf), :117-126 (Xf: xmin_[dir-1] + idx * dx_[dir-1])
struct UniformCartesian {
  Real xmin_[3] = {0, 0, 0}, dx_[3] = {1, 1, 1};
  template <int dir, class... Args>
  Real Dxf(Args...) const { return dx_[dir - 1]; }
  template <int dir, int face>
  Real Xf(const int idx) const { return xmin_[dir - 1] + idx * dx_[dir - 1]; }
};
// P:mesh/domain.hpp:183-290
struct IndexShape {
  int is(IndexDomain) const { return 0; }
  int ie(IndexDomain) const { return 0; }
  int js(IndexDomain) const { return 0; }
  int je(IndexDomain) const { return 0; }
  int ks(IndexDomain) const { return 0; }
  int ke(IndexDomain) const { return 0; }
  int ncellsi(IndexDomain) const { return 1; }
  int ncellsj(IndexDomain) const { return 1; }
  int ncellsk(IndexDomain) const { return 1; }
};
// P:mesh/meshblock.hpp:122 (cellbounds), coords, gid
struct MeshBlock {
  IndexShape cellbounds;
  UniformCartesian coords;
  int gid = 0;
};
// P:interface/params.hpp: Params::Get<T>(key); P:interface/state_descriptor.hpp: Param<T>(key)
struct StateDescriptor {
  template <class T>
  const T &Param(const std::string &) const { static T v{}; return v; }
};
struct Packages_t {                                              // P:interface/packages.hpp
  std::shared_ptr<StateDescriptor> &Get(const std::string &) { static std::shared_ptr<StateDescriptor> p = std::make_shared<StateDescriptor>(); return p; }
};
struct Mesh {                                                    // P:mesh/mesh.hpp
  int ndim = 3;
  Packages_t packages;
};
// P:interface/meshblock_data.hpp:75 (GetBlockPointer), :262 (Get(base_name, sparse_id))
template <class T>
struct MeshBlockData {
  MeshBlock *GetBlockPointer() const { return nullptr; }
  Variable<T> &Get(const std::string &, int = -1) const { static Variable<T> v; return v; }
};
// P:interface/mesh_data.hpp:203-204 (GetMeshPointer / GetParentPointer), :282 (GetBlockData),
// :454 (NumBlocks)
template <class T>
struct MeshData {
  Mesh *GetParentPointer() const { return nullptr; }
  int NumBlocks() const { return 0; }
  const std::shared_ptr<MeshBlockData<T>> &GetBlockData(int) const { static std::shared_ptr<MeshBlockData<T>> p; return p; }
};
// P:time_integration/staged_integrator.hpp: gam0, gam1, beta, dt, nstages
struct LowStorageIntegrator {
  int nstages = 2;
  Real dt = 0;
  std::vector<Real> gam0, gam1, beta;
};
namespace Globals { extern int nghost; }                         // P:globals.hpp
}  // namespace parthenon

// src/artemis.hpp:78-105
enum class Coordinates { cartesian, cylindrical, spherical1D, spherical2D, spherical3D, axisymmetric, null };
enum class RSolver { hllc, hlle, llf, null };
enum class ReconstructionMethod { pcm, plm, ppm, null };
#define PARTHENON_FAIL(msg) throw std::runtime_error(msg)        // P:utils/error_checking.hpp
#include <stdexcept>
