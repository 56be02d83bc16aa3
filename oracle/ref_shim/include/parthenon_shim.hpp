// parthenon_shim.hpp -- a MOCK of the slice of Parthenon/Kokkos that the Artemis hot-path
// sources use (TEST INFRASTRUCTURE ONLY; written for this repo, it contains no reference code).
//
// Purpose: compile the reference's OWN hot-path sources, unmodified and where they lie under
// /root/reference/src (utils/fluxes/*, geometry/*, utils/integrators/artemis_integrator.hpp,
// derived/fill_derived.cpp, rotating_frame/rotating_frame.hpp, utils/artemis_utils.hpp), into
// oracle/_ref/libartemis_ref.so without Parthenon, Kokkos, singularity-eos or cmake.  The mock
// supplies: Real, the KOKKOS_*/PARTHENON_* macros, par_for / par_for_outer / par_for_inner as
// plain (OpenMP) loops, ScratchPad2D, a UniformCartesian with Parthenon's Xf arithmetic
// (P:coordinates/uniform_cartesian.hpp:30-36,147-152), MeshData / Mesh / StateDescriptor
// lookalikes holding borrowed numpy arrays, and SparsePack lookalikes with the index-based
// (`pack(b, n, k, j, i)`, `pack.flux(b, dir, n, k, j, i)`), face (`pack(b, TE::F1, n, ...)`) and
// type-based (`pack(b, gas::cons::density(n), k, j, i)`) accessors.
#pragma once
#include <algorithm>
#include <any>
#include <array>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <limits>
#include <map>
#include <memory>
#include <set>
#include <sstream>
#include <string>
#include <tuple>
#include <typeindex>
#include <unordered_map>
#include <vector>
#ifdef _OPENMP
#include <omp.h>
#endif

using Real = double;

#define KOKKOS_INLINE_FUNCTION inline
#define KOKKOS_FORCEINLINE_FUNCTION inline
#define KOKKOS_FUNCTION
#define KOKKOS_LAMBDA [=]
#define SQR(x) ((x) * (x))

[[noreturn]] inline void shim_fail(const char *msg, const char *file, int line) {
  std::fprintf(stderr, "### ref_shim FAIL: %s (%s:%d)\n", msg, file, line);
  std::abort();
}
inline void shim_fail(const std::stringstream &msg, const char *file, int line) {
  shim_fail(msg.str().c_str(), file, line);
}
#define PARTHENON_FAIL(msg) shim_fail(msg, __FILE__, __LINE__)
#define PARTHENON_REQUIRE(cond, msg)                                                       \
  do {                                                                                     \
    if (!(cond)) shim_fail(msg, __FILE__, __LINE__);                                       \
  } while (0)

#define PARTHENON_DEBUG_REQUIRE(cond, msg) ((void)0)  // Release build: compiled out
#define PARTHENON_AUTO_LABEL "shim"

// loop-pattern tags (only passed through)
struct shim_loop_tag {};
static constexpr shim_loop_tag DEFAULT_LOOP_PATTERN{}, DEFAULT_OUTER_LOOP_PATTERN{},
    DEFAULT_INNER_LOOP_PATTERN{};

namespace Kokkos {
template <class T>
struct Min {  // reducer handle of par_reduce
  T &ref;
  explicit Min(T &r) : ref(r) {}
};
struct MemoryUnmanaged {};
template <class...>
struct View {
  template <class... A>
  View(A &&...) {}
};
}  // namespace Kokkos

namespace parthenon {

enum CoordinateDirection { NODIR = -1, X0DIR = 0, X1DIR = 1, X2DIR = 2, X3DIR = 3 };
enum class TaskStatus { complete, incomplete, iterate, fail };
enum class AmrTag : int { derefine = -1, same = 0, refine = 1 };
enum class TopologicalElement : std::size_t { CC = 0, F1 = 3, F2 = 4, F3 = 5 };
enum class IndexDomain { entire, interior };
struct IndexRange { int s = 0, e = 0; };
struct DevExecSpace {};
namespace Globals { inline int nghost = 0; }

// ---- coordinates -------------------------------------------------------------------------
struct UniformCartesian {
  std::array<Real, 3> xmin_{}, dx_{};
  template <int dir>
  Real Xf(const int idx) const { return xmin_[dir - 1] + idx * dx_[dir - 1]; }
};
using Coordinates_t = UniformCartesian;

// ---- multilevel operators (src/utils/refinement/*.hpp) ------------------------------------
// SIGN: P:config.hpp.in:86.  ParArrayND<Real, VariableState>: the 7-index accessor the
// prolongation / restriction stencils use, over ONE borrowed [nvar][nk][nj][ni] array
// (cell-centred fields only: element index, l and m are always 0).
#define SIGN(x) (((x) < 0.0) ? -1.0 : 1.0)
using TE = TopologicalElement;
struct VariableState {};
template <class T, class State = void>
struct ParArrayND {
  T *data = nullptr;
  int nvar = 0, nk = 0, nj = 0, ni = 0;
  T &operator()(int /*el*/, int /*l*/, int /*m*/, int n, int k, int j, int i) const {
    return data[(((size_t)n * nk + k) * nj + j) * ni + i];
  }
};

// ---- metadata flags / pack options ---------------------------------------------------------
enum class MetadataFlag { Conserved, WithFluxes, FillGhost };
struct Metadata {
  static constexpr MetadataFlag Conserved = MetadataFlag::Conserved;
  static constexpr MetadataFlag WithFluxes = MetadataFlag::WithFluxes;
  static constexpr MetadataFlag FillGhost = MetadataFlag::FillGhost;
};
enum class PDOpt { WithFluxes, Coarse, Flatten };

// ---- variable-name types ---------------------------------------------------------------------
namespace variable_names {
template <bool REGEX>
struct base_t {
  int idx = 0;
  base_t() = default;
  explicit base_t(int i) : idx(i) {}
};
struct any : public base_t<true> {
  using base_t<true>::base_t;
  static std::string name() { return ".*"; }
};
}  // namespace variable_names

// ---- state containers ------------------------------------------------------------------------
// One field = `ncomp` dense [nk][nj][ni] arrays per block (component c of block b at
// data + (b*ncomp + c)*cells); flux[d] likewise; face[d] are the three elements of a
// Face field with their own dims.
struct Field {
  std::string name;
  int ncomp = 0;
  bool conserved = false;
  // component c of block b lives at ptr + b*bstride + c*cells (bstride in elements), so a
  // field can be a slice of a larger [nb][nvar][nk][nj][ni] slab
  Real *data = nullptr;
  size_t bstride = 0;
  Real *flux[3] = {nullptr, nullptr, nullptr};
  size_t flux_bstride = 0;
  Real *face[3] = {nullptr, nullptr, nullptr};
  size_t face_bstride = 0;
};

// small integer id per variable-name type (O(1) type-based pack access)
inline int &shim_type_counter() {
  static int n = 0;
  return n;
}
template <class V>
inline int shim_type_id() {
  static const int id = shim_type_counter()++;
  return id;
}
constexpr int kShimMaxTypes = 64;

// Params: the type-erased key/value store of a package (P:interface/params.hpp)
class Params {
 public:
  template <class T>
  void Add(const std::string &key, T v) { params_[key] = std::move(v); }
  template <class T>
  const T &Get(const std::string &key) const {
    auto it = params_.find(key);
    if (it == params_.end()) shim_fail(("missing param " + key).c_str(), __FILE__, __LINE__);
    const T *p = std::any_cast<T>(&it->second);
    if (!p) shim_fail(("param type mismatch " + key).c_str(), __FILE__, __LINE__);
    return *p;
  }

 private:
  std::map<std::string, std::any> params_;
};

class StateDescriptor {
 public:
  StateDescriptor() = default;
  explicit StateDescriptor(const std::string &) {}
  template <class T>
  const T &Param(const std::string &key) const { return params_.Get<T>(key); }
  template <class T>
  void AddParam(const std::string &key, T v) { params_.Add(key, std::move(v)); }
  Params &AllParams() { return params_; }

 private:
  Params params_;
};

// 1-D device array with a host mirror (here both are host memory)
template <class T>
class ParArray1D {
 public:
  ParArray1D() = default;
  ParArray1D(const std::string &, int n) : d_(std::make_shared<std::vector<T>>(n)) {}
  ParArray1D GetHostMirror() const {
    ParArray1D m;
    m.d_ = std::make_shared<std::vector<T>>(d_->size());
    return m;
  }
  void DeepCopy(const ParArray1D &src) { *d_ = *src.d_; }
  T &operator()(int i) const { return (*d_)[i]; }

 private:
  std::shared_ptr<std::vector<T>> d_;
};

struct Packages_t {
  std::map<std::string, std::shared_ptr<StateDescriptor>> pkgs;
  std::shared_ptr<StateDescriptor> &Get(const std::string &n) {
    auto it = pkgs.find(n);
    if (it == pkgs.end()) shim_fail(("missing package " + n).c_str(), __FILE__, __LINE__);
    return it->second;
  }
  const std::shared_ptr<StateDescriptor> &Get(const std::string &n) const {
    auto it = pkgs.find(n);
    if (it == pkgs.end()) shim_fail(("missing package " + n).c_str(), __FILE__, __LINE__);
    return it->second;
  }
};

class Mesh {
 public:
  int ndim = 3;
  Packages_t packages;
  std::shared_ptr<StateDescriptor> resolved_packages = std::make_shared<StateDescriptor>();
};

// ParameterInput: "block/name" -> text, filled by the harness (the reference's Initialize
// functions read it exactly as they read a parsed input deck)
class ParameterInput {
 public:
  std::map<std::string, std::string> kv;
  void Set(const std::string &block, const std::string &name, const std::string &v) {
    kv[block + "/" + name] = v;
  }
  void Set(const std::string &block, const std::string &name, Real v) {
    std::ostringstream o;
    o.precision(17);
    o << v;
    kv[block + "/" + name] = o.str();
  }
  bool DoesBlockExist(const std::string &block) const {
    const std::string pre = block + "/";
    auto it = kv.lower_bound(pre);
    return it != kv.end() && it->first.compare(0, pre.size(), pre) == 0;
  }
  bool Has(const std::string &b, const std::string &n) const { return kv.count(b + "/" + n) != 0; }
  std::string GetString(const std::string &b, const std::string &n) {
    auto it = kv.find(b + "/" + n);
    if (it == kv.end()) shim_fail(("missing input " + b + "/" + n).c_str(), __FILE__, __LINE__);
    return it->second;
  }
  std::string GetOrAddString(const std::string &b, const std::string &n, const std::string &d) {
    return Has(b, n) ? kv[b + "/" + n] : d;
  }
  Real GetReal(const std::string &b, const std::string &n) { return std::stod(GetString(b, n)); }
  Real GetOrAddReal(const std::string &b, const std::string &n, Real d) {
    return Has(b, n) ? std::stod(kv[b + "/" + n]) : d;
  }
  int GetOrAddInteger(const std::string &b, const std::string &n, int d) {
    return Has(b, n) ? std::stoi(kv[b + "/" + n]) : d;
  }
  bool GetOrAddBoolean(const std::string &b, const std::string &n, bool d) {
    return Has(b, n) ? (kv[b + "/" + n] == "true" || kv[b + "/" + n] == "1") : d;
  }
  template <class T>
  std::vector<T> GetVector(const std::string &b, const std::string &n) {
    std::vector<T> out;
    std::stringstream ss(GetString(b, n));
    std::string tok;
    while (std::getline(ss, tok, ',')) out.push_back((T)std::stod(tok));
    return out;
  }
};

template <class T>
class MeshData {
 public:
  Mesh *pm = nullptr;
  int nb = 0, ni = 0, nj = 0, nk = 0, fni = 0, fnj = 0, fnk = 0;
  IndexRange ib, jb, kb;  // interior
  std::vector<UniformCartesian> coords;
  std::vector<Field> fields;

  Mesh *GetParentPointer() const { return pm; }
  int NumBlocks() const { return nb; }
  IndexRange GetBoundsI(IndexDomain d) const { return d == IndexDomain::interior ? ib : IndexRange{0, ni - 1}; }
  IndexRange GetBoundsJ(IndexDomain d) const { return d == IndexDomain::interior ? jb : IndexRange{0, nj - 1}; }
  IndexRange GetBoundsK(IndexDomain d) const { return d == IndexDomain::interior ? kb : IndexRange{0, nk - 1}; }
  const Field *Find(const std::string &n) const {
    for (auto &f : fields)
      if (f.name == n) return &f;
    return nullptr;
  }
};
template <class T>
class MeshBlockData : public MeshData<T> {};

struct MeshBlockDataCollection {
  std::shared_ptr<MeshBlockData<Real>> base = std::make_shared<MeshBlockData<Real>>();
  std::shared_ptr<MeshBlockData<Real>> &Get() { return base; }
};
class MeshBlock {
 public:
  MeshBlockDataCollection meshblock_data;
};

// ---- packs ---------------------------------------------------------------------------------
// Flattened view over the selected fields of one MeshData, in the order the variable types
// were listed (Parthenon: SparsePack index = position in the descriptor, components of a
// vector field consecutive).
class SparsePackShim {
 public:
  int nb = 0, nvar = 0, nj = 0, ni = 0, fnj = 0, fni = 0;
  size_t cells = 0, fcells = 0;
  const UniformCartesian *coords = nullptr;
  struct Entry {
    Real *data, *flux[3], *face[3];
    size_t bstride, flux_bstride, face_bstride;
    int comp;
  };
  std::vector<Entry> ent;  // [nvar]
  int type_off[kShimMaxTypes], type_size[kShimMaxTypes];
  SparsePackShim() {
    for (int t = 0; t < kShimMaxTypes; ++t) { type_off[t] = -1; type_size[t] = 0; }
  }

  int GetNBlocks() const { return nb; }
  int GetLowerBound(int) const { return 0; }
  int GetUpperBound(int) const { return nvar - 1; }
  int GetMaxNumberOfVars() const { return nvar; }
  const Coordinates_t &GetCoordinates(int b) const { return coords[b]; }

  Real &operator()(int b, int n, int k, int j, int i) const {
    const Entry &e = ent[n];
    return e.data[b * e.bstride + e.comp * cells + ((size_t)k * nj + j) * ni + i];
  }
  Real &flux(int b, int dir, int n, int k, int j, int i) const {
    const Entry &e = ent[n];
    return e.flux[dir - 1][b * e.flux_bstride + e.comp * cells + ((size_t)k * nj + j) * ni + i];
  }
  Real &operator()(int b, TopologicalElement el, int n, int k, int j, int i) const {
    const Entry &e = ent[n];
    const int d = (int)el - (int)TopologicalElement::F1;
    return e.face[d][b * e.face_bstride + e.comp * fcells + ((size_t)k * fnj + j) * fni + i];
  }
  template <class V, class = decltype(V::name())>
  Real &operator()(int b, const V &v, int k, int j, int i) const {
    return (*this)(b, type_off[shim_type_id<V>()] + v.idx, k, j, i);
  }
  struct VarHandle {
    int sparse_id;
  };
  template <class V, class = decltype(V::name())>
  VarHandle operator()(int, const V &v) const {
    return VarHandle{v.idx};  // sparse pools are registered with ids 0..nspecies-1
  }
  template <class V, class = decltype(V::name())>
  Real &operator()(int b, TopologicalElement el, const V &v, int k, int j, int i) const {
    return (*this)(b, el, type_off[shim_type_id<V>()] + v.idx, k, j, i);
  }
  template <class V, class = decltype(V::name())>
  Real &flux(int b, int dir, const V &v, int k, int j, int i) const {
    return flux(b, dir, type_off[shim_type_id<V>()] + v.idx, k, j, i);
  }
  template <class V, class = decltype(V::name())>
  int GetSize(int, const V &) const {
    return type_size[shim_type_id<V>()];
  }
};
template <class... Ts>
struct SparsePack : public SparsePackShim {
  struct Descriptor;
};

template <class... Ts>
struct PackDescriptorShim {
  std::vector<MetadataFlag> flags;
  bool any_ = false;

  SparsePackShim GetPack(const MeshData<Real> *md) const {
    SparsePackShim p;
    p.nb = md->nb; p.nj = md->nj; p.ni = md->ni; p.fnj = md->fnj; p.fni = md->fni;
    p.cells = (size_t)md->nk * md->nj * md->ni;
    p.fcells = (size_t)md->fnk * md->fnj * md->fni;
    p.coords = md->coords.data();
    auto add_field = [&](const Field &f, int tid) {
      if (tid >= 0) {
        if (tid >= kShimMaxTypes) shim_fail("too many variable types", __FILE__, __LINE__);
        p.type_off[tid] = (int)p.ent.size();
        p.type_size[tid] = f.ncomp;
      }
      for (int c = 0; c < f.ncomp; ++c)
        p.ent.push_back({f.data, {f.flux[0], f.flux[1], f.flux[2]},
                         {f.face[0], f.face[1], f.face[2]}, f.bstride, f.flux_bstride,
                         f.face_bstride, c});
    };
    if (any_) {
      for (auto &f : md->fields) {
        bool ok = true;
        for (auto fl : flags)
          if (fl == MetadataFlag::Conserved && !f.conserved) ok = false;
        if (ok) add_field(f, -1);
      }
    } else {
      (AddNamed<Ts>(md, add_field), ...);
    }
    p.nvar = (int)p.ent.size();
    return p;
  }

 private:
  template <class V, class F>
  static void AddNamed(const MeshData<Real> *md, F &add_field) {
    const Field *f = md->Find(V::name());
    if (f) add_field(*f, shim_type_id<V>());  // absent (sparse, unallocated) field: size 0
  }
};

template <class... Ts, class P>
PackDescriptorShim<Ts...> MakePackDescriptor(P *, const std::vector<MetadataFlag> &flags = {},
                                             const std::set<PDOpt> & = {}) {
  PackDescriptorShim<Ts...> d;
  d.flags = flags;
  d.any_ = (std::is_same_v<Ts, variable_names::any> || ...);
  return d;
}

// ---- scratch + loops --------------------------------------------------------------------------
struct ScratchArena {
  Real *base = nullptr;
  size_t used = 0;
};
struct team_mbr_t {
  mutable ScratchArena arena;
  ScratchArena &team_scratch(int) const { return arena; }
  void team_barrier() const {}
};
template <class T>
class ScratchPad2D {
 public:
  ScratchPad2D() = default;
  ScratchPad2D(ScratchArena &a, int n, int m) : p_(a.base + a.used), m_(m) { a.used += (size_t)n * m; }
  T &operator()(int n, int i) const { return p_[(size_t)n * m_ + i]; }
  static int shmem_size(int n, int m) { return (int)sizeof(T) * n * m; }

 private:
  T *p_ = nullptr;
  int m_ = 0;
};
template <class T>
class ScratchPad1D {
 public:
  ScratchPad1D() = default;
  ScratchPad1D(ScratchArena &a, int n) : p_(a.base + a.used) { a.used += (size_t)n; }
  T &operator()(int i) const { return p_[i]; }
  static int shmem_size(int n) { return (int)sizeof(T) * n; }

 private:
  T *p_ = nullptr;
};

template <class F>
inline void par_for_inner(shim_loop_tag, const team_mbr_t &, int il, int iu, const F &f) {
  for (int i = il; i <= iu; ++i) f(i);
}
// (b, k, j) teams
template <class F>
inline void par_for_outer(shim_loop_tag, const char *, DevExecSpace, int scr_bytes, int, int b0,
                          int b1, int k0, int k1, int j0, int j1, const F &f) {
#pragma omp parallel
  {
    std::vector<Real> buf(scr_bytes / sizeof(Real) + 8);
#pragma omp for collapse(3) schedule(static)
    for (int b = b0; b <= b1; ++b)
      for (int k = k0; k <= k1; ++k)
        for (int j = j0; j <= j1; ++j) {
          team_mbr_t m;
          m.arena.base = buf.data();
          m.arena.used = 0;
          f(m, b, k, j);
        }
  }
}
// (b, k) teams
template <class F>
inline void par_for_outer(shim_loop_tag, const char *, DevExecSpace, int scr_bytes, int, int b0,
                          int b1, int k0, int k1, const F &f) {
#pragma omp parallel
  {
    std::vector<Real> buf(scr_bytes / sizeof(Real) + 8);
#pragma omp for collapse(2) schedule(static)
    for (int b = b0; b <= b1; ++b)
      for (int k = k0; k <= k1; ++k) {
        team_mbr_t m;
        m.arena.base = buf.data();
        m.arena.used = 0;
        f(m, b, k);
      }
  }
}
template <class F>
inline void par_for(shim_loop_tag, const char *, DevExecSpace, int b0, int b1, int k0, int k1,
                    int j0, int j1, int i0, int i1, const F &f) {
#pragma omp parallel for collapse(3) schedule(static)
  for (int b = b0; b <= b1; ++b)
    for (int k = k0; k <= k1; ++k)
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) f(b, k, j, i);
}

// par_reduce with a Min reducer over (b, k, j, i): min is order-independent, so the OpenMP
// reduction returns the same bits as any other traversal
struct loop_pattern_mdrange_tag_t {};
static constexpr loop_pattern_mdrange_tag_t loop_pattern_mdrange_tag{};
template <class F>
inline void par_reduce(loop_pattern_mdrange_tag_t, const char *, DevExecSpace, int b0, int b1,
                       int k0, int k1, int j0, int j1, int i0, int i1, const F &f,
                       Kokkos::Min<Real> red) {
  Real m = std::numeric_limits<Real>::max();
#pragma omp parallel for collapse(3) schedule(static) reduction(min : m)
  for (int b = b0; b <= b1; ++b)
    for (int k = k0; k <= k1; ++k)
      for (int j = j0; j <= j1; ++j)
        for (int i = i0; i <= i1; ++i) f(b, k, j, i, m);
  red.ref = m;
}

class LowStorageIntegrator {
 public:
  Real dt = 0.0;
  std::vector<Real> gam0, gam1, beta;
};

namespace driver { namespace prelude {} }
namespace package { namespace prelude {} }
}  // namespace parthenon
