// artemis.hpp (MOCK) -- stands in for /root/reference/src/artemis.hpp when the reference's
// hot-path sources are compiled into oracle/_ref (see parthenon_shim.hpp).  The include
// order of the build recipe makes `#include "artemis.hpp"` inside the reference sources
// resolve here.  It declares the names those sources expect from the real header: the
// variable-name types, the selection enums (same enumerator order as src/artemis.hpp:78-105,
// which is also the integer encoding of include/ab200.h) and the Big/Fuzz/Null/NewArray
// helpers.  TEST INFRASTRUCTURE ONLY.
#ifndef ARTEMIS_ARTEMIS_HPP_
#define ARTEMIS_ARTEMIS_HPP_
#include "parthenon_shim.hpp"

using namespace parthenon;

#define SHIM_VARIABLE(ns, varname)                                                         \
  struct varname : public parthenon::variable_names::base_t<false> {                       \
    using parthenon::variable_names::base_t<false>::base_t;                                \
    static std::string name() { return #ns "." #varname; }                                 \
  }
namespace gas {
namespace cons { SHIM_VARIABLE(gas.cons, density); SHIM_VARIABLE(gas.cons, total_energy);
                 SHIM_VARIABLE(gas.cons, internal_energy); SHIM_VARIABLE(gas.cons, momentum); }
namespace prim { SHIM_VARIABLE(gas.prim, density); SHIM_VARIABLE(gas.prim, pressure);
                 SHIM_VARIABLE(gas.prim, velocity); SHIM_VARIABLE(gas.prim, sie); }
namespace face { SHIM_VARIABLE(gas.face, velocity); }
namespace diff { SHIM_VARIABLE(gas.diff, momentum); SHIM_VARIABLE(gas.diff, energy); }
}  // namespace gas
namespace dust {
namespace cons { SHIM_VARIABLE(dust.cons, density); SHIM_VARIABLE(dust.cons, momentum); }
namespace prim { SHIM_VARIABLE(dust.prim, density); SHIM_VARIABLE(dust.prim, velocity); }
}  // namespace dust
#undef SHIM_VARIABLE

enum class Coordinates { cartesian, cylindrical, spherical1D, spherical2D, spherical3D,
                         axisymmetric, null };
enum class RSolver { hllc, hlle, llf, null };
enum class ReconstructionMethod { pcm, plm, ppm, null };
enum class Fluid { gas, dust, null };

template <typename T = Real>
inline constexpr auto Big() { return std::numeric_limits<T>::max(); }
template <typename T = Real>
inline constexpr auto Fuzz() {
  if constexpr (std::is_same_v<T, float>) return 1e-22;
  return 1e-99;
}
template <typename T = Real>
inline constexpr auto Null() { return std::numeric_limits<T>::quiet_NaN(); }
template <>
inline constexpr auto Null<int>() { return Big<int>(); }
template <typename T, int N>
inline auto NewArray(T val = Null<T>()) {
  std::array<T, N> arr;
  for (int i = 0; i < N; i++) arr[i] = val;
  return arr;
}
#endif  // ARTEMIS_ARTEMIS_HPP_
