// utils/eos/eos.hpp (MOCK) -- the reference's EOS is singularity::Variant<singularity::IdealGas>
// (src/utils/eos/eos.hpp:26; singularity-eos 1.9.1, pinned 3f78b83).  singularity-eos cannot be
// compiled here (its ports-of-call/spiner submodules are not checked out), so the five
// IdealGas members the hot path and the diffusion operators call are restated from the
// published source external/singularity-eos/singularity-eos/eos/eos_ideal.hpp:60-64,90-93,
// 122-126,134-137,140-143.
#ifndef UTILS_EOS_HPP_
#define UTILS_EOS_HPP_
#include "artemis.hpp"
namespace ArtemisUtils {
static constexpr int lambda_max_vals = 1;
class EOS {
 public:
  EOS() = default;
  EOS(Real gm1, Real Cv) : _Cv(Cv), _gm1(gm1) {}
  template <class L = Real *>
  Real PressureFromDensityInternalEnergy(const Real rho, const Real sie, L && = nullptr) const {
    const Real v = _gm1 * rho * sie;
    return 0.0 > v ? 0.0 : v;  // MYMAX(0.0, v), eos_ideal.hpp:32
  }
  template <class L = Real *>
  Real BulkModulusFromDensityInternalEnergy(const Real rho, const Real sie, L && = nullptr) const {
    const Real v = (_gm1 + 1) * _gm1 * rho * sie;
    return 0.0 > v ? 0.0 : v;
  }
  template <class L = Real *>
  Real TemperatureFromDensityInternalEnergy(const Real, const Real sie, L && = nullptr) const {
    const Real v = sie / _Cv;
    return 0.0 > v ? 0.0 : v;  // MYMAX(0.0, sie / _Cv), eos_ideal.hpp:63
  }
  template <class L = Real *>
  Real SpecificHeatFromDensityInternalEnergy(const Real, const Real, L && = nullptr) const {
    return _Cv;
  }
  template <class L = Real *>
  Real GruneisenParamFromDensityInternalEnergy(const Real, const Real, L && = nullptr) const {
    return _gm1;  // eos_ideal.hpp:150-154
  }
  template <class L = Real *>
  Real GruneisenParamFromDensityTemperature(const Real, const Real, L && = nullptr) const {
    return _gm1;
  }

 private:
  Real _Cv = 1.0, _gm1 = 0.0;
};
}  // namespace ArtemisUtils
#endif
