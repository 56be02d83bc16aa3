// MOCK: multilevel prolongation is outside what oracle/_ref compiles (uniform meshes only).
// The real header is what brings geometry/geometry.hpp into utils/artemis_utils.hpp.
#pragma once
#include "geometry/geometry.hpp"
