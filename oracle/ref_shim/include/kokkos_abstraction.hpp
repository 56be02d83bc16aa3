// MOCK of <kokkos_abstraction.hpp>: see parthenon_shim.hpp
#pragma once
#include "parthenon_shim.hpp"
