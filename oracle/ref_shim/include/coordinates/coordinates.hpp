// MOCK of <coordinates/coordinates.hpp>: see parthenon_shim.hpp
#pragma once
#include "parthenon_shim.hpp"
