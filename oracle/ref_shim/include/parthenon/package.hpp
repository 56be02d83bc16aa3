// MOCK of <parthenon/package.hpp>: see parthenon_shim.hpp
#pragma once
#include "parthenon_shim.hpp"
