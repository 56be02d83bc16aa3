// MOCK of <interface/variable_state.hpp>: see parthenon_shim.hpp
#pragma once
#include "parthenon_shim.hpp"
