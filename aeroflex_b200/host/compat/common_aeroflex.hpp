// compat shim: the reference includes "common_aeroflex.hpp" (src/common/common_aeroflex.hpp)
#pragma once
#include "rans/common.h"
