// compat shim: the reference includes "database/database.hpp" (src/common/database/database.hpp); only database::airfoil is needed by rans
#pragma once
#include "../rans/common.h"
