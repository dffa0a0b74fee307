/* oracle/shim/boost/filesystem/operations.hpp -- TEST INFRASTRUCTURE ONLY (see path.hpp). */
#pragma once
#include "path.hpp"
