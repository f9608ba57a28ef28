// util.hpp -- name-compatibility shim: reference code does #include "util.hpp" for the boundary types.
#pragma once
#include "scrooge_types.hpp"
