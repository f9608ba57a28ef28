// util.hpp -- name-compatibility shim: reference code does #include "util.hpp" for the boundary types and the
// dataset readers (reference src/util.hpp).
#pragma once
#include "scrooge_types.hpp"
#include "scrooge_io.hpp"
