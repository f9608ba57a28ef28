// sim_runtime.h -- launch interface of the host simulation (tests/sim).  TEST INFRASTRUCTURE.
#pragma once
#include <cuda_runtime.h>   // the stand-in in tests/sim/shim

namespace sim {
// Runs body(arg) once per thread of a grid x block launch, CTA after CTA; smem = the array the kernels' `extern __shared__`
// declaration names (every CTA sees it as its dynamic shared memory).
void launch(unsigned grid, unsigned block, void *smem, void (*body)(void *), void *arg);
// 0 (default): fibers run in thread order; otherwise every scheduling sweep uses a fresh pseudo-random order derived from it
extern uint64_t schedule_seed;
}
