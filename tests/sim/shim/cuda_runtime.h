// cuda_runtime.h STAND-IN for the host simulation of the alignment kernels (tests/sim): TEST INFRASTRUCTURE, never part of
// the product.  tests/sim/sim_kernels.cpp compiles scrooge_b200/csrc/sg_align_delta.cuh (the kernel source itself, with
// -DSG_SIM) with g++ against this header: every CUDA thread of a CTA is a fiber (ucontext) of one host thread, warp
// collectives (__all_sync, __shfl_*_sync, __syncwarp) and __syncthreads are rendezvous between fibers, shared memory is a
// host array, atomics are plain read-modify-writes (one host thread).  What it checks is the kernel's LOGIC, statement
// by statement, against the oracle on the CPU-only box; what it cannot check is what the few inline-PTX blocks do on the
// hardware (each has a C++ twin under SG_SIM right beside it) and anything about timing.
#pragma once
#include <stdint.h>
#include <stddef.h>
#include <string.h>
#include <type_traits>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __shared__
#define __align__(n) __attribute__((aligned(n)))

struct uint2 { uint32_t x, y; };
struct uint3 { uint32_t x, y, z; };
struct alignas(16) uint4 { uint32_t x, y, z, w; };
static inline uint2 make_uint2(uint32_t x, uint32_t y) { return uint2{x, y}; }
static inline uint4 make_uint4(uint32_t x, uint32_t y, uint32_t z, uint32_t w) { return uint4{x, y, z, w}; }

namespace sim {
// state of the running fiber / CTA (tests/sim/sim_runtime.cpp)
struct ThreadCtx { uint3 tid; };
extern ThreadCtx *cur;
extern uint3 block_idx, block_dim, grid_dim;
extern unsigned char *smem_base;               // the CTA's dynamic shared memory
unsigned ballot(bool pred);                    // warp rendezvous: bit k = predicate of lane k
unsigned alive_mask();                         // lanes of the current warp that have not returned from the kernel
uint64_t exchange(uint64_t v, int src_lane);   // warp rendezvous: every lane reads src_lane's v
void cta_barrier();
// counters for the modelled memory traffic of a launch (what the fibers count through SG_SIM_COUNT)
extern uint64_t counters[8];
}  // namespace sim

#define threadIdx (sim::cur->tid)
#define blockIdx (sim::block_idx)
#define blockDim (sim::block_dim)
#define gridDim (sim::grid_dim)

static inline int __popc(uint32_t x) { return __builtin_popcount(x); }
static inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
static inline int __ffs(int x) { return __builtin_ffs(x); }
static inline uint32_t __brev(uint32_t x)
{
    x = ((x >> 1) & 0x55555555u) | ((x & 0x55555555u) << 1);
    x = ((x >> 2) & 0x33333333u) | ((x & 0x33333333u) << 2);
    x = ((x >> 4) & 0x0F0F0F0Fu) | ((x & 0x0F0F0F0Fu) << 4);
    return __builtin_bswap32(x);
}
// PRMT (default mode): result byte k = byte (selector nibble k & 7) of {b, a}; nibble bit 3 replicates that byte's sign
static inline uint32_t __byte_perm(uint32_t a, uint32_t b, uint32_t s)
{
    const uint64_t src = ((uint64_t)b << 32) | a;
    uint32_t r = 0;
    for (int k = 0; k < 4; k++) {
        const uint32_t sel = (s >> (4 * k)) & 0xFu;
        uint32_t byte = (uint32_t)(src >> (8 * (sel & 7u))) & 0xFFu;
        if (sel & 8u) byte = (byte & 0x80u) ? 0xFFu : 0u;
        r |= byte << (8 * k);
    }
    return r;
}
static inline uint32_t __funnelshift_l(uint32_t lo, uint32_t hi, uint32_t sh)
{
    return (uint32_t)(((((uint64_t)hi << 32) | lo) << (sh & 31u)) >> 32);
}
static inline uint32_t __funnelshift_r(uint32_t lo, uint32_t hi, uint32_t sh)
{
    return (uint32_t)((((uint64_t)hi << 32) | lo) >> (sh & 31u));
}
template <typename T> static inline T __ldg(const T *p) { return *p; }
static inline uint32_t min(uint32_t a, uint32_t b) { return a < b ? a : b; }
static inline uint32_t max(uint32_t a, uint32_t b) { return a > b ? a : b; }

static inline unsigned long long atomicAdd(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; *p = o + v; return o; }
static inline unsigned long long atomicMin(unsigned long long *p, unsigned long long v) { const unsigned long long o = *p; if (v < o) *p = v; return o; }

static inline int __all_sync(unsigned, int pred) { const unsigned v = sim::ballot(pred != 0); return (v & sim::alive_mask()) == sim::alive_mask(); }
static inline int __any_sync(unsigned, int pred) { return sim::ballot(pred != 0) != 0u; }
static inline unsigned __ballot_sync(unsigned, int pred) { return sim::ballot(pred != 0); }
static inline void __syncwarp(unsigned = 0xFFFFFFFFu) { (void)sim::ballot(true); }
static inline void __syncthreads() { sim::cta_barrier(); }
template <typename T> static inline T __shfl_sync(unsigned, T v, int src)
{
    static_assert(sizeof(T) <= 8, "shuffle of at most 8 bytes");
    uint64_t u = 0;
    memcpy(&u, &v, sizeof(T));
    u = sim::exchange(u, src & 31);
    memcpy(&v, &u, sizeof(T));
    return v;
}
template <typename T> static inline T __shfl_up_sync(unsigned m, T v, unsigned delta)
{
    const int lane = (int)(threadIdx.x & 31u);
    return __shfl_sync(m, v, lane >= (int)delta ? lane - (int)delta : lane);
}
template <typename T> static inline T __shfl_down_sync(unsigned m, T v, unsigned delta)
{
    const int lane = (int)(threadIdx.x & 31u);
    return __shfl_sync(m, v, lane + (int)delta < 32 ? lane + (int)delta : lane);
}
template <typename T> static inline T __shfl_xor_sync(unsigned m, T v, int x) { return __shfl_sync(m, v, (int)(threadIdx.x & 31u) ^ x); }

// shared-window address of a generic pointer into the CTA's dynamic shared memory
static inline size_t __cvta_generic_to_shared(const void *p) { return (size_t)((const unsigned char *)p - sim::smem_base); }
