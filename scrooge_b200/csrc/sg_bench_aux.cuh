// sg_bench_aux.cuh -- kernels of libscrooge_b200_bench.so: measurement, synthetic-data and checking helpers.  NOT part of
// the product library (libscrooge_b200.so holds the path only): the full-batch run checker, the synthetic generators
// and the integer-ALU peak probe.
#pragma once
#include <cstdint>
#include <cuda_runtime.h>
#include "sg_synth.h"

namespace sg {

// ---- full-batch consistency check of the compacted runs (measurement / test helper) -----------------------
// The size-independent properties of reference validateCigarString (src/tests.cu:106-169) that need no sequence data,
// for EVERY alignment of a batch: every run has a count in [1, max_count]; the counts of =,X,I sum to the query
// length, those of =,X,D to the consumed reference prefix, those of X,I,D to the edit distance.  One warp per alignment.
__global__ void __launch_bounds__(256) check_runs_kernel(const uint8_t *__restrict__ runs, const uint64_t *__restrict__ run_off,
                                                         uint64_t n, const uint64_t *__restrict__ query_len,
                                                         const int64_t *__restrict__ edit, const uint64_t *__restrict__ ref_consumed,
                                                         uint32_t max_count, unsigned long long *__restrict__ n_bad)
{
    const int lane = threadIdx.x & 31;
    const uint64_t warps = (uint64_t)gridDim.x * (blockDim.x / 32);
    for (uint64_t a = (uint64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5); a < n; a += warps) {
        const uint64_t r0 = run_off[a], r1 = run_off[a + 1];
        uint32_t q = 0, t = 0, e = 0, bad = 0;
        for (uint64_t k = r0 + lane; k < r1; k += 32) {
            const uint32_t b = runs[k], op = b >> 6;
            uint32_t c = b & 63u;
            bad |= ((c == 0u && max_count <= 63u) || c > max_count) ? 1u : 0u;
            if (c == 0u) c = 63u;     // W - O > 63: a byte with count 0 stands for 63 more of its op (SG_RUN_COUNT)
            q += op != 3u ? c : 0u;   // '=', 'X', 'I' consume the query
            t += op != 2u ? c : 0u;   // '=', 'X', 'D' consume the text
            e += op != 0u ? c : 0u;   // 'X', 'I', 'D' are edits
        }
#pragma unroll
        for (int o = 16; o; o >>= 1) {
            q += __shfl_xor_sync(0xFFFFFFFFu, q, o);
            t += __shfl_xor_sync(0xFFFFFFFFu, t, o);
            e += __shfl_xor_sync(0xFFFFFFFFu, e, o);
            bad |= __shfl_xor_sync(0xFFFFFFFFu, bad, o);
        }
        if (lane == 0 && (bad || q != query_len[a] || t != ref_consumed[a] || (int64_t)e != edit[a])) atomicAdd(n_bad, 1ull);
    }
}

// ---- synthetic pairs ----------------------------------------------------------------------------------
__global__ void __launch_bounds__(128) synth_pairs_kernel(SgSynthParams p, uint64_t first_pair, uint64_t n_pairs,
                                                           char *__restrict__ text, uint64_t text_stride,
                                                           uint64_t *__restrict__ text_len, char *__restrict__ reads)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n_pairs) return;
    char *t = text + k * text_stride;
    const uint64_t tl = sg_synth_pair(p, first_pair + k, t, reads + k * (uint64_t)p.read_len);
    text_len[k] = tl;
    for (uint64_t x = tl; x < text_stride; x++) t[x] = 'A';  // keep the whole slot packable
}

__global__ void __launch_bounds__(256) synth_genome_kernel(uint64_t seed, uint64_t first, uint64_t n, char *__restrict__ out)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = sg_synth_genome_base(seed, first + i);
}

__global__ void __launch_bounds__(128) synth_reads_kernel(SgSynthParams p, uint64_t first, uint64_t n, const char *__restrict__ genome,
                                                           uint64_t genome_len, char *__restrict__ reads, uint64_t *__restrict__ pos)
{
    const uint64_t k = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    pos[k] = sg_synth_read_from_genome(p, first + k, genome, genome_len, reads + k * (uint64_t)p.read_len);
}

// ---- integer-ALU peak probe -----------------------------------------------------------------------------
// Independent chains of the DC recurrence's own instructions.  kind 0: LOP3 only; 1: SHF (funnel shift) only;
// 2: two LOP3 per SHF (the DC mix); 3: LOP3 + IMAD alternating (alu pipe + fma pipe); 4-6: one DC entry
// (4 LOP3 + a 64-bit shift left by one) with the shift done as IMAD+SHF, IMAD.SHL+IMAD.WIDE, or IMAD.HI+IMAD+SHL.
#ifndef SG_SIM   // the peak probe is chains of inline PTX: nothing for the host simulation (tests/sim) to check
template <int KIND>
__global__ void __launch_bounds__(256) int32_peak_kernel(uint32_t *__restrict__ sink, int iters, uint32_t seed)
{
    constexpr int CH = 8;
    uint32_t a[CH], b[CH], c[CH];
#pragma unroll
    for (int k = 0; k < CH; k++) {
        a[k] = seed + threadIdx.x * 2654435761u + k;
        b[k] = a[k] * 40503u + 17u;
        c[k] = b[k] ^ 0x9E3779B9u;
    }
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int rep = 0; rep < 4; rep++) {
#pragma unroll
            for (int k = 0; k < CH; k++) {
                if (KIND == 0) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[k]) : "r"(c[k]), "r"(a[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(c[k]) : "r"(a[k]), "r"(b[k]));
                } else if (KIND == 1) {
                    asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(a[k]) : "r"(b[k]));
                    asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(b[k]) : "r"(c[k]));
                    asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(c[k]) : "r"(a[k]));
                } else if (KIND == 2) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[k]) : "r"(c[k]), "r"(a[k]));
                    asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(c[k]) : "r"(a[k]));
                } else if (KIND == 3) {
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("mad.lo.u32 %0, %0, 3, %1;" : "+r"(b[k]) : "r"(c[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(c[k]) : "r"(a[k]), "r"(b[k]));
                    asm volatile("mad.lo.u32 %0, %0, 5, %1;" : "+r"(a[k]) : "r"(b[k]));
                } else if (KIND == 4) {
                    // one DC entry as the kernel issues it today: 4 LOP3 + lo shift on the fma pipe + funnel shift
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[k]) : "r"(c[k]), "r"(a[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(c[k]) : "r"(a[k]), "r"(b[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x1E;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("shf.l.wrap.b32 %0, %1, %0, 1;" : "+r"(b[k]) : "r"(a[k]));
                    asm volatile("mad.lo.u32 %0, %0, 2, %1;" : "+r"(c[k]) : "r"(b[k]));
                } else if (KIND == 5) {
                    // 4 LOP3 + 64-bit shift entirely on the fma pipe: hi*2 then mad.wide(lo, 2, {0, hi*2})
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[k]) : "r"(c[k]), "r"(a[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(c[k]) : "r"(a[k]), "r"(b[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x1E;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("{ .reg .b64 t, u; .reg .b32 z; shl.b32 z, %1, 1; mov.b64 t, {0, z}; mad.wide.u32 u, %0, 2, t; mov.b64 {%0, %1}, u; }"
                                 : "+r"(b[k]), "+r"(c[k]));
                } else {
                    // 4 LOP3 + 64-bit shift on the fma pipe through the carry: lo*2, mulhi(lo,2), hi*2+carry
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xE8;" : "+r"(b[k]) : "r"(c[k]), "r"(a[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0xCA;" : "+r"(c[k]) : "r"(a[k]), "r"(b[k]));
                    asm volatile("lop3.b32 %0, %0, %1, %2, 0x1E;" : "+r"(a[k]) : "r"(b[k]), "r"(c[k]));
                    asm volatile("{ .reg .b32 cy; mul.hi.u32 cy, %0, 2; mad.lo.u32 %1, %1, 2, cy; shl.b32 %0, %0, 1; }"
                                 : "+r"(b[k]), "+r"(c[k]));
                }
            }
        }
    }
    uint32_t acc = 0;
#pragma unroll
    for (int k = 0; k < CH; k++) acc ^= a[k] ^ b[k] ^ c[k];
    if (acc == 0x12345678u) sink[0] = acc;  // keep the chains alive
}
#endif  // !SG_SIM
// ops per loop iteration; kinds 4-6 count one "DC entry" (4 LOP3 + a 64-bit shift) as 6 ops
constexpr int kPeakOpsPerIter[7] = {8 * 4 * 3, 8 * 4 * 3, 8 * 4 * 3, 8 * 4 * 4, 8 * 4 * 6, 8 * 4 * 6, 8 * 4 * 6};

}  // namespace sg
