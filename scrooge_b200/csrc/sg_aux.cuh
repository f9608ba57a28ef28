// sg_aux.cuh -- the kernels around the aligner: sequence ingest (ASCII -> 2 bit), CIGAR run compaction
// (scan + gather).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace sg {

// ---- ingest -----------------------------------------------------------------------------------------
// Replaces reference single_ascii_to_twobit_string (src/genasm_gpu.cu:640-685), where every block
// redundantly converts the whole blob with 32 threads and byte loads.  Here a thread converts 16 bases:
// one 16-byte load, SWAR conversion of 4 bases per 32-bit register, one 4-byte store; a warp reads 512
// contiguous bytes and writes 128.  HBM-bound: 1 B read + 0.25 B written per base.
//
// Bits 2:1 of a letter (either case) are A 0, C 1, T 2, G 3 -- the code with G and T swapped.  pack4 works on these raw
// codes (4 bases of a register -> 8 bits, base k at bits 2k+1:2k, in the TOP byte of the result); the swap is undone
// once per packed word: code = raw ^ (raw >> 1 & 0x55555555) (src/genasm_cpu.cpp:87-90: A0 C1 G2 T3).
// Validity: every byte of v is 0x0k, so the low 16 bits of v | v >> 12 are the PRMT selector (k0, k2, k1, k3): one permute
// of "ACTG" reproduces the letters the codes stand for, one permute brings the (case-folded) input into the same order.
__device__ __forceinline__ uint32_t pack4(uint32_t w, uint32_t &bad)
{
    const uint32_t v = (w >> 1) & 0x03030303u;                  // raw 2-bit code in each byte
    const uint32_t e = __byte_perm(0x47544341u /* "ACTG" */, 0u, v | (v >> 12));
    const uint32_t u = __byte_perm(w & 0xDFDFDFDFu, 0u, 0x3120u);  // folded case, bytes (b0, b2, b1, b3)
    bad |= e ^ u;
    return v * 0x01041040u;                                     // 4 codes -> bits 31:24
}

// the four top bytes of a..d as one word (a's in the lowest byte), G/T swap undone
__device__ __forceinline__ uint32_t pack16(uint32_t a, uint32_t b, uint32_t c, uint32_t d)
{
    const uint32_t lo = __byte_perm(a, b, 0x0073u), hi = __byte_perm(c, d, 0x0073u);   // (a.b3, b.b3, 0.., ..)
    const uint32_t raw = __byte_perm(lo, hi, 0x5410u);
    return raw ^ ((raw >> 1) & 0x55555555u);
}

__device__ __forceinline__ void pack_word_slow(const char *__restrict__ ascii, uint64_t n_bases, uint64_t base, uint32_t &out, uint32_t &bad)
{
    out = 0;   // tail (or padding) word: byte by byte, missing bases are zero
    for (int k = 0; k < 16; k++) {
        if (base + k < n_bases) {
            uint32_t b1 = 0;
            uint32_t raw = pack4((uint32_t)(uint8_t)ascii[base + k] | 0x41414100u, b1) >> 24;
            bad |= b1 ? 1u : 0u;
            raw &= 3u;
            out |= (raw ^ (raw >> 1)) << (2 * k);
        }
    }
}

__device__ __forceinline__ void pack_report_bad(const char *__restrict__ ascii, uint64_t n_bases, uint64_t base, unsigned long long *bad_pos)
{
    for (int k = 0; k < 16 && base + k < n_bases; k++) {   // rare: find the first offending base of this word
        const uint32_t c = (uint32_t)(uint8_t)ascii[base + k] & 0xDFu;
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') {
            atomicMin(bad_pos, (unsigned long long)(base + k));
            break;
        }
    }
}

// Two words per thread and iteration, `stride` words apart (both 16-byte loads are in flight before either is used:
// 32 bytes per thread outstanding is what it takes to cover the HBM latency at this occupancy).
__global__ void __launch_bounds__(256) pack_2bit_kernel(const char *__restrict__ ascii, uint64_t n_bases,
                                                         uint32_t *__restrict__ packed, uint64_t n_words,
                                                         unsigned long long *__restrict__ bad_pos, uint64_t w_first)
{
    const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
    for (uint64_t w0 = w_first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; w0 < n_words; w0 += 2 * stride) {
        const uint64_t w1 = w0 + stride;
        const uint64_t b0 = w0 * 16ull, b1 = w1 * 16ull;
        const bool full0 = b0 + 16ull <= n_bases, full1 = w1 < n_words && b1 + 16ull <= n_bases;
        uint4 v0 = make_uint4(0, 0, 0, 0), v1 = make_uint4(0, 0, 0, 0);
        if (full0) v0 = __ldg(reinterpret_cast<const uint4 *>(ascii + b0));
        if (full1) v1 = __ldg(reinterpret_cast<const uint4 *>(ascii + b1));
        uint32_t bad0 = 0, bad1 = 0, out0, out1 = 0;
        if (full0) out0 = pack16(pack4(v0.x, bad0), pack4(v0.y, bad0), pack4(v0.z, bad0), pack4(v0.w, bad0));
        else pack_word_slow(ascii, n_bases, b0, out0, bad0);
        packed[w0] = out0;
        if (w1 < n_words) {
            if (full1) out1 = pack16(pack4(v1.x, bad1), pack4(v1.y, bad1), pack4(v1.z, bad1), pack4(v1.w, bad1));
            else pack_word_slow(ascii, n_bases, b1, out1, bad1);
            packed[w1] = out1;
        }
        if (bad0) pack_report_bad(ascii, n_bases, b0, bad_pos);
        if (bad1) pack_report_bad(ascii, n_bases, b1, bad_pos);
    }
}

// The same conversion with the ASCII staged through shared memory by the bulk-copy engine (cp.async.bulk = TMA's 1-D form,
// completion on an mbarrier): one elected thread keeps kPackStages tiles of kPackTile bytes in flight per CTA, all threads
// convert a landed tile (conflict-free 16-byte shared loads, coalesced 4-byte stores).  It covers the whole tiles of a blob,
// the tail goes to pack_2bit_kernel (SG_PACK=plain: everything does).  Measured on the benchmark's 11.4 G-base text blob:
// 2.19 ms = 6.48 TB/s of read + written bytes (0.99 of the measured copy peak) against 2.51 ms = 5.67 TB/s for the plain
// kernel; tile x stages 8 KB x 4 / x 8: 2.30 / 2.26 ms, 16 KB x 3 / x 4 / x 6: 2.26 / 2.28 / 2.19, 32 KB x 2 / x 3: 2.21 / 2.23.
#ifndef SG_PACK_TILE
#define SG_PACK_TILE 16384
#endif
#ifndef SG_PACK_STAGES
#define SG_PACK_STAGES 6
#endif
constexpr int kPackTile = SG_PACK_TILE;      // ASCII bytes per stage (16 KB = 1024 packed words)
constexpr int kPackStages = SG_PACK_STAGES;  // 6 x 16 KB = 96 KB of shared memory per CTA, two CTAs per SM
// The "side" geometry: 4 x 8 KB = 32 KB per CTA, small enough to run in the CTA slot the alignment kernel leaves free on
// every SM (SG_PACK_SIDE: ingest of the next batch beside the alignment of the current one); alone it is 5 % slower.
constexpr int kPackSideTile = 8192;
constexpr int kPackSideStages = 4;

#ifndef SG_SIM   // mbarrier / cp.async.bulk have no host twin: the simulation (tests/sim) runs the plain ingest kernel only
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t phase)
{
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t"
        "}" ::"r"(bar), "r"(phase) : "memory");
}

template <int kPackTile, int kPackStages>
__global__ void __launch_bounds__(256) pack_2bit_bulk_kernel(const char *__restrict__ ascii, uint64_t n_tiles,
                                                              uint32_t *__restrict__ packed, unsigned long long *__restrict__ bad_pos)
{
    extern __shared__ __align__(128) uint8_t pack_smem[];
    __shared__ __align__(8) uint64_t bars[kPackStages];
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(pack_smem);
    const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
    const uint64_t n_bases = n_tiles * (uint64_t)kPackTile;
    // tiles of this CTA: blockIdx.x, blockIdx.x + gridDim.x, ...
    const uint64_t mine = n_tiles > blockIdx.x ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    auto issue = [&](uint64_t k) {   // elected thread: tile k of this CTA into stage k % kPackStages
        const uint32_t st = (uint32_t)(k % kPackStages);
        const char *src = ascii + (blockIdx.x + k * (uint64_t)gridDim.x) * (uint64_t)kPackTile;
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + st * 8u), "r"((uint32_t)kPackTile) : "memory");
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem0 + st * (uint32_t)kPackTile),
                     "l"(src), "r"((uint32_t)kPackTile), "r"(bar0 + st * 8u)
                     : "memory");
    };
    if (threadIdx.x == 0) {
        for (int s = 0; s < kPackStages; s++) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + s * 8u) : "memory");
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0)
        for (uint64_t k = 0; k < mine && k < (uint64_t)kPackStages; k++) issue(k);
    for (uint64_t k = 0; k < mine; k++) {
        const uint32_t st = (uint32_t)(k % kPackStages), phase = (uint32_t)((k / kPackStages) & 1u);
        mbar_wait(bar0 + st * 8u, phase);
        const uint64_t tile = blockIdx.x + k * (uint64_t)gridDim.x;
        const uint4 *sv = reinterpret_cast<const uint4 *>(pack_smem + st * kPackTile);
        uint32_t *dst = packed + tile * (uint64_t)(kPackTile / 16);
#pragma unroll
        for (int j = 0; j < kPackTile / 16 / 256; j++) {
            const uint4 v = sv[j * 256 + threadIdx.x];
            uint32_t bad = 0;
            const uint32_t w = pack16(pack4(v.x, bad), pack4(v.y, bad), pack4(v.z, bad), pack4(v.w, bad));
            dst[j * 256 + threadIdx.x] = w;
            if (bad) pack_report_bad(ascii, n_bases, (tile * (uint64_t)(kPackTile / 16) + j * 256 + threadIdx.x) * 16ull, bad_pos);
        }
        __syncthreads();   // every thread is done with the stage before it is refilled
        if (threadIdx.x == 0 && k + kPackStages < mine) {
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy reads before the async-proxy write
            issue(k + kPackStages);
        }
    }
}

#endif  // !SG_SIM

// ---- CIGAR run compaction -----------------------------------------------------------------------------
// The aligner writes each alignment's runs into its own slab slot (capacity known in advance, no device
// allocator, no linked list: cf. reference src/cuda_list.hpp).  Compaction = exclusive scan of the run
// counts, then a gather into one dense byte array that is copied to the host in a single transfer.

constexpr int kScanBlock = 256;
constexpr int kScanItems = 8;  // elements per thread
constexpr int kScanTile = kScanBlock * kScanItems;

__device__ __forceinline__ uint64_t block_exclusive_scan(uint64_t v, uint64_t *total)
{
    __shared__ uint64_t warp_sums[kScanBlock / 32];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    uint64_t x = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        uint64_t y = __shfl_up_sync(0xFFFFFFFFu, x, o);
        if (lane >= o) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t s = lane < kScanBlock / 32 ? warp_sums[lane] : 0;
#pragma unroll
        for (int o = 1; o < kScanBlock / 32; o <<= 1) {
            uint64_t y = __shfl_up_sync(0xFFFFFFFFu, s, o);
            if (lane >= o) s += y;
        }
        if (lane < kScanBlock / 32) warp_sums[lane] = s;
    }
    __syncthreads();
    const uint64_t warp_off = wid ? warp_sums[wid - 1] : 0;
    if (total) *total = warp_sums[kScanBlock / 32 - 1];
    const uint64_t res = warp_off + x - v;
    __syncthreads();
    return res;
}

// pass 1: per-tile totals
__global__ void __launch_bounds__(kScanBlock) scan_tile_sums_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                     uint64_t *__restrict__ tile_sums)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile;
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        const uint64_t idx = base + (uint64_t)k * kScanBlock + threadIdx.x;
        if (idx < n) s += in[idx];
    }
    uint64_t total;
    block_exclusive_scan(s, &total);
    if (threadIdx.x == 0) tile_sums[blockIdx.x] = total;
}

// pass 2: exclusive scan of the tile totals in one block (n_tiles is small: n / 2048)
__global__ void __launch_bounds__(kScanBlock) scan_tile_offsets_kernel(uint64_t *__restrict__ tile_sums, uint64_t n_tiles)
{
    uint64_t carry = 0;
    for (uint64_t base = 0; base < n_tiles; base += kScanBlock) {
        const uint64_t idx = base + threadIdx.x;
        const uint64_t v = idx < n_tiles ? tile_sums[idx] : 0;
        uint64_t total;
        const uint64_t ex = block_exclusive_scan(v, &total);
        if (idx < n_tiles) tile_sums[idx] = carry + ex;
        carry += total;
    }
}

// pass 3: final offsets; out[n] = grand total
__global__ void __launch_bounds__(kScanBlock) scan_finish_kernel(const uint32_t *__restrict__ in, uint64_t n,
                                                                  const uint64_t *__restrict__ tile_offs,
                                                                  uint64_t *__restrict__ out)
{
    const uint64_t base = (uint64_t)blockIdx.x * kScanTile + (uint64_t)threadIdx.x * kScanItems;
    uint32_t v[kScanItems];
    uint64_t s = 0;
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        v[k] = base + k < n ? in[base + k] : 0u;
        s += v[k];
    }
    uint64_t ex = block_exclusive_scan(s, nullptr) + tile_offs[blockIdx.x];
#pragma unroll
    for (int k = 0; k < kScanItems; k++) {
        if (base + k < n) out[base + k] = ex;
        ex += v[k];
        if (base + k == n - 1) out[n] = ex;
    }
}

// gather: GROUP lanes per alignment copy its runs from the slab slot to the dense array.  Source and destination have
// unrelated byte alignments, so the body is copied as 32-bit words aligned to the DESTINATION, each assembled from the
// two aligned source words it straddles (one funnel shift); at most 3 head and 3 tail bytes go byte by byte.
template <int GROUP>
__global__ void __launch_bounds__(256) gather_runs_kernel(const uint8_t *__restrict__ slab, const uint64_t *__restrict__ slab_off,
                                                           const uint32_t *__restrict__ nruns, const uint64_t *__restrict__ run_off,
                                                           uint64_t n, uint8_t *__restrict__ runs)
{
    const uint64_t gid = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / GROUP;
    const uint32_t sub = threadIdx.x % GROUP;
    const uint64_t ngroups = ((uint64_t)gridDim.x * blockDim.x) / GROUP;
    for (uint64_t a = gid; a < n; a += ngroups) {
        const uint8_t *src = slab + slab_off[a];
        uint8_t *dst = runs + run_off[a];
        const uint32_t cnt = nruns[a];
        const uint32_t head = min(cnt, (uint32_t)((4u - ((uint32_t)(uintptr_t)dst & 3u)) & 3u));
        if (sub < head) dst[sub] = src[sub];
        const uint32_t body = (cnt - head) >> 2;                       // whole destination words
        const uint8_t *s = src + head;
        const uint32_t m = (uint32_t)(uintptr_t)s & 3u;
        const uint32_t *sw = reinterpret_cast<const uint32_t *>(s - m);
        uint32_t *dw = reinterpret_cast<uint32_t *>(dst + head);
        const uint32_t src_words = (m + 4u * body + 3u) >> 2;          // aligned source words that hold the body
        for (uint32_t w = sub; w < body; w += GROUP) {
            const uint32_t lo = sw[w];
            const uint32_t hi = (m != 0u && w + 1u < src_words) ? sw[w + 1u] : 0u;
            dw[w] = __funnelshift_r(lo, hi, 8u * m);
        }
        const uint32_t done = head + 4u * body;
        if (sub < cnt - done) dst[done + sub] = src[done + sub];
    }
}

}  // namespace sg
