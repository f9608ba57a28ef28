// sg_align_generic.cuh -- the alignment kernel for ANY window configuration (W, O) chosen at run time.
//
// The reference fixes the window size W and the overlap O at compile time (-DCLI_W/-DCLI_O, src/genasm_cpu.cpp:22-35,
// src/genasm_gpu.cu:1-63) and its evaluation sweeps both (scripts/profile.py:66-100 cpu_sweep_wo / cpu_sweep_o,
// :595-640 accuracy sweeps over W in 32, 64, 96, 128).  genasm_delta_kernel<W> (sg_align_delta.cuh) is tuned for the two
// configurations the reference ships -- 64/33 and 32/17 -- with everything unrolled around W - O = 31 or 15 traceback
// steps in one 32-bit word.  This kernel runs the same algorithm with W, O as kernel parameters:
//
//   * vectors of NW = ceil(W / 32) words (template parameter, 1..8), pattern position J at bit 32*NW-1-J: a window narrower
//     than the vector is a window with more padding rows, which the recurrence already handles (the last window of every
//     read has m < W);
//   * W columns per window whatever n is (the "matches nothing" mask stands in for the columns i >= n), one column =
//     delta_column<NW> as in the tuned kernel;
//   * the op planes A = V | H, B = ~V & (H | E) of the W-O+1 traceback columns kept for the top ceil((W-O)/32) words;
//   * the traceback as a per-lane loop over up to 2 (W-O) <= 126 steps into four-word register streams, run-length encoded
//     after the walk as in the tuned kernel.
//
// Limits: 2 <= W <= 256, 0 <= O < W, W - O <= 128.  A run is one byte, (op << 6) | count, and a run can be W - O long: with W - O > 63 (the
// WIDE instantiations: eight-word op streams) a longer run is split into bytes with count 0, each meaning "63 more".
// Same one-lane-per-alignment mapping, work queue and outputs as the tuned kernel; one warp per CTA, shared memory sized
// at launch.  Checked bit-exact against the unmodified reference built at thirteen further window
// configurations (tests/test_gpu_parity.py::test_window_configurations, goldens in tests/golden/golden_w*_o*.json) and
// against the tuned kernels at 64/33 and 32/17 (::test_generic_kernel_equals_tuned_kernels).
#pragma once
#include "sg_align_delta.cuh"

namespace sg {

struct GenericGeom {
    int W;      // window size in characters (columns per window)
    int TBL;    // W - O: traceback limit (src/genasm_cpu.cpp:50)
    int NWT;    // plane words kept per traceback column: pattern positions 0..TBL-1
    uint32_t *planes;   // GP kernels: per-CTA plane scratch in global memory, generic_plane_words() words each
};

// op planes of one warp: [column][word][plane][lane]
__host__ __device__ inline int generic_plane_words(int TBL) { return (TBL + 1) * ((TBL + 31) / 32) * 2 * 32; }

// shared memory of one warp: pattern masks and the window's text codes, plus the op planes unless they live in global
// memory (GP: a window configuration with W - O > 32 needs 20-32 KB of planes per warp, which would leave 6-9 warps per
// SM; in global memory the planes of all resident warps fit in the L2 cache and occupancy is bounded by registers)
__host__ __device__ inline int generic_smem_words(int NW, int W, int TBL, bool global_planes)
{
    const int pm = 5 * NW * 32;                 // [base code 0..3, 4 = "matches nothing"][word][lane]
    const int tw = ((W + 15) / 16) * 32;        // [text word][lane]: the window's 2-bit codes
    return pm + tw + (global_planes ? 0 : generic_plane_words(TBL));
}

// NW-word addition with carry propagation for any NW: one add.cc / addc.cc chain (a single asm statement, so that nothing
// can come between the carry-setting and the carry-using instructions)
#ifndef SG_SIM   // the host simulation (tests/sim) adds with a plain carry loop, see add_vec_any
template <int NW> struct AddChain;
template <> struct AddChain<1> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[1], const uint32_t (&b)[1], uint32_t (&s)[1])
    {
        asm("add.u32 %0, %1, %2;"
            : "=r"(s[0])
            : "r"(a[0]), "r"(b[0]));
    }
};
template <> struct AddChain<2> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[2], const uint32_t (&b)[2], uint32_t (&s)[2])
    {
        asm("add.cc.u32 %0, %2, %4; addc.u32 %1, %3, %5;"
            : "=r"(s[0]), "=r"(s[1])
            : "r"(a[0]), "r"(a[1]), "r"(b[0]), "r"(b[1]));
    }
};
template <> struct AddChain<3> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[3], const uint32_t (&b)[3], uint32_t (&s)[3])
    {
        asm("add.cc.u32 %0, %3, %6; addc.cc.u32 %1, %4, %7; addc.u32 %2, %5, %8;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(b[0]), "r"(b[1]), "r"(b[2]));
    }
};
template <> struct AddChain<4> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[4], const uint32_t (&b)[4], uint32_t (&s)[4])
    {
        asm("add.cc.u32 %0, %4, %8; addc.cc.u32 %1, %5, %9; addc.cc.u32 %2, %6, %10; addc.u32 %3, %7, %11;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]));
    }
};
template <> struct AddChain<5> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[5], const uint32_t (&b)[5], uint32_t (&s)[5])
    {
        asm("add.cc.u32 %0, %5, %10; addc.cc.u32 %1, %6, %11; addc.cc.u32 %2, %7, %12; addc.cc.u32 %3, %8, %13; addc.u32 %4, %9, %14;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]));
    }
};
template <> struct AddChain<6> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[6], const uint32_t (&b)[6], uint32_t (&s)[6])
    {
        asm("add.cc.u32 %0, %6, %12; addc.cc.u32 %1, %7, %13; addc.cc.u32 %2, %8, %14; addc.cc.u32 %3, %9, %15; addc.cc.u32 %4, %10, %16; addc.u32 %5, %11, %17;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]));
    }
};
template <> struct AddChain<7> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[7], const uint32_t (&b)[7], uint32_t (&s)[7])
    {
        asm("add.cc.u32 %0, %7, %14; addc.cc.u32 %1, %8, %15; addc.cc.u32 %2, %9, %16; addc.cc.u32 %3, %10, %17; addc.cc.u32 %4, %11, %18; addc.cc.u32 %5, %12, %19; addc.u32 %6, %13, %20;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]));
    }
};
template <> struct AddChain<8> {
    static __device__ __forceinline__ void run(const uint32_t (&a)[8], const uint32_t (&b)[8], uint32_t (&s)[8])
    {
        asm("add.cc.u32 %0, %8, %16; addc.cc.u32 %1, %9, %17; addc.cc.u32 %2, %10, %18; addc.cc.u32 %3, %11, %19; addc.cc.u32 %4, %12, %20; addc.cc.u32 %5, %13, %21; addc.cc.u32 %6, %14, %22; addc.u32 %7, %15, %23;"
            : "=r"(s[0]), "=r"(s[1]), "=r"(s[2]), "=r"(s[3]), "=r"(s[4]), "=r"(s[5]), "=r"(s[6]), "=r"(s[7])
            : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(a[4]), "r"(a[5]), "r"(a[6]), "r"(a[7]), "r"(b[0]), "r"(b[1]), "r"(b[2]), "r"(b[3]), "r"(b[4]), "r"(b[5]), "r"(b[6]), "r"(b[7]));
    }
};
#endif  // !SG_SIM
template <int NW>
__device__ __forceinline__ void add_vec_any(const uint32_t (&a)[NW], const uint32_t (&b)[NW], uint32_t (&s)[NW])
{
#ifdef SG_SIM
    uint64_t c = 0;
    for (int k = 0; k < NW; k++) {
        c += (uint64_t)a[k] + (uint64_t)b[k];
        s[k] = (uint32_t)c;
        c >>= 32;
    }
#else
    AddChain<NW>::run(a, b, s);
#endif
}

// delta_column (sg_align_delta.cuh) for any NW
template <int NW>
__device__ __forceinline__ void delta_column_any(uint32_t (&Pv)[NW], uint32_t (&Mv)[NW], const uint32_t (&pm)[NW], uint32_t (&Ph)[NW])
{
    uint32_t t[NW], s[NW], x[NW], Mh[NW], Phs[NW], Mhs[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) t[k] = ~pm[k] & Pv[k];
    add_vec_any<NW>(t, Pv, s);
#pragma unroll
    for (int k = 0; k < NW; k++) {
        x[k] = (s[k] ^ Pv[k]) | ~pm[k];
        Ph[k] = Mv[k] | ~(x[k] | Pv[k]);
        Mh[k] = Pv[k] & x[k];
    }
    shl1<NW>(Ph, Phs);
    shl1<NW>(Mh, Mhs);
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t xv = ~pm[k] | Mv[k];
        Pv[k] = Mhs[k] | ~(xv | Phs[k]);
        Mv[k] = Phs[k] & xv;
    }
}

// `count` bases starting at base `pos` of a packed blob as little-endian words out[0..NOUT): only the words that hold one of
// those bases are read (a window at the very end of a blob must not touch what follows the blob's padding)
template <int NOUT>
__device__ __forceinline__ void load_bases(const uint32_t *__restrict__ blob, uint64_t pos, int count, uint32_t (&out)[NOUT])
{
    const uint64_t first = pos >> 4;
    const uint64_t last = count > 0 ? (pos + (uint64_t)count - 1ull) >> 4 : first;
    const uint32_t sh = (uint32_t)(pos & 15u) * 2u;
    uint32_t w[NOUT + 1];
#pragma unroll
    for (int k = 0; k <= NOUT; k++) w[k] = (count > 0 && first + (uint64_t)k <= last) ? __ldg(blob + first + k) : 0u;
#pragma unroll
    for (int k = 0; k < NOUT; k++) out[k] = __funnelshift_r(w[k], w[k + 1], sh);
}

template <int NW, bool GP, bool WIDE>
__global__ void __launch_bounds__(32) genasm_generic_kernel(const AlignParams P, const GenericGeom G)
{
    constexpr int NTW = 2 * NW;                 // text / pattern words of a full-width window (16 bases each)
    extern __shared__ __align__(16) uint32_t smem_all[];
    const int lane = threadIdx.x & 31;
    const int W = G.W, TBL = G.TBL, NWT = G.NWT;
    uint32_t *pm_s = smem_all + lane;                                   // + (code * NW + k) * 32
    uint32_t *tw_s = smem_all + 5 * NW * 32 + lane;                     // + word * 32
    // + ((column * NWT + kk) * 2 + plane) * 32
    uint32_t *tb_s = (GP ? G.planes + (size_t)blockIdx.x * (size_t)generic_plane_words(TBL) : smem_all + 5 * NW * 32 + ((W + 15) / 16) * 32) + lane;

    const bool want_cigar = !(P.flags & 1u);

    bool have = false, drained = false;
    uint64_t pair = 0, t_pos = 0, t_begin = 0, t_end = 0, q_pos = 0, q_end = 0;
    int64_t ed = 0;
    uint8_t *out = nullptr, *out_end = nullptr;
    uint32_t nruns = 0;
    uint64_t entries = 0;
    bool overflow = false;

    while (true) {
        // ---- work queue: a lane without an alignment takes the next one (as genasm_delta_kernel) ----
        if (!have && !drained) {
            while (true) {
                uint64_t idx = atomicAdd(P.counter, 1ull);
                if (idx >= P.n) { drained = true; break; }
                if (P.order) idx = P.order[idx];   // longest-first launch order (reference src/tests.cu:377)
                uint64_t ql = P.query_len[idx];
                if (ql == 0) {  // zero windows: distance 0, empty CIGAR (src/tests.cu:243,246)
                    P.edit[idx] = 0;
                    P.ref_consumed[idx] = 0;
                    P.nruns[idx] = 0;
                    P.status[idx] = 0;
                    if (P.dc_entries) P.dc_entries[idx] = 0;
                    if (P.windows) P.windows[idx] = 0;
                    continue;
                }
                pair = idx;
                t_begin = t_pos = P.text_start[idx];
                t_end = t_pos + P.text_len[idx];
                q_pos = P.query_start[idx];
                q_end = q_pos + ql;
                ed = 0;
                nruns = 0;
                entries = 0;
                overflow = false;
                if (want_cigar) {
                    out = P.slab + P.slab_off[idx];
                    out_end = P.slab + P.slab_off[idx + 1];
                }
                have = true;
                break;
            }
        }
        if (__all_sync(0xFFFFFFFFu, !have)) break;

        // ---- window setup (src/genasm_cpu.cpp:411-420, pattern masks :178-198) ----
        uint32_t Pv[NW], Mv[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) { Pv[k] = 0; Mv[k] = 0; }
        int n = 0, m = 0;
        if (have) {
            const uint64_t tl = t_end - t_pos, ql = q_end - q_pos;
            n = tl < (uint64_t)W ? (int)tl : W;
            m = ql < (uint64_t)W ? (int)ql : W;
            uint32_t tw[NTW], pw[NTW];
            load_bases<NTW>(P.text, t_pos, n, tw);
            load_bases<NTW>(P.query, q_pos, m, pw);
#pragma unroll
            for (int k = 0; k < NTW; k++)
                if (k * 16 < W) tw_s[k * 32] = tw[k];
            uint32_t p0[NW], p1[NW], hm[NW];
            pattern_planes<NW>(pw, p0, p1);
            ones_shl<NW>(32 * NW - m, hm);
#pragma unroll
            for (int k = 0; k < NW; k++) {
                pm_s[(0 * NW + k) * 32] = (p1[k] | p0[k]) & hm[k];
                pm_s[(1 * NW + k) * 32] = (p1[k] | ~p0[k]) & hm[k];
                pm_s[(2 * NW + k) * 32] = (~p1[k] | p0[k]) & hm[k];
                pm_s[(3 * NW + k) * 32] = (~p1[k] | ~p0[k]) & hm[k];
                pm_s[(4 * NW + k) * 32] = hm[k];   // a character that matches nothing: columns i >= n
                Pv[k] = hm[k];                      // boundary column D(n,J) = m-J (src/genasm_cpu.cpp:225-231)
            }
        }
        __syncwarp();

        // ---- DC: columns W-1 .. 0 (src/genasm_cpu.cpp:210-288 as +-1 deltas, see sg_align_delta.cuh) ----
        // one 16-base text word at a time, the code of the current column kept in the word's top two bits
        for (int wi = (W - 1) >> 4; wi >= 0; wi--) {
            uint32_t cw = tw_s[wi * 32];
            const int top = W - 1 - 16 * wi < 15 ? W - 1 - 16 * wi : 15;
            cw <<= (15 - top) * 2;
            uint32_t *tbp = tb_s + ((16 * wi + top) * NWT * 2) * 32;   // planes of column 16 wi + top
#pragma unroll 4
            for (int ii = top; ii >= 0; ii--) {
                const int i = 16 * wi + ii;
                uint32_t code = cw >> 30;
                cw <<= 2;
                if (i >= n) code = 4u;
                uint32_t pm[NW], Ph[NW];
                const uint32_t *pmc = pm_s + code * (NW * 32);
#pragma unroll
                for (int k = 0; k < NW; k++) pm[k] = pmc[k * 32];
                delta_column_any<NW>(Pv, Mv, pm, Ph);
                if (i <= TBL) {
#pragma unroll
                    for (int kk = 0; kk < NW; kk++) {
                        if (kk < NWT) {
                            const int k = NW - 1 - kk;
                            tbp[kk * 64] = Pv[k] | Ph[k];
                            tbp[kk * 64 + 32] = ~Pv[k] & (Ph[k] | pm[k]);
                        }
                    }
                }
                tbp -= NWT * 64;
            }
        }
        __syncwarp();
        if (!have) continue;

        {   // window distance d_w = D(0,0) (src/genasm_cpu.cpp:278-283)
            int dw = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) dw += __popc(Pv[k]) - __popc(Mv[k]);
            entries += (uint64_t)(dw + 1) * (uint64_t)(n + 1) + kWindowUnit;
        }

        // ---- TB (src/genasm_cpu.cpp:290-409): op = 2A + B = 0 '=', 1 'X', 2 'I', 3 'D'; the two bits of step k go to bit k of
        // two register-resident streams (at most 2 (W-O) <= 126 steps), as in genasm_delta_kernel's generic walk ----
        constexpr int SW = WIDE ? 8 : 4;                  // WIDE: W - O up to 128, at most 256 steps per window
        const int jmax = m < TBL ? m : TBL;
        int i = 0, j = 0;
        uint32_t hs[SW], ls[SW];
        {
            const uint32_t col_words = (uint32_t)NWT * 64u;    // planes of one column
            const uint32_t *addr = tb_s;                       // column i, word j >> 5
            uint32_t mask = 0x80000000u;                       // pattern position j & 31, one-hot from the top
            uint32_t ca = addr[0], cb = addr[32];
            bool more = true;                                  // jmax >= 1 and TBL >= 1
#pragma unroll
            for (int w = 0; w < SW; w++) {
                uint32_t h = 0u, l = 0u, bit = 1u;
                if (more) {
                    do {
                        const bool hi = (ca & mask) != 0u;
                        const bool lo = (cb & mask) != 0u;
                        if (hi) h |= bit;
                        if (lo) l |= bit;
                        bit <<= 1;
                        if (!(hi && !lo)) { i++; addr += col_words; }          // every op but 'I' consumes a text character
                        if (!(hi && lo)) {                                     // every op but 'D' consumes a pattern character
                            j++;
                            mask = __funnelshift_r(mask, mask, 1);
                            if (mask == 0x80000000u) addr += 64;               // next plane word of the column
                        }
                        more = j < jmax && i < TBL;
                        if (more) { ca = addr[0]; cb = addr[32]; }
                    } while (bit != 0u && more);
                }
                hs[w] = h;
                ls[w] = l;
            }
        }

        // ---- RLE on the streams: per-window runs, flushed at window end, never merged across windows (quirk Q2); the
        // procedure of genasm_delta_kernel for SW stream words ----
        uint32_t e[SW];
        uint32_t edits = 0u, nb = 0u;
        int steps = j;                                   // every step but a 'D' consumes a pattern character
#pragma unroll
        for (int w = 0; w < SW; w++) {
            steps += __popc(hs[w] & ls[w]);
            edits += __popc(hs[w] | ls[w]);              // every op but '=' is an edit
            const uint32_t hn = __funnelshift_r(hs[w], w + 1 < SW ? hs[w + 1] : 0u, 1);
            const uint32_t ln = __funnelshift_r(ls[w], w + 1 < SW ? ls[w + 1] : 0u, 1);
            e[w] = (hs[w] ^ hn) | (ls[w] ^ ln);
        }
#pragma unroll
        for (int w = 0; w < SW; w++) {
            const int last = steps - 1 - 32 * w;         // steps >= 1: a window with m >= 1 takes at least one step
            if (last >= 0 && last < 32) e[w] |= 1u << last;
            if (last < 31) e[w] &= last < 0 ? 0u : (2u << last) - 1u;   // nothing beyond the last step
            nb += __popc(e[w]);
        }
        if constexpr (WIDE) {
            // a run can be up to 127 long and the run byte holds 6 bits: a run of c > 63 is written as (c - 1) / 63 bytes
            // with count 0 ("63 more of this op follow", SG_RUN_COUNT) and one byte with the rest; nruns counts bytes
            uint8_t *o = out;
            nb = 0u;
            int st = -1;
#pragma unroll
            for (int w = 0; w < SW; w++) {
                uint32_t ew = e[w];
                const uint32_t h7 = __funnelshift_l(hs[w], hs[w], 7), l6 = __funnelshift_l(ls[w], ls[w], 6);
                while (ew) {
                    const int p = __ffs((int)ew) - 1;
                    const uint32_t rh = __funnelshift_r(h7, h7, p), rl = __funnelshift_r(l6, l6, p);
                    const uint32_t opb = ((rh & 0x80u) | (rl & ~0x80u)) & 0xC0u;
                    uint32_t c = (uint32_t)(p - st);
                    while (true) {
                        const uint32_t piece = c > 63u ? 0u : c;
                        if (want_cigar) {
                            if (o < out_end) *o++ = (uint8_t)(opb | piece);
                            else overflow = true;
                        }
                        nb++;
                        if (c <= 63u) break;
                        c -= 63u;
                    }
                    st = p;
                    ew &= ew - 1u;
                }
                st -= 32;
            }
            out = o;
        } else {
            const bool fits = !want_cigar || (uint64_t)(out_end - out) >= (uint64_t)nb;
            if (!fits) overflow = true;
            if (want_cigar && fits) {
                uint8_t *o = out;
                out += nb;
                int st = -1;                                 // step before the current run's first, relative to word w
#pragma unroll
                for (int w = 0; w < SW; w++) {
                    uint32_t ew = e[w];
                    const uint32_t h7 = __funnelshift_l(hs[w], hs[w], 7), l6 = __funnelshift_l(ls[w], ls[w], 6);
                    while (ew) {
                        const int p = __ffs((int)ew) - 1;
                        const uint32_t rh = __funnelshift_r(h7, h7, p), rl = __funnelshift_r(l6, l6, p);
                        const uint32_t t = (rh & 0x80u) | (rl & ~0x80u);
                        *o++ = (uint8_t)((t & 0xC0u) | (uint32_t)(p - st));
                        st = p;
                        ew &= ew - 1u;
                    }
                    st -= 32;
                }
            }
        }
        nruns += nb;
        t_pos += (uint64_t)i;
        q_pos += (uint64_t)j;
        ed += edits;
        if (q_pos >= q_end) {
            P.edit[pair] = ed;
            P.ref_consumed[pair] = t_pos - t_begin;
            P.nruns[pair] = overflow ? 0u : nruns;   // nothing valid in the slot: the compaction must not read past it
            P.status[pair] = overflow ? 5 : 0;
            if (P.dc_entries) P.dc_entries[pair] = entries & (kWindowUnit - 1);
            if (P.windows) P.windows[pair] = (uint32_t)(entries >> 40);
            have = false;
        }
    }
}

}  // namespace sg
