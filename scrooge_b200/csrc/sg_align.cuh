// sg_align.cuh -- the B200 alignment kernel: windowed GenASM distance calculation (DC), traceback (TB)
// and per-window CIGAR run-length encoding, one alignment per lane.
//
// What it computes is bit-identical to reference src/genasm_cpu.cpp:210-438 (genasm_dc, genasm_tb,
// genasm); how it computes it is new:
//
//  * Mapping: ONE LANE PER ALIGNMENT, persistent warps, a lane-granular work queue.  The reference GPU
//    kernel (src/genasm_gpu.cu:296-420) gives one 64-thread block to an alignment, a thread to a text
//    column, and needs n+k barrier-separated wavefront steps per window with <= 50 % of lanes busy.
//    Here no lane ever waits for another lane's data: there are no shuffles, no barriers, and every lane
//    runs the full recurrence on its own registers.
//  * DC order: column-major in chunks of G rows.  The reference walks d outer / i inner and keeps a
//    W+1-entry forefront in memory.  Here the G entries R[d0..d0+G-1][i+1] of the previous column live
//    in registers (multi-word, one 32-bit register per word) together with their <<1 copies, and a
//    column step is, per 32-bit word and row: one shift and two LOP3
//        u      = (sP[r] | pm) & P[r-1]                  (P = previous column, sP = P << 1; off the critical path)
//        C[r]   = u & sP[r-1] & sC[r-1]                  (= mat & del & sub & ins of src/genasm_cpu.cpp:247-251)
//        sC[r]  = C[r] << 1
//    Two register sets alternate between "previous" and "current" column (the column loop is unrolled by
//    two), so no register is ever copied.
//    The last row of a chunk is kept as a forefront in shared memory so that a lane whose window needs
//    more than G rows continues with rows d0+G.. in its next phase (early termination at chunk
//    granularity; the reported distance is exact).
//  * Left-aligned vectors: pattern position J lives at bit W-1-J whatever m is (the reference puts it at
//    bit m-1-J, src/genasm_cpu.cpp:59,185-189).  The early-termination test is then always the sign bit
//    and the traceback bits of the first W-O pattern positions are always in the top word (this is
//    DENT, src/genasm_cpu.cpp:200-208,258-267, without the extract/insert step).
//  * SENE/DENT taken one step further: instead of storing R[d][i] for every row d (reference:
//    (W-O+1)*(K+1) half vectors = 8.3 KB per alignment), the DC accumulates, per traceback column i,
//    two words over all rows
//        V_i = OR_d ( R[d][i] & ~(R[d][i] << 1) )   bit J set <=> D(i,J+1) = D(i,J) - 1   (insertion edge)
//        H_i = OR_d ( R[d][i] & ~R[d][i+1] )        bit J set <=> D(i+1,J) = D(i,J) - 1   (deletion edge)
//    where D(i,J) = min{d : bit J of R[d][i] is 0} is the edit-distance matrix the R rows encode.  That is
//    8 bytes per column, 256 B per alignment, independent of the window distance; the mismatch bits
//    E_i = pm[text[i]] are read back from the pattern-mask table.  The traceback's tests (src/genasm_cpu.cpp:321-343)
//        can_ins = d>0 && zero(R[d-1][i],   J+1)  <=> D(i,J+1)   <= d-1
//        can_del = d>0 && zero(R[d-1][i+1], J)    <=> D(i+1,J)   <= d-1
//        can_sub = d>0 && zero(R[d-1][i+1], J+1)  <=> D(i+1,J+1) <= d-1
//    are evaluated with d == D(i,J) (an invariant of the walk from d_w = D(0,0)), so can_ins <=> V_i(J),
//    can_del <=> H_i(J), and when neither holds can_sub <=> text[i] != pattern[J] <=> E_i(J) by the DP
//    recurrence.  The j == m-1 special case and the i >= n limit fall out of the same bits.
//
// Per-warp shared memory (W=64): pattern masks 1 KB + forefront 16.25 KB + V/H 8 KB = 25.25 KB (8 warps
// per SM), every array laid out [column][lane] so that all accesses are bank-conflict free whatever
// column each lane is at.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace sg {


template <int W> struct WinCfg;
// O: window overlap; G: rows per DC chunk (about the typical window distance + 1: 10 kbp reads at 10 % need 7.9 rows
// per 64-column window on average, 150 bp reads at 5 % need 2.6 rows per 32-column window)
template <> struct WinCfg<64> { static constexpr int O = 33; static constexpr int G = 8; };
template <> struct WinCfg<32> { static constexpr int O = 17; static constexpr int G = 4; };

// Per-warp shared memory.  TMEM = false: the forefront lives in shared memory (one-warp CTAs, 8 per SM at W=64).
// TMEM = true: the forefront lives in tensor memory -- W columns x NW words per lane, private to the lane, addressed
// with a warp-uniform column: exactly the access shape of tcgen05.ld/st.32x32b -- which frees 16 KB of shared
// memory per warp (four-warp CTAs owning 128 TMEM columns each at W=64, 4 CTAs = 16 warps per SM).
template <int W, bool TMEM = false> struct SmemLayout {
    static constexpr int NW = W / 32;               // 32-bit words per bitvector
    static constexpr int TBL = W - WinCfg<W>::O;    // TB_LIMIT (src/genasm_cpu.cpp:50)
    static constexpr int TBCOLS = TBL + 1;          // traceback columns kept (0..TBL)
    static constexpr int PM_WORDS = 4 * NW * 32;    // [base code][lane][NW]
    static constexpr int FF_WORDS = TMEM ? 0 : (W + 1) * NW * 32;  // [column][lane][NW]
    static constexpr int TB_WORDS = TBCOLS * 2 * 32;    // [column][lane][V,H]
    static constexpr int STAGE_WORDS = TMEM ? W * 32 / 4 : 0;  // [run][lane] bytes (smem variant stages runs in FF)
    static constexpr int WORDS_PER_WARP = PM_WORDS + FF_WORDS + TB_WORDS + STAGE_WORDS;
    static constexpr int BYTES_PER_WARP = WORDS_PER_WARP * 4;
    static constexpr int WARPS_PER_CTA = TMEM ? 4 : 1;
    static constexpr int TMEM_COLS = W * NW < 32 ? 32 : W * NW;   // power of two >= 32: 128 (W=64), 32 (W=32)
    static constexpr int BYTES_PER_CTA = BYTES_PER_WARP * WARPS_PER_CTA + (TMEM ? 16 : 0);
};

#ifndef SG_SIM   // the host simulation (tests/sim) runs the delta kernel only: tensor-memory PTX has no host twin
// ---- tensor memory as per-lane scratch (sm_100a tcgen05) ----------------------------------------------
template <int NW> __device__ __forceinline__ void tmem_ld(uint32_t taddr, uint32_t (&v)[NW]);
template <> __device__ __forceinline__ void tmem_ld<1>(uint32_t taddr, uint32_t (&v)[1])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x1.b32 {%0}, [%1];" : "=r"(v[0]) : "r"(taddr));
}
template <> __device__ __forceinline__ void tmem_ld<2>(uint32_t taddr, uint32_t (&v)[2])
{
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr));
}
template <int NW> __device__ __forceinline__ void tmem_st(uint32_t taddr, const uint32_t (&v)[NW]);
template <> __device__ __forceinline__ void tmem_st<1>(uint32_t taddr, const uint32_t (&v)[1])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x1.b32 [%0], {%1};" ::"r"(taddr), "r"(v[0]));
}
template <> __device__ __forceinline__ void tmem_st<2>(uint32_t taddr, const uint32_t (&v)[2])
{
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(v[0]), "r"(v[1]));
}
__device__ __forceinline__ void tmem_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

#endif  // !SG_SIM

// The per-lane work counter carries the DC entries in its low 40 bits and the window count above them (one
// accumulator, no extra register): good for reads up to ~500 Mbp.
constexpr uint64_t kWindowUnit = 1ull << 40;

struct AlignParams {
    const uint32_t *text;
    const uint64_t *text_start;
    const uint64_t *text_len;
    const uint32_t *query;
    const uint64_t *query_start;
    const uint64_t *query_len;
    uint64_t n;
    uint32_t flags;
    uint8_t *slab;
    const uint64_t *slab_off;
    unsigned long long *counter;
    int64_t *edit;
    uint64_t *ref_consumed;
    uint32_t *nruns;
    uint8_t *status;
    uint64_t *dc_entries;  // optional: sum over windows of (d_w+1)*(n+1), the early-termination-minimal DC work
    uint32_t *windows;     // optional: number of windows of the alignment
    const uint32_t *order; // optional: the queue hands out alignment order[k] as its k-th item (a permutation of 0..n-1)
};

// ---- small helpers -------------------------------------------------------------------------------

// W bases (2W bits) starting at base `pos` of a packed blob, as 2W/32 little-endian words.
template <int NWIN>
__device__ __forceinline__ void load_window(const uint32_t *__restrict__ blob, uint64_t pos, uint32_t (&out)[NWIN])
{
    const uint32_t *p = blob + (pos >> 4);
    const uint32_t sh = (uint32_t)(pos & 15u) * 2u;
    uint32_t w[NWIN + 1];
#pragma unroll
    for (int k = 0; k <= NWIN; k++) w[k] = __ldg(p + k);
#pragma unroll
    for (int k = 0; k < NWIN; k++) out[k] = __funnelshift_r(w[k], w[k + 1], sh);
}

// even bits of x gathered into the low 16 bits
__device__ __forceinline__ uint32_t compress_even(uint32_t x)
{
    x &= 0x55555555u;
    x = (x | (x >> 1)) & 0x33333333u;
    x = (x | (x >> 2)) & 0x0F0F0F0Fu;
    x = (x | (x >> 4)) & 0x00FF00FFu;
    x = (x | (x >> 8)) & 0x0000FFFFu;
    return x;
}

// Left-aligned pattern bit planes of a window: bit 31-J of word NW-1 <-> pattern[J] for J < 32, and so on
// downwards.  plane0 = low bit of the base code, plane1 = high bit.
template <int NW>
__device__ __forceinline__ void pattern_planes(const uint32_t (&pw)[2 * NW], uint32_t (&p0)[NW], uint32_t (&p1)[NW])
{
#pragma unroll
    for (int k = 0; k < NW; k++) {
        uint32_t a = pw[2 * k], b = pw[2 * k + 1];
        uint32_t lo = compress_even(a) | (compress_even(b) << 16);
        uint32_t hi = compress_even(a >> 1) | (compress_even(b >> 1) << 16);
        // pattern positions 32k..32k+31 go to word NW-1-k, bit-reversed
        p0[NW - 1 - k] = __brev(lo);
        p1[NW - 1 - k] = __brev(hi);
    }
}

// ~0 << s over NW words (s may be >= 32*NW -> 0)
template <int NW>
__device__ __forceinline__ void ones_shl(int s, uint32_t (&out)[NW])
{
#pragma unroll
    for (int k = 0; k < NW; k++) {
        int t = s - 32 * k;  // shift seen by word k
        out[k] = t <= 0 ? 0xFFFFFFFFu : (t >= 32 ? 0u : (0xFFFFFFFFu << t));
    }
}

template <int NW>
__device__ __forceinline__ void shl1(const uint32_t (&in)[NW], uint32_t (&out)[NW])
{
    out[0] = in[0] << 1;
#pragma unroll
    for (int k = 1; k < NW; k++) out[k] = __funnelshift_l(in[k - 1], in[k], 1);
}

// NW-word shared-memory vector access (one LDS.64 / STS.64 when NW == 2; pointers are 8-byte aligned)
template <int NW> __device__ __forceinline__ void lds_vec(const uint32_t *p, uint32_t (&v)[NW]);
template <> __device__ __forceinline__ void lds_vec<1>(const uint32_t *p, uint32_t (&v)[1]) { v[0] = p[0]; }
template <> __device__ __forceinline__ void lds_vec<2>(const uint32_t *p, uint32_t (&v)[2])
{
    uint2 t = *reinterpret_cast<const uint2 *>(p);
    v[0] = t.x; v[1] = t.y;
}
template <int NW> __device__ __forceinline__ void sts_vec(uint32_t *p, const uint32_t (&v)[NW]);
template <> __device__ __forceinline__ void sts_vec<1>(uint32_t *p, const uint32_t (&v)[1]) { p[0] = v[0]; }
template <> __device__ __forceinline__ void sts_vec<2>(uint32_t *p, const uint32_t (&v)[2])
{
    *reinterpret_cast<uint2 *>(p) = make_uint2(v[0], v[1]);
}

#ifndef SG_SIM   // row-wise formulation (SG_DC=rows): not simulated
// G rows of one column: entries and their << 1
template <int NW> struct RowSet {
    static constexpr int G = WinCfg<NW * 32>::G;
    uint32_t C[G][NW];
    uint32_t S[G][NW];
};

// Boundary column R[d][n] = ones << d (src/genasm_cpu.cpp:225-231,239-245), left-aligned: row r of the chunk is
// ~0 << (W - m + d0 + r), and each row is the previous one shifted by one.
template <int W, int NW>
__device__ __forceinline__ void dc_boundary(RowSet<NW> &N, int m, int d0, uint32_t &V)
{
    ones_shl<NW>(W - m + d0, N.C[0]);
#pragma unroll
    for (int r = 0; r < RowSet<NW>::G; r++) {
        shl1<NW>(N.C[r], N.S[r]);
        V |= N.C[r][NW - 1] & ~N.S[r][NW - 1];
        if (r + 1 < RowSet<NW>::G) {
#pragma unroll
            for (int k = 0; k < NW; k++) N.C[r + 1][k] = N.S[r][k];
        }
    }
}

// One DC column: P holds column i+1, N receives column i.  F is row d0-1 of column i (the forefront left
// by the previous chunk; ignored when fm == ~0, i.e. in the first chunk of a window), XF carries
// (F & F<<1) | fm from column i+1 in and from column i out.  With TBCOL the insertion/deletion edge words
// of the column are OR-accumulated into V/H (top word only: the first W-O pattern positions).
template <int NW, bool TBCOL>
__device__ __forceinline__ void dc_column(const RowSet<NW> &P, RowSet<NW> &N, const uint32_t (&pm)[NW],
                                          const uint32_t (&F)[NW], const uint32_t fm, uint32_t (&XF)[NW], uint32_t &V,
                                          uint32_t &H)
{
    constexpr int G = RowSet<NW>::G;
    constexpr int TOP = NW - 1;
    uint32_t sF[NW];
    shl1<NW>(F, sF);
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t u = (P.S[0][k] | pm[k]) & XF[k];
        N.C[0][k] = u & (sF[k] | fm);
        XF[k] = (F[k] & sF[k]) | fm;
    }
    shl1<NW>(N.C[0], N.S[0]);
#pragma unroll
    for (int r = 1; r < G; r++) {
#pragma unroll
        for (int k = 0; k < NW; k++) {
            const uint32_t u = (P.S[r][k] | pm[k]) & P.C[r - 1][k];
            N.C[r][k] = u & P.S[r - 1][k] & N.S[r - 1][k];
        }
        shl1<NW>(N.C[r], N.S[r]);
    }
    if (TBCOL) {
#pragma unroll
        for (int r = 0; r < G; r++) {
            H |= N.C[r][TOP] & ~P.C[r][TOP];
            V |= N.C[r][TOP] & ~N.S[r][TOP];
        }
    }
}

// ---- the kernel ------------------------------------------------------------------------------------

template <int W, bool TMEM>
__global__ void __launch_bounds__(SmemLayout<W, TMEM>::WARPS_PER_CTA * 32, TMEM ? 4 : 1) genasm_align_kernel(const AlignParams P)
{
    using L = SmemLayout<W, TMEM>;
    constexpr int NW = L::NW;
    constexpr int NWIN = 2 * NW;  // words of a 2-bit window
    constexpr int TBL = L::TBL;
    constexpr int TBCOLS = L::TBCOLS;
    constexpr int G = WinCfg<W>::G;
    constexpr int TOP = NW - 1;
    constexpr int PMS = NW * 32;   // words between the masks of consecutive base codes
    constexpr int FFS = NW * 32;   // words between forefront columns
    constexpr int TBS = 2 * 32;    // words between traceback columns

    extern __shared__ __align__(16) uint32_t smem_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *smem = smem_all + warp * L::WORDS_PER_WARP;
    uint32_t *pm_s = smem + lane * NW;
    uint32_t *ff_s = smem + L::PM_WORDS + lane * NW;                      // forefront (smem variant only)
    uint32_t *tb_s = smem + L::PM_WORDS + L::FF_WORDS + lane * 2;
    uint8_t *stage_s = reinterpret_cast<uint8_t *>(smem + L::PM_WORDS + L::FF_WORDS + L::TB_WORDS) + lane;  // TMEM variant
    (void)ff_s; (void)stage_s;

    // tensor-memory forefront: one allocation per CTA, lanes 32*warp.. belong to this warp
    uint32_t tff = 0;
    if constexpr (TMEM) {
        uint32_t *tmem_addr_s = smem_all + L::WARPS_PER_CTA * L::WORDS_PER_WARP;
        if (warp == 0) {
            const uint32_t dst = (uint32_t)__cvta_generic_to_shared(tmem_addr_s);
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst), "r"((uint32_t)L::TMEM_COLS));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::);
        }
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        tff = *tmem_addr_s + ((uint32_t)(warp * 32) << 16);
    }
    auto ff_load = [&](const int i, uint32_t (&F)[NW]) {   // TMEM: asynchronous, ff_wait() before F is read
        if constexpr (TMEM) tmem_ld<NW>(tff + (uint32_t)(i * NW), F);
        else lds_vec<NW>(ff_s + i * FFS, F);
    };
    auto ff_wait = [&]() { if constexpr (TMEM) tmem_wait_ld(); };
    auto ff_store = [&](const int i, const uint32_t (&v)[NW]) {
        if constexpr (TMEM) tmem_st<NW>(tff + (uint32_t)(i * NW), v);
        else sts_vec<NW>(ff_s + i * FFS, v);
    };

    const bool want_cigar = !(P.flags & 1u);

    // lane state
    bool have = false, drained = false;
    uint64_t pair = 0, t_pos = 0, t_begin = 0, t_end = 0, q_pos = 0, q_end = 0;
    int64_t ed = 0;
    uint8_t *out = nullptr, *out_end = nullptr;
    uint32_t nruns = 0;
    uint64_t entries = 0;
    bool overflow = false;
    int d0 = 0, n = -1, m = 0;
    uint32_t tw[NWIN];
#pragma unroll
    for (int k = 0; k < NWIN; k++) tw[k] = 0;

    while (true) {
        // ---- work queue: a lane without an alignment takes the next one ----------------------------
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
                d0 = 0;
                have = true;
                break;
            }
        }
        if (__all_sync(0xFFFFFFFFu, !have)) break;

        // ---- window setup (first chunk of a window) ------------------------------------------------
        if (have && d0 == 0) {
            uint64_t tl = t_end - t_pos, ql = q_end - q_pos;
            n = tl < (uint64_t)W ? (int)tl : W;
            m = ql < (uint64_t)W ? (int)ql : W;
            load_window<NWIN>(P.text, t_pos, tw);
            uint32_t pw[NWIN];
            load_window<NWIN>(P.query, q_pos, pw);
            uint32_t p0[NW], p1[NW], hm[NW];
            pattern_planes<NW>(pw, p0, p1);
            ones_shl<NW>(W - m, hm);
            // pm[c] = ((p1 ^ C1) | (p0 ^ C0)) & hm : zero where pattern[J] == c (src/genasm_cpu.cpp:178-198),
            // low W-m bits zero (left-aligned representation)
            uint32_t m0[NW], m1[NW], m2[NW], m3[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) {
                m0[k] = (p1[k] | p0[k]) & hm[k];
                m1[k] = (p1[k] | ~p0[k]) & hm[k];
                m2[k] = (~p1[k] | p0[k]) & hm[k];
                m3[k] = (~p1[k] | ~p0[k]) & hm[k];
            }
            sts_vec<NW>(pm_s + 0 * PMS, m0);
            sts_vec<NW>(pm_s + 1 * PMS, m1);
            sts_vec<NW>(pm_s + 2 * PMS, m2);
            sts_vec<NW>(pm_s + 3 * PMS, m3);
        }
        const int nn = have ? n : -1;
        const uint32_t fm = d0 == 0 ? 0xFFFFFFFFu : 0u;  // "no row above this chunk"

        // ---- DC: one chunk of G rows, columns n .. 0 -------------------------------------------------
        RowSet<NW> A, B;
        uint32_t XF[NW];  // (F & F << 1) | fm of row d0-1 at the previous column
        // lanes without work (queue drained) ride along in the fast path: whatever they compute stays in their own
        // scratch and is never read
        const bool uniform = __all_sync(0xFFFFFFFFu, !have || n == W);

        // boundary column: forefront in, chunk's last row out, insertion edges of the column if it is a TB column
        auto boundary_column = [&](RowSet<NW> &Nv, const int i) {
            // row d0-1 of a boundary column is itself a boundary value, ones << (d0-1): no forefront access
            uint32_t F[NW], sF[NW];
            ones_shl<NW>(W - m + d0 - 1, F);
            shl1<NW>(F, sF);
#pragma unroll
            for (int k = 0; k < NW; k++) XF[k] = (F[k] & sF[k]) | fm;
            uint32_t V = 0;
            dc_boundary<W, NW>(Nv, m, d0, V);
            if (i < TBCOLS) {
                uint32_t vh[2];
                lds_vec<2>(tb_s + i * TBS, vh);
                vh[0] = (vh[0] & ~fm) | V;
                vh[1] = vh[1] & ~fm;
                sts_vec<2>(tb_s + i * TBS, vh);
            }
        };
        // text column i with base code at bits 31:30 of cw: Pv (column i+1) -> Nv (column i)
        auto text_column = [&](const RowSet<NW> &Pv, RowSet<NW> &Nv, const int i, const uint32_t cw, const bool TBCOL,
                               const uint32_t (&F)[NW]) {
            uint32_t pm[NW];
            const uint32_t *pmp = pm_s + (cw >> 30) * PMS;
            lds_vec<NW>(pmp, pm);
            uint32_t V = 0, H = 0;
            if (TBCOL) dc_column<NW, true>(Pv, Nv, pm, F, fm, XF, V, H);
            else dc_column<NW, false>(Pv, Nv, pm, F, fm, XF, V, H);
            if (TBCOL) {
                uint32_t vh[2];
                lds_vec<2>(tb_s + i * TBS, vh);
                vh[0] = (vh[0] & ~fm) | V;
                vh[1] = (vh[1] & ~fm) | H;
                sts_vec<2>(tb_s + i * TBS, vh);
            }
        };

        if constexpr (TMEM) __syncwarp();  // tcgen05.ld/st are warp-collective: reconverge after the per-lane setup
        if (uniform) {
            // fast path: every lane of the warp has a full text window; no per-column tests
            boundary_column(A, W);
            uint32_t Fa[NW], Fb[NW];
            ff_load(W - 1, Fa);
#pragma unroll
            for (int blk = NWIN - 1; blk >= 0; blk--) {
                uint32_t cw = tw[blk];
                const bool TBCOL = blk * 16 < TBCOLS;  // W=64: blocks 0,1 (columns 0..31); W=32: block 0
#pragma unroll 2
                for (int ii = 15; ii >= 1; ii -= 2) {
                    const int i = blk * 16 + ii;
                    ff_wait();
                    ff_load(i - 1, Fb);              // forefront of the next column is in flight during this one
                    text_column(A, B, i, cw, TBCOL, Fa);
                    ff_store(i, B.C[G - 1]);
                    cw <<= 2;
                    ff_wait();
                    ff_load(i >= 2 ? i - 2 : 0, Fa);
                    text_column(B, A, i - 1, cw, TBCOL, Fb);
                    ff_store(i - 1, A.C[G - 1]);
                    cw <<= 2;
                }
            }
            ff_wait();
        } else {
            // generic path (some lane's text is running out, n < W): per-lane start column.  Column i lives in
            // set A when i is even and in set B when it is odd, exactly as in the fast path, so a lane that
            // starts at its own boundary column n joins the alternation without any register copies.
            auto generic_column = [&](const RowSet<NW> &Pv, RowSet<NW> &Nv, const int i) {
                uint32_t F[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) F[k] = 0;
                if (i < W) {  // forefront access is warp-uniform (tcgen05.ld/st are collective), use of it is per lane
                    ff_load(i, F);
                    ff_wait();
                }
                if (i <= nn) {
                    if (i == nn) {
                        boundary_column(Nv, i);
                    } else {
                        const uint32_t word = i >= 48 ? tw[NWIN - 1] : (i >= 32 ? tw[NWIN > 2 ? 2 : 0] : (i >= 16 ? tw[1] : tw[0]));
                        text_column(Pv, Nv, i, word << (30 - 2 * (i & 15)), i < TBCOLS, F);
                    }
                }
                if constexpr (TMEM) __syncwarp();
                if (i < W) ff_store(i, Nv.C[G - 1]);
            };
            generic_column(B, A, W);
            for (int i = W - 1; i >= 1; i -= 2) {
                generic_column(A, B, i);
                generic_column(B, A, i - 1);
            }
        }
        if constexpr (TMEM) tmem_wait_st();  // the next phase reads this forefront back

        if (!have) continue;

        // ---- early termination: first row of the chunk whose sign bit is clear ------------------------
        int above = 0;  // rows of this chunk with the sign bit still set (monotone in r)
#pragma unroll
        for (int r = 0; r < G; r++) above += (int)(A.C[r][TOP] >> 31);
        if (above == G) {  // not within this chunk: continue with rows d0+G.. next phase
            d0 += G;
            continue;
        }
        entries += (uint64_t)(d0 + above + 1) * (uint64_t)(n + 1) + kWindowUnit;  // d_w = d0 + above
        d0 = 0;

        // ---- TB: walk the V/H words and the mismatch bits from (0,0) ----------------------------------
        // can_ins <=> V_i(J), can_del <=> H_i(J), else can_sub <=> pm[text[i]](J); priority I > D > X > '='
        // (src/genasm_cpu.cpp:321-370).  Branch-free per step: all lanes of the warp walk in lock step.  The
        // text position is carried by the column pointer, the pattern position by the one-hot mask; finished
        // runs are staged in this lane's (now dead) forefront slots, one word per run.
        // finished runs are staged per lane: in the dead forefront slots (smem variant) or in a byte array (TMEM variant);
        // `stage` advances by FFS per run in both
        auto stage_put = [&](const int at, const uint32_t run) {
            if constexpr (TMEM) stage_s[(at / FFS) * 32] = (uint8_t)run;
            else ff_s[at] = run;
        };
        auto stage_get = [&](const int at) -> uint32_t {
            if constexpr (TMEM) return stage_s[(at / FFS) * 32];
            else return ff_s[at];
        };
        const int jmax = m < TBL ? m : TBL;
        const uint32_t mask_end = 0x80000000u >> jmax;   // jmax <= W-O <= 31
        int tcol = 0;                                    // word offset of column i in the V/H array
        int stage = 0;                                   // word offset of the next staged run
        uint32_t mask = 0x80000000u;
        uint32_t tlo = tw[0], thi = tw[1];  // shifting text window, current base code = tlo & 3
        uint32_t prev = 0u, cnt = 0u;
        uint32_t vh[2], e;
        lds_vec<2>(tb_s, vh);
        e = pm_s[(tlo & 3u) * PMS + TOP];
        while (mask != mask_end && tcol != TBL * TBS) {
            const bool is_i = (vh[0] & mask) != 0;
            const bool has_d = (vh[1] & mask) != 0;
            const bool has_x = (e & mask) != 0;
            uint32_t op = has_x ? 1u : 0u;  // 0 '=', 1 'X', 2 'I', 3 'D'
            op = has_d ? 3u : op;
            op = is_i ? 2u : op;
            const bool brk = op != prev;
            if (brk && cnt != 0u) {
                stage_put(stage, prev * 64u + cnt);
                stage += FFS;
            }
            cnt = brk ? 1u : cnt + 1u;
            prev = op;
            if (!is_i) {
                tcol += TBS;
                tlo = __funnelshift_r(tlo, thi, 2);
                thi >>= 2;
            }
            if (is_i || !has_d) mask >>= 1;
            // next column's words (reloaded even when the column did not move: keeps the step branch-free)
            lds_vec<2>(tb_s + tcol, vh);
            e = pm_s[(tlo & 3u) * PMS + TOP];
        }
        if (cnt != 0u) {  // runs are flushed at window end, never merged across windows (quirk Q2)
            stage_put(stage, prev * 64u + cnt);
            stage += FFS;
        }
        const int i = tcol / TBS;
        const int j = __clz(mask);
        const uint32_t nb = (uint32_t)(stage / FFS);
        uint32_t edits = 0u;
        const bool fits = !want_cigar || (uint64_t)(out_end - out) >= (uint64_t)nb;
        for (int k = 0; k < stage; k += FFS) {
            const uint32_t run = stage_get(k);
            edits += run >= 64u ? (run & 63u) : 0u;   // every op but '=' is an edit
            if (want_cigar && fits) *out++ = (uint8_t)run;
        }
        if (!fits) overflow = true;
        nruns += nb;
        ed += edits;
        t_pos += (uint64_t)i;
        q_pos += (uint64_t)j;
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
    if constexpr (TMEM) {
        __syncwarp();
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncthreads();
        if (warp == 0) {
            const uint32_t base = tff & 0x0000FFFFu;  // warp 0: lane field is zero
            asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tff), "r"((uint32_t)L::TMEM_COLS));
            (void)base;
        }
    }
}
#endif  // !SG_SIM

}  // namespace sg
