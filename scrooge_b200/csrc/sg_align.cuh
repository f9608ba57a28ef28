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
//    column step is, per 32-bit word and row: one funnel shift, one AND, two LOP3
//        X[r]   = P[r] & sP[r]                          (P = previous column, sP = P << 1)
//        C[r]   = ((sP[r] | pm) & X[r-1]) & sC[r-1]      (= mat & del & sub & ins of src/genasm_cpu.cpp:247-251)
//        sC[r]  = C[r] << 1
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
//    where D(i,J) = min{d : bit J of R[d][i] is 0} is the edit-distance matrix the R rows encode.  With
//    E_i = pm[text[i]] (mismatch bits) this is 12 bytes per column, 384 B per alignment, independent
//    of the window distance.  The traceback's tests (src/genasm_cpu.cpp:321-343)
//        can_ins = d>0 && zero(R[d-1][i],   J+1)  <=> D(i,J+1)   <= d-1
//        can_del = d>0 && zero(R[d-1][i+1], J)    <=> D(i+1,J)   <= d-1
//        can_sub = d>0 && zero(R[d-1][i+1], J+1)  <=> D(i+1,J+1) <= d-1
//    are evaluated with d == D(i,J) (an invariant of the walk from d_w = D(0,0)), so can_ins <=> V_i(J),
//    can_del <=> H_i(J), and when neither holds can_sub <=> text[i] != pattern[J] <=> E_i(J) by the DP
//    recurrence.  The j == m-1 special case and the i >= n limit fall out of the same bits.
//
// Per-warp shared memory (W=64): pattern masks 1 KB + forefront 16.25 KB + V/H/E 12 KB = 29.25 KB,
// every array laid out [column][lane] so that all accesses are bank-conflict free whatever column each
// lane is at.
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>

namespace sg {

constexpr int kRowsPerChunk = 8;  // G

template <int W> struct WinCfg;
template <> struct WinCfg<64> { static constexpr int O = 33; };
template <> struct WinCfg<32> { static constexpr int O = 17; };

template <int W> struct SmemLayout {
    static constexpr int NW = W / 32;               // 32-bit words per bitvector
    static constexpr int TBL = W - WinCfg<W>::O;    // TB_LIMIT (src/genasm_cpu.cpp:50)
    static constexpr int TBCOLS = TBL + 1;          // traceback columns kept (0..TBL)
    static constexpr int PM_WORDS = 4 * NW * 32;
    static constexpr int FF_WORDS = (W + 1) * NW * 32;
    static constexpr int TB_WORDS = 3 * TBCOLS * 32;
    static constexpr int WORDS_PER_WARP = PM_WORDS + FF_WORDS + TB_WORDS;
    static constexpr int BYTES_PER_WARP = WORDS_PER_WARP * 4;
};

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

// ---- the kernel ------------------------------------------------------------------------------------

template <int W>
__global__ void __launch_bounds__(32) genasm_align_kernel(const AlignParams P)
{
    using L = SmemLayout<W>;
    constexpr int NW = L::NW;
    constexpr int NWIN = 2 * NW;  // words of a 2-bit window
    constexpr int TBL = L::TBL;
    constexpr int TBCOLS = L::TBCOLS;
    constexpr int G = kRowsPerChunk;
    constexpr int TOP = NW - 1;

    extern __shared__ uint32_t smem[];
    const int lane = threadIdx.x;
    // [c][lane][NW], [col][lane][NW], [3][col][lane]
    uint32_t *pm_s = smem + lane * NW;
    uint32_t *ff_s = smem + L::PM_WORDS + lane * NW;
    uint32_t *tbv_s = smem + L::PM_WORDS + L::FF_WORDS + lane;
    uint32_t *tbh_s = tbv_s + TBCOLS * 32;
    uint32_t *tbe_s = tbh_s + TBCOLS * 32;

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
                uint64_t ql = P.query_len[idx];
                if (ql == 0) {  // zero windows: distance 0, empty CIGAR (src/tests.cu:243,246)
                    P.edit[idx] = 0;
                    P.ref_consumed[idx] = 0;
                    P.nruns[idx] = 0;
                    P.status[idx] = 0;
                    if (P.dc_entries) P.dc_entries[idx] = 0;
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
#pragma unroll
            for (int k = 0; k < NW; k++) {
                pm_s[0 * NW * 32 + k] = (p1[k] | p0[k]) & hm[k];
                pm_s[1 * NW * 32 + k] = (p1[k] | ~p0[k]) & hm[k];
                pm_s[2 * NW * 32 + k] = (~p1[k] | p0[k]) & hm[k];
                pm_s[3 * NW * 32 + k] = (~p1[k] | ~p0[k]) & hm[k];
            }
        }
        const int nn = have ? n : -1;
        const uint32_t fm = d0 == 0 ? 0xFFFFFFFFu : 0u;  // "no row above this chunk"

        // ---- DC: one chunk of G rows, columns n .. 0 -------------------------------------------------
        uint32_t C[G][NW], S[G][NW];   // previous column entries and their << 1
        uint32_t XFp[NW];               // F & (F << 1) of row d0-1 at the previous column
#pragma unroll
        for (int k = 0; k < NW; k++) XFp[k] = 0xFFFFFFFFu;
#pragma unroll
        for (int r = 0; r < G; r++)
#pragma unroll
            for (int k = 0; k < NW; k++) { C[r][k] = 0; S[r][k] = 0; }

        // One column step.  i is the text column, c the base code of text[i] (unused on the boundary column),
        // TBCOL says at compile time whether the column can be visited by the traceback.
        auto column = [&](const int i, const uint32_t c, const bool TBCOL) {
            uint32_t *ffp = ff_s + i * (NW * 32);
            uint32_t F[NW], sF[NW];
            lds_vec<NW>(ffp, F);
#pragma unroll
            for (int k = 0; k < NW; k++) F[k] |= fm;
            shl1<NW>(F, sF);
            uint32_t V = 0, H = 0, E = 0;
            if (i == nn) {
                // boundary column: R[d][n] = ones << d  (src/genasm_cpu.cpp:225-231,239-245), left-aligned
#pragma unroll
                for (int r = 0; r < G; r++) {
                    ones_shl<NW>(W - m + d0 + r, C[r]);
                    shl1<NW>(C[r], S[r]);
                    V |= C[r][TOP] & ~S[r][TOP];
                }
            } else {
                uint32_t pm[NW];
                lds_vec<NW>(pm_s + c * (NW * 32), pm);
                E = pm[TOP];
                uint32_t aboveS[NW], aboveX[NW];
#pragma unroll
                for (int k = 0; k < NW; k++) { aboveS[k] = sF[k] | fm; aboveX[k] = XFp[k]; }
#pragma unroll
                for (int r = 0; r < G; r++) {
                    uint32_t Xr[NW], newC[NW];
#pragma unroll
                    for (int k = 0; k < NW; k++) {
                        Xr[k] = C[r][k] & S[r][k];
                        newC[k] = ((S[r][k] | pm[k]) & aboveX[k]) & aboveS[k];
                    }
                    if (TBCOL) H |= newC[TOP] & ~C[r][TOP];
#pragma unroll
                    for (int k = 0; k < NW; k++) { C[r][k] = newC[k]; aboveX[k] = Xr[k]; }
                    shl1<NW>(C[r], S[r]);
                    if (TBCOL) V |= C[r][TOP] & ~S[r][TOP];
#pragma unroll
                    for (int k = 0; k < NW; k++) aboveS[k] = S[r][k];
                }
            }
            // row d0-1 of this column is "topright" for the next column
#pragma unroll
            for (int k = 0; k < NW; k++) XFp[k] = (F[k] & sF[k]) | fm;
            sts_vec<NW>(ffp, C[G - 1]);
            if (TBCOL) {
                const int o = i * 32;
                tbv_s[o] = (tbv_s[o] & ~fm) | V;
                tbh_s[o] = (tbh_s[o] & ~fm) | H;
                if (fm) tbe_s[o] = E;
            }
        };

        if (nn == W) column(W, 0u, false);
#pragma unroll
        for (int blk = NWIN - 1; blk >= 0; blk--) {
            const uint32_t word = tw[blk];
            const bool TBCOL = (blk * 16 + 15) < TBCOLS;
#pragma unroll 2
            for (int ii = 15; ii >= 0; ii--) {
                const int i = blk * 16 + ii;
                if (i > nn) continue;
                const uint32_t c = (word >> (ii * 2)) & 3u;
                column(i, c, TBCOL);
            }
        }

        if (!have) continue;

        // ---- early termination: first row of the chunk whose sign bit is clear ------------------------
        int above = 0;  // rows of this chunk with the sign bit still set (monotone in r)
#pragma unroll
        for (int r = 0; r < G; r++) above += (int)(C[r][TOP] >> 31);
        if (above == G) {  // not within this chunk: continue with rows d0+G.. next phase
            d0 += G;
            continue;
        }
        entries += (uint64_t)(d0 + above + 1) * (uint64_t)(n + 1);  // d_w = d0 + above
        d0 = 0;

        // ---- TB: walk the V/H/E words from (0,0) ------------------------------------------------------
        int i = 0, j = 0;
        uint32_t mask = 0x80000000u;
        uint32_t cur_op = 4u, cur_cnt = 0u, edits = 0u;
        while (j < m && i < TBL && j < TBL) {
            const uint32_t v = tbv_s[i * 32], h = tbh_s[i * 32], e = tbe_s[i * 32];
            uint32_t op;  // 0 '=', 1 'X', 2 'I', 3 'D'; priority I > D > X > '=' (src/genasm_cpu.cpp:346-370)
            if (v & mask) op = 2u;
            else if (h & mask) op = 3u;
            else if (e & mask) op = 1u;
            else op = 0u;
            if (op != 2u) i++;
            if (op != 3u) { j++; mask >>= 1; }
            if (op != 0u) edits++;
            if (op != cur_op) {
                if (cur_cnt) {
                    if (want_cigar) {
                        if (out < out_end) *out++ = (uint8_t)((cur_op << 6) | cur_cnt);
                        else overflow = true;
                    }
                    nruns++;
                }
                cur_op = op;
                cur_cnt = 1u;
            } else {
                cur_cnt++;
            }
        }
        if (cur_cnt) {  // runs are flushed at window end, never merged across windows (quirk Q2)
            if (want_cigar) {
                if (out < out_end) *out++ = (uint8_t)((cur_op << 6) | cur_cnt);
                else overflow = true;
            }
            nruns++;
        }
        ed += edits;
        t_pos += (uint64_t)i;
        q_pos += (uint64_t)j;
        if (q_pos >= q_end) {
            P.edit[pair] = ed;
            P.ref_consumed[pair] = t_pos - t_begin;
            P.nruns[pair] = nruns;
            P.status[pair] = overflow ? 5 : 0;
            if (P.dc_entries) P.dc_entries[pair] = entries;
            have = false;
        }
    }
}

}  // namespace sg
