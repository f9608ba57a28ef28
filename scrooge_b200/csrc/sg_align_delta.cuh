// sg_align_delta.cuh -- the delta-encoded alignment kernel: same results as genasm_align_kernel (sg_align.cuh), bit
// for bit, with the window's distance calculation done COLUMN-WISE ON +-1 DELTAS instead of row-wise on K+1 threshold
// vectors.
//
// What the reference computes (src/genasm_cpu.cpp:210-288): R[d][i], d = 0..d_w, i = n..0, with
//     bit J of R[d][i] == 0   <=>   D(i,J) <= d,
// where D(i,J) is the window's semi-global edit-distance matrix
//     D(i,m) = 0,  D(n,J) = m-J,  D(i,J) = min( D(i+1,J+1) + [text[i] != pattern[J]],  D(i,J+1) + 1,  D(i+1,J) + 1 ).
// Everything the path returns is a function of D alone: the window distance d_w = D(0,0) (:278-283) and the
// traceback's tests (:321-343), which sg_align.cuh already evaluates on two words per text column,
//     V_i bit J  <=>  D(i,J+1) = D(i,J) - 1        H_i bit J  <=>  D(i+1,J) = D(i,J) - 1.
// The row-wise kernel builds V_i/H_i by OR-ing over d_w+1 (in practice 8 or 16) rows of R, 6 instructions per row
// and column.  But V_i and H_i ARE the vertical / horizontal "+1" delta vectors of D, and adjacent cells of D differ
// by -1, 0 or +1, so a whole column of D is two bit-vectors (Pv: +1, Mv: -1) that follow from the previous column
// with one addition and a handful of logic operations (Myers 1999, Hyyro 2001 -- here run from column n down to 0 on
// left-aligned vectors: pattern position J at bit W-1-J, padding bits below W-m behave as extra rows of zeros):
//     x   = (((Eq & Pv) + Pv) ^ Pv) | Eq              Eq = ~pm[text[i]]
//     Ph  = Mv | ~(x | Pv)          Mh = Pv & x       horizontal deltas column i+1 -> i       (H_i = Ph)
//     Pv' = (Mh << 1) | ~(Eq | Mv | (Ph << 1))        Mv' = (Ph << 1) & (Eq | Mv)             (V_i = Pv')
// That is 20 instructions per column at W=64 whatever the window distance is -- against 6 x 11.2 computed rows at
// 10 % error -- with no chunks, no forefront, no early-termination granularity, and d_w = popc(Pv) - popc(Mv) at
// column 0.  A text that runs out (n < W) needs no special column code either: a column whose character matches
// nothing (pm = all real bits set) maps the boundary state Pv = ones << (W-m), Mv = 0 onto itself, so lanes with
// n < W feed that fifth "code" to columns i >= n.
//
// The algorithmic work the roofline is quoted on stays the reference's: sum over windows of (d_w+1)(n+1) entries
// (SURVEY.md section 8d), counted per alignment exactly as before.
//
// Traceback.  For every traceback column the DC stores the op the walk would take at each pattern position as two bit
// planes, A = V | H and B = ~V & (H | E) (E = pm[text[i]]): op = 2A + B = 0 '=', 1 'X', 2 'I', 3 'D', which is the
// reference's priority I > D > X > '=' (src/genasm_cpu.cpp:346-370).  The walk is one LDS.64 and two bit tests per
// step, and the two op bits of step k go to bit k of two register-resident bit streams.  Run-length encoding, the edit
// count and the run count happen after the walk, on the streams, with bit tricks (run boundaries = hs ^ hs >> 1 |
// ls ^ ls >> 1, edits = popc(hs | ls)): one loop iteration per RUN instead of bookkeeping per step.
//
// Per-warp shared memory (W=64): pattern masks 5 x 256 B + traceback columns 32 x 256 B = 9.25 KB (four-warp CTAs of 37 KB: 24 warps per SM);
// every array [column][lane] so that all accesses are conflict free.
#pragma once
#include "sg_align.cuh"

// What was tried on this kernel and measured on a B200 against the version below (1 M x 10 kbp pairs, alignment kernel alone,
// +-0.01 ms run to run; profiles/r01_kernel_variants_ab.txt).  The code of the variants that lost is gone (git history has it):
//   * pattern masks of column i-1 fetched from shared memory while column i is computed: 37.0 -> 37.8 ms (the short-scoreboard
//     stalls the ncu source view shows at the first use of an LDS are already covered by the other warps of the scheduler, and
//     two more registers are live);
//   * the 64-bit addition and/or the two shifts of a column on the fma pipe (IMAD / IMAD.WIDE.U32 with run-time multipliers 1
//     and 2 that ptxas cannot turn back into alu-pipe instructions; a column drops from 18 to 15 alu-pipe instructions): both
//     39.8 ms, addition only 36.6 -> 39.8 ms, shifts only 36.6 -> 38.0 ms.  IMAD.WIDE costs about 6 issue cycles more than the
//     IADD3 it replaces;
//   * the next window's text and pattern words requested right after the traceback, so that their L2 latency is covered by the
//     run-length encoding: 37.0 -> 37.9 ms (69 registers instead of 62);
//   * fast traceback steps accumulating their stream bits (and the column address) with predicated IMADs instead of the 89
//     VIADDs per window the compiler emits: no change at all (36.59 vs 36.59 ms) -- VIADD does not compete with LOP3/SHF for the
//     alu pipe, and the traceback is not what that pipe waits for;
//   * a leaner run-emission loop (op bits rotated into place, one output pointer): 36.90 -> 36.59 ms, KEPT; the same loop
//     handling two runs per iteration: 36.46 -> 37.52 ms;
//   * the pattern bit planes of the window setup gathered by four bit selects per plane and 16-base word with plain shifts
//     (IMAD.SHL, the fma pipe's shift form) and one hand-placed LOP3 per select instead of the compiler's compress_even
//     (LEA.HI + LOP3 per stage, ~80 alu-pipe instructions per window): 36.64 -> 36.48 ms, KEPT; the same with the shifts as
//     multiplications by opaque constants: 36.58 -> 36.84 ms;
//   * the base code of a column brought to bits 31:30 by an IMAD and to the mask table's stride by one SHF + one IMAD, instead
//     of four masked copies per text word + one PRMT per column (16 instead of 23 alu-pipe instructions per 16 columns, 32 more
//     IMADs, 72 registers): 36.58 -> 37.62 ms.  A real multiplication is not a free ride on the fma pipe: both pipes issue once
//     per two cycles per scheduler; what is traded is issue slots, not pipe time;
//   * windows with fewer than W-O pattern characters taking the unchecked walk too, their op streams cut where the pattern ran
//     out (see the traceback): 150 bp reads 2.01 -> 2.04 G alignments/s at 64/33, 2.21 -> 2.25 at 32/17, KEPT;
//   * runs stored as whole words instead of bytes (EMIT below): 11.87 -> 11.65 ms per 303 104 pairs, 18.04 -> 12.74 ms when most
//     windows are mostly edits, KEPT as what the host API launches (profiles/r02_variant_ab.jsonl).

namespace sg {

#ifdef SG_STATS
__device__ unsigned long long g_delta_stats[4];   // warp iterations: uniform DC, generic DC, fast TB, generic TB
#endif

template <int W> struct DeltaLayout {
    static constexpr int NW = W / 32;
    static constexpr int TBL = W - WinCfg<W>::O;
    static constexpr int TBCOLS = TBL + 1;
    static constexpr int PMS = 64;                      // words between the masks of consecutive base codes (256 B: see the DC loop)
    static constexpr int PM_WORDS = 5 * PMS;            // [base code 0..3, 4 = "matches nothing"][lane][NW]
    static constexpr int TB_WORDS = TBCOLS * 2 * 32;    // [column][lane][A,B]: the traceback's op per pattern position, 2 bit planes
    static constexpr int WORDS_PER_WARP = PM_WORDS + TB_WORDS;
    static constexpr int BYTES_PER_WARP = WORDS_PER_WARP * 4;
    static constexpr int WARPS_PER_CTA = 4;             // 4 x 9.25 KB + 1 KB reserved = 38 KB: 6 CTAs = 24 warps per SM
    static constexpr int BYTES_PER_CTA = BYTES_PER_WARP * WARPS_PER_CTA;
};

// NW-word addition with carry propagation
template <int NW> __device__ __forceinline__ void add_vec(const uint32_t (&a)[NW], const uint32_t (&b)[NW], uint32_t (&s)[NW]);
template <> __device__ __forceinline__ void add_vec<1>(const uint32_t (&a)[1], const uint32_t (&b)[1], uint32_t (&s)[1]) { s[0] = a[0] + b[0]; }
template <> __device__ __forceinline__ void add_vec<2>(const uint32_t (&a)[2], const uint32_t (&b)[2], uint32_t (&s)[2])
{
    const uint64_t r = (((uint64_t)a[1] << 32) | a[0]) + (((uint64_t)b[1] << 32) | b[0]);
    s[0] = (uint32_t)r;
    s[1] = (uint32_t)(r >> 32);
}

// The host simulation (tests/sim, -DSG_SIM) compiles this file with g++: every inline-PTX block below has a C++ twin next
// to it, and SG_SIM_COUNT feeds the simulation's event counters (nothing in the product build).
#ifdef SG_SIM
#define SG_SIM_COUNT(k, v) (sim::counters[k] += (uint64_t)(v))
#else
#define SG_SIM_COUNT(k, v) ((void)0)
#endif

// two consecutive words at shared-window address addr
__device__ __forceinline__ void lds_pair(uint32_t addr, uint32_t &a, uint32_t &b)
{
#ifdef SG_SIM
    const uint32_t *p = reinterpret_cast<const uint32_t *>(sim::smem_base + addr);
    a = p[0];
    b = p[1];
#else
    asm volatile("ld.shared.v2.u32 {%0, %1}, [%2];" : "=r"(a), "=r"(b) : "r"(addr));
#endif
}

// Odd bits of x (bit 2k+1, k = 0..15) gathered into bits 16+k of the result; its low half is garbage.  x1 = x << 1.
// Each stage keeps the upper of two neighbouring groups where it is and takes the lower one from a shifted copy -- a
// bit select, so the garbage never carries into the payload -- and the shifted copies are plain shifts, which the compiler
// emits as IMAD.SHL (the fma pipe's shift form): 4 alu-pipe + 3 fma-pipe instructions.
__device__ __forceinline__ uint32_t bit_select(uint32_t a, uint32_t b, uint32_t m)   // (a & m) | (b & ~m) as ONE LOP3
{
#ifdef SG_SIM
    return (a & m) | (b & ~m);
#else
    uint32_t d;
    asm("lop3.b32 %0, %1, %2, %3, 0xE4;" : "=r"(d) : "r"(a), "r"(b), "r"(m));
    return d;
#endif
}
__device__ __forceinline__ uint32_t gather_odd_hi(uint32_t x, uint32_t x1)
{
    // plain shifts (the compiler makes them IMAD.SHL, the shift form of the fma pipe) and hand-placed LOP3s (left to
    // itself it splits every select into two LOP3 with simplified masks)
    uint32_t y = bit_select(x, x1, 0x88888888u);
    y = bit_select(y, y << 2, 0xC0C0C0C0u);
    y = bit_select(y, y << 4, 0xF000F000u);
    y = bit_select(y, y << 8, 0xFF000000u);
    return y;
}

// pattern_planes (sg_align.cuh) with gather_odd_hi: plane 1 = the odd bits of the 2-bit codes, plane 0 = the odd bits of
// the word shifted left by one
template <int NW>
__device__ __forceinline__ void pattern_planes_gather(const uint32_t (&pw)[2 * NW], uint32_t (&p0)[NW], uint32_t (&p1)[NW])
{
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t a = pw[2 * k], b = pw[2 * k + 1];
        const uint32_t a1 = a << 1, b1 = b << 1, a2 = a << 2, b2 = b << 2;
        const uint32_t ha = gather_odd_hi(a, a1), hb = gather_odd_hi(b, b1);
        const uint32_t la = gather_odd_hi(a1, a2), lb = gather_odd_hi(b1, b2);
        // pattern positions 32k..32k+31 go to word NW-1-k, bit-reversed
        p0[NW - 1 - k] = __brev(__byte_perm(la, lb, 0x7632));
        p1[NW - 1 - k] = __brev(__byte_perm(ha, hb, 0x7632));
    }
}

// One column: (Pv, Mv) of column i+1 -> column i, Ph = the horizontal +1 deltas between them.
template <int NW>
__device__ __forceinline__ void delta_column(uint32_t (&Pv)[NW], uint32_t (&Mv)[NW], const uint32_t (&pm)[NW], uint32_t (&Ph)[NW])
{
    uint32_t t[NW], s[NW], x[NW], Mh[NW], Phs[NW], Mhs[NW];
#pragma unroll
    for (int k = 0; k < NW; k++) t[k] = ~pm[k] & Pv[k];
    add_vec<NW>(t, Pv, s);
#pragma unroll
    for (int k = 0; k < NW; k++) {
        x[k] = (s[k] ^ Pv[k]) | ~pm[k];
        Ph[k] = Mv[k] | ~(x[k] | Pv[k]);
        Mh[k] = Pv[k] & x[k];
    }
    shl1<NW>(Ph, Phs);   // carry-in 0: D(i,m) = 0 for every column
    shl1<NW>(Mh, Mhs);
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint32_t xv = ~pm[k] | Mv[k];
        Pv[k] = Mhs[k] | ~(xv | Phs[k]);
        Mv[k] = Phs[k] & xv;
    }
}

// EMIT selects how the runs of a window reach the alignment's slab slot:
//   0  one byte store per run (the default: cheapest in instructions -- 6.4 runs per window at 10 % error);
//   1  runs collected in a register and stored as whole 32-bit words (one PRMT per run, one store per four runs, the
//      pending bytes carried across windows and flushed with the alignment's last window).  Needs slots that start and end
//      on 4-byte boundaries (SG_FLAG_SLOTS_ALIGNED4: a promise of the caller).  For windows that are mostly edits -- a
//      spurious candidate location of read mapping walks ~45 one-step runs per window -- the byte stores are what the
//      kernel waits for: every lane's store is a separate 32-byte sector write in L1 and L2 (32 wavefronts per
//      instruction), ~20 x 32 per window and warp against the ~1 000 cycles a window's arithmetic takes.
//      (The same with two registers and 64-bit stores -- two PRMTs per run, one store per eight runs -- was measured too:
//      11.60 ms where words take 11.65 on 10 kbp pairs, 1.630 against 1.613 ms on 150 bp reads at 32/17, no difference on
//      spurious candidates; not kept.)
template <int W, int EMIT = 0>
__global__ void __launch_bounds__(DeltaLayout<W>::WARPS_PER_CTA * 32) genasm_delta_kernel(const AlignParams P)
{
    using L = DeltaLayout<W>;
    constexpr int NW = L::NW;
    constexpr int NWIN = 2 * NW;
    constexpr int TBL = L::TBL;
    constexpr int TBCOLS = L::TBCOLS;
    constexpr int TOP = NW - 1;
    constexpr int PMS = L::PMS;    // words between the masks of consecutive base codes
    constexpr int TBS = 2 * 32;    // words between traceback columns

    extern __shared__ __align__(16) uint32_t smem_all[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    uint32_t *smem = smem_all + warp * L::WORDS_PER_WARP;
    uint32_t *pm_s = smem + lane * NW;
    uint32_t *tb_s = smem + L::PM_WORDS + lane * 2;

    const bool want_cigar = !(P.flags & 1u);
    const bool want_stats = P.dc_entries != nullptr || P.windows != nullptr;

    bool have = false, drained = false;
    uint64_t pair = 0, t_pos = 0, t_begin = 0, t_end = 0, q_pos = 0, q_end = 0;
    int64_t ed = 0;
    uint8_t *out = nullptr, *out_end = nullptr;
    uint32_t nruns = 0;
    uint32_t acc = 0u;             // EMIT == 1: the runs not yet stored, newest in the top byte
    uint64_t entries = 0;
    bool overflow = false;
    int n = -1, m = 0;
    uint32_t tw[NWIN], pw[NWIN];   // the window's text / pattern words
#pragma unroll
    for (int k = 0; k < NWIN; k++) { tw[k] = 0; pw[k] = 0; }

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
                have = true;
                break;
            }
        }
        if (__all_sync(0xFFFFFFFFu, !have)) break;

        // ---- window setup ------------------------------------------------------------------------------
        uint32_t Pv[NW], Mv[NW];
#pragma unroll
        for (int k = 0; k < NW; k++) { Pv[k] = 0; Mv[k] = 0; }
        n = -1;
        if (have) {
            uint64_t tl = t_end - t_pos, ql = q_end - q_pos;
            n = tl < (uint64_t)W ? (int)tl : W;
            m = ql < (uint64_t)W ? (int)ql : W;
            load_window<NWIN>(P.text, t_pos, tw);
            load_window<NWIN>(P.query, q_pos, pw);
            uint32_t p0[NW], p1[NW], hm[NW];
            pattern_planes_gather<NW>(pw, p0, p1);
            ones_shl<NW>(W - m, hm);
            // pm[c]: zero where pattern[J] == c (src/genasm_cpu.cpp:178-198) and in the W-m padding bits
            uint32_t m0[NW], m1[NW], m2[NW], m3[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) {
                m0[k] = (p1[k] | p0[k]) & hm[k];
                m1[k] = (p1[k] | ~p0[k]) & hm[k];
                m2[k] = (~p1[k] | p0[k]) & hm[k];
                m3[k] = (~p1[k] | ~p0[k]) & hm[k];
                Pv[k] = hm[k];   // boundary column D(n,J) = m-J (src/genasm_cpu.cpp:225-231): every vertical delta is +1
            }
            sts_vec<NW>(pm_s + 0 * PMS, m0);
            sts_vec<NW>(pm_s + 1 * PMS, m1);
            sts_vec<NW>(pm_s + 2 * PMS, m2);
            sts_vec<NW>(pm_s + 3 * PMS, m3);
            sts_vec<NW>(pm_s + 4 * PMS, hm);   // a character that matches nothing: columns i >= n
        }
        const bool uniform = __all_sync(0xFFFFFFFFu, !have || n == W);
        // The first TB_LIMIT steps of a traceback cannot end a window that holds at least TB_LIMIT pattern characters
        // (after k steps i <= k and j <= k), so they are walked without any end test.  A window with fewer pattern
        // characters (the last window of a read) walks them unchecked too: the only limit it can overrun is j == m, the
        // steps beyond that point read whatever the planes hold below the pattern (in bounds: at most TB_LIMIT columns
        // and TB_LIMIT mask shifts) and are cut off afterwards, where the m-th pattern-consuming step is found in the
        // op streams.  (Before, one short window among a warp's 32 sent the whole warp through the checked loop: with
        // 150 bp reads that was nearly every iteration.)
        const bool tb_fast = true;
#ifdef SG_STATS
        if (lane == 0) {
            atomicAdd(&g_delta_stats[uniform ? 0 : 1], 1ull);
            atomicAdd(&g_delta_stats[tb_fast ? 2 : 3], 1ull);
        }
#endif

        // ---- DC: columns W-1 .. 0, two delta vectors per lane ----------------------------------------------
        // lanes without work ride along on whatever their scratch holds; nothing of it is ever read
        // Code size matters here: the SM's instruction cache holds 32 KB and 24 warps at different points of the loop
        // body share it, so the 64 columns are NOT fully unrolled: two loops (columns without / with a traceback
        // store) of NWIN/2 iterations over one 16-column text word each.
        auto columns = [&](auto uni) {
            constexpr bool UNI = decltype(uni)::value;
            constexpr int HB = NWIN / 2;   // text words per half
            constexpr uint32_t NONE = 4u * PMS * 4u;   // byte offset of the "matches nothing" masks
            static_assert(TBCOLS == HB * 16, "traceback columns = the lower half of the window");
            const char *pmb = reinterpret_cast<const char *>(pm_s);
#pragma unroll
            for (int half = 1; half >= 0; half--) {
#pragma unroll 1
                for (int b = HB - 1; b >= 0; b--) {
                    uint32_t cw = tw[half * HB];
                    if (HB == 2 && b == 1) cw = tw[half * HB + (HB - 1)];
                    const int nrel = n - (half * HB + b) * 16;          // columns ii < nrel of this word hold text
                    uint32_t *tbp = tb_s + b * 16 * TBS;
                    // base codes as shared-memory offsets: four masked copies hold the codes of columns 4q+r in byte q,
                    // and one byte permute per column moves that byte to bits 15:8 (code * 256 B, the mask table's stride)
                    uint32_t cq[4];
                    cq[0] = cw & 0x03030303u;
                    cq[1] = (cw >> 2) & 0x03030303u;
                    cq[2] = (cw >> 4) & 0x03030303u;
                    cq[3] = (cw >> 6) & 0x03030303u;
#pragma unroll
                    for (int ii = 15; ii >= 0; ii--) {
                        uint32_t pm[NW], Ph[NW];
                        uint32_t off = __byte_perm(cq[ii & 3], 0u, 0x4404u | ((uint32_t)(ii >> 2) << 4));
                        if (!UNI) off = ii < nrel ? off : NONE;
                        lds_vec<NW>(reinterpret_cast<const uint32_t *>(pmb + off), pm);
                        delta_column<NW>(Pv, Mv, pm, Ph);
                        if (half == 0) {
                            const uint32_t v = Pv[TOP], hh = Ph[TOP], e = pm[TOP];
                            *reinterpret_cast<uint2 *>(tbp + ii * TBS) = make_uint2(v | hh, ~v & (hh | e));
                        }
                    }
                }
            }
        };
        if (uniform) columns(std::true_type{});
        else columns(std::false_type{});

        if (!have) continue;

        if (want_stats) {   // window distance d_w = D(0,0) = sum of the vertical deltas of column 0 (src/genasm_cpu.cpp:278-283);
                            // nothing the path returns needs it: only the work counters of the measurement interface do
            int dw = 0;
#pragma unroll
            for (int k = 0; k < NW; k++) dw += __popc(Pv[k]) - __popc(Mv[k]);
            entries += (uint64_t)(dw + 1) * (uint64_t)(n + 1) + kWindowUnit;
        }

        // ---- TB: walk the op planes from (0,0); the op of step k goes to bit k of two bit streams (hi, lo) ----
        const int jmax = m < TBL ? m : TBL;
        const uint32_t mask_end = 0x80000000u >> jmax;   // jmax <= W-O <= 31
        constexpr int SW = (2 * TBL + 31) / 32;          // stream words: at most 2*TB_LIMIT steps, 32 per word
        uint32_t tcol = (uint32_t)__cvta_generic_to_shared(tb_s);   // shared-memory address of column i
        const uint32_t tb_begin = tcol, tb_end = tcol + TBL * TBS * 4;
        uint32_t mask = 0x80000000u;                     // pattern position j, one-hot from the top
        uint32_t hs[SW], ls[SW];
        uint32_t ca, cb;
        lds_pair(tcol, ca, cb);
        uint32_t h0 = 0u, l0 = 0u, bit0 = 1u;
        if (tb_fast) {
#pragma unroll
            for (int k = 0; k < TBL; k++) {
                // one step, written out as predicated instructions (the compiler's version spends selects on them):
                //   hi/lo = the op's two bits; every op but 'I' (hi & !lo) consumes a text character (next column),
                //   every op but 'D' (hi & lo) consumes a pattern character (mask >>= 1)
#if defined(SG_SIM)
                {
                    const bool hi = (ca & mask) != 0u, lo = (cb & mask) != 0u;
                    if (hi) h0 |= 1u << k;
                    if (lo) l0 |= 1u << k;
                    if (!(hi && !lo)) tcol += TBS * 4;
                    if (!(hi && lo)) mask >>= 1;
                    lds_pair(tcol, ca, cb);
                }
#else
                asm volatile(
                    "{\n\t"
                    ".reg .pred ph, pl, pi, pd;\n\t"
                    ".reg .b32 t;\n\t"
                    "and.b32 t, %2, %4;\n\t"
                    "setp.ne.u32 ph, t, 0;\n\t"
                    "and.b32 t, %3, %4;\n\t"
                    "setp.ne.u32 pl, t, 0;\n\t"
                    "@ph or.b32 %0, %0, %6;\n\t"
                    "@pl or.b32 %1, %1, %6;\n\t"
                    "and.pred pd, ph, pl;\n\t"
                    "not.pred pi, pl;\n\t"
                    "and.pred pi, pi, ph;\n\t"
                    "@!pi add.u32 %5, %5, %7;\n\t"
                    "@!pd shr.u32 %4, %4, 1;\n\t"
                    "ld.shared.v2.u32 {%2, %3}, [%5];\n\t"
                    "}"
                    : "+r"(h0), "+r"(l0), "+r"(ca), "+r"(cb), "+r"(mask), "+r"(tcol)
                    : "r"(1u << k), "n"(TBS * 4));
#endif
            }
            bit0 = TBL < 32 ? 1u << (TBL & 31) : 0u;
            if (m < TBL) {
                constexpr uint32_t kWalked = TBL < 32 ? (1u << (TBL & 31)) - 1u : 0xFFFFFFFFu;
                const uint32_t nd = ~(h0 & l0) & kWalked;          // steps that consumed a pattern character (all but 'D')
                if (__popc(nd) >= m) {
                    // the pattern ran out within these steps: the position of the m-th set bit of nd, by halving
                    uint32_t pos = 0u;
                    int r = m;
#pragma unroll
                    for (int half = 16; half >= 1; half >>= 1) {
                        const int c = __popc((nd >> pos) & ((1u << half) - 1u));
                        if (c < r) { r -= c; pos += (uint32_t)half; }
                    }
                    const uint32_t keep = (2u << pos) - 1u;        // steps 0..pos are the walk (pos <= TBL-1 <= 30)
                    h0 &= keep;
                    l0 &= keep;
                    const int took_text = __popc(~(h0 & ~l0) & keep);   // every step but an 'I' consumed a text character
                    tcol = tb_begin + (uint32_t)took_text * (uint32_t)(TBS * 4);
                    mask = mask_end;                               // j == m: the checked loop below has nothing left to do
                }
                // else: fewer than m pattern characters consumed so far (deletions): the state is that of a checked walk
                // after TB_LIMIT steps, and the loop below goes on from it
            }
        }
#pragma unroll
        for (int w = 0; w < SW; w++) {
            uint32_t h = w == 0 ? h0 : 0u, l = w == 0 ? l0 : 0u;
            uint32_t bit = w == 0 ? bit0 : 1u;
            if (bit != 0u && mask != mask_end && tcol != tb_end) {
                do {
                    const bool hi = (ca & mask) != 0u;
                    const bool lo = (cb & mask) != 0u;
                    if (hi) h |= bit;
                    if (lo) l |= bit;
                    bit <<= 1;
                    if (!(hi && !lo)) tcol += TBS * 4;
                    if (!(hi && lo)) mask >>= 1;
                    lds_pair(tcol, ca, cb);
                } while (bit != 0u && mask != mask_end && tcol != tb_end);
            }
            hs[w] = h;
            ls[w] = l;
        }
        const int i = (int)((tcol - tb_begin) / (TBS * 4));
        const int j = __clz(mask);
        t_pos += (uint64_t)i;
        q_pos += (uint64_t)j;

        // ---- RLE on the streams: per-window runs, flushed at window end, never merged across windows (quirk Q2) ----
        // a run ends at step k when op k+1 differs (the streams are zero beyond the last step) and at the last step
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
            nb += __popc(e[w]);
        }
        const bool fits = !want_cigar || (uint64_t)(out_end - out) >= (uint64_t)nb;
        if (!fits) overflow = true;
        if (want_cigar && fits) {
            // one byte per run, (op << 6) | length.  The op bits of step p are brought to bits 7 and 6 by one rotation each
            // (the streams are pre-rotated by 7 and 6 per word), merged by one LOP3 and joined with the length by another.
            uint8_t *o = out;
            out += nb;                                   // the run count is known: the loop carries one pointer only
            int st = -1;                                 // step before the current run's first, relative to word w
#pragma unroll
            for (int w = 0; w < SW; w++) {
                uint32_t ew = e[w];
                const uint32_t h7 = __funnelshift_l(hs[w], hs[w], 7), l6 = __funnelshift_l(ls[w], ls[w], 6);
                while (ew) {
                    const int p = __ffs((int)ew) - 1;
                    const uint32_t rh = __funnelshift_r(h7, h7, p), rl = __funnelshift_r(l6, l6, p);
                    const uint32_t t = (rh & 0x80u) | (rl & ~0x80u);
                    if constexpr (EMIT == 1) {
                        // the run's byte enters the accumulator from the top (bytes 3..1 move down one place); the slot is
                        // 4-byte aligned, so the pointer's low bits say how many bytes are pending, and the fourth completes a word
                        acc = __byte_perm(acc, (t & 0xC0u) | (uint32_t)(p - st), 0x4321);
                        o++;
                        if (((uint32_t)(uintptr_t)o & 3u) == 0u) {
                            *reinterpret_cast<uint32_t *>(o - 4) = acc;
                            SG_SIM_COUNT(1, 1);
                        }
                    } else {
                        *o++ = (uint8_t)((t & 0xC0u) | (uint32_t)(p - st));
                        SG_SIM_COUNT(0, 1);
                    }
                    st = p;
                    ew &= ew - 1u;
                }
                st -= 32;
            }
        }
        nruns += nb;
        ed += edits;
        if (q_pos >= q_end) {
            if constexpr (EMIT == 1) {
                // the pending 1..3 runs: moved down to the low bytes and stored as a word that ends inside the slot (its end is
                // 4-byte aligned); the bytes above them are padding nobody reads
                const uint32_t pend = (uint32_t)(uintptr_t)out & 3u;
                if (want_cigar && pend != 0u) {
                    *reinterpret_cast<uint32_t *>(out - pend) = acc >> (32u - 8u * pend);
                    SG_SIM_COUNT(1, 1);
                }
            }
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
