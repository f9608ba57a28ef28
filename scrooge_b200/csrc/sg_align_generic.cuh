// sg_align_generic.cuh -- the alignment kernel for ANY window configuration (W, O) chosen at run time.
//
// The reference fixes the window size W and the overlap O at compile time (-DCLI_W/-DCLI_O, src/genasm_cpu.cpp:22-35,
// src/genasm_gpu.cu:1-63) and its evaluation sweeps both (scripts/profile.py:66-100 cpu_sweep_wo / cpu_sweep_o,
// :595-640 accuracy sweeps over W in 32, 64, 96, 128).  genasm_delta_kernel<W> (sg_align_delta.cuh) is tuned for the two
// configurations the reference ships -- 64/33 and 32/17 -- with everything unrolled around W - O = 31 or 15 traceback
// steps in one 32-bit word.  This kernel runs the same algorithm with W, O as kernel parameters:
//
//   * vectors of NW = ceil(W / 32) words (template parameter, 1..4), pattern position J at bit 32*NW-1-J: a window narrower
//     than the vector is a window with more padding rows, which the recurrence already handles (the last window of every
//     read has m < W);
//   * W columns per window whatever n is (the "matches nothing" mask stands in for the columns i >= n), one column =
//     delta_column<NW> as in the tuned kernel;
//   * the op planes A = V | H, B = ~V & (H | E) of the W-O+1 traceback columns kept for the top ceil((W-O)/32) words;
//   * the traceback as a plain per-lane loop with the run-length encoding done during the walk (the tuned kernel's
//     register-resident op streams assume at most 64 steps).
//
// Limits: 2 <= W <= 128, 0 <= O < W, W - O <= 63 (a run is one byte, (op << 6) | count, and a run can be W - O long).
// Same one-lane-per-alignment mapping, work queue and outputs as the tuned kernel; one warp per CTA, shared memory sized
// at launch.  Checked bit-exact against the unmodified reference built at nine further window
// configurations (tests/test_gpu_parity.py::test_window_configurations, goldens in tests/golden/golden_w*_o*.json) and
// against the tuned kernels at 64/33 and 32/17 (::test_generic_kernel_equals_tuned_kernels).
#pragma once
#include "sg_align_delta.cuh"

namespace sg {

struct GenericGeom {
    int W;      // window size in characters (columns per window)
    int TBL;    // W - O: traceback limit (src/genasm_cpu.cpp:50)
    int NWT;    // plane words kept per traceback column: pattern positions 0..TBL-1
};

__host__ __device__ inline int generic_smem_words(int NW, int W, int TBL)
{
    const int NWT = (TBL + 31) / 32;
    const int pm = 5 * NW * 32;                 // [base code 0..3, 4 = "matches nothing"][word][lane]
    const int tw = ((W + 15) / 16) * 32;        // [text word][lane]: the window's 2-bit codes
    const int tb = (TBL + 1) * NWT * 2 * 32;    // [column][word][plane][lane]
    return pm + tw + tb;
}

// NW-word addition with carry propagation for any NW
template <int NW>
__device__ __forceinline__ void add_vec_any(const uint32_t (&a)[NW], const uint32_t (&b)[NW], uint32_t (&s)[NW])
{
    uint32_t c = 0u;
#pragma unroll
    for (int k = 0; k < NW; k++) {
        const uint64_t r = (uint64_t)a[k] + (uint64_t)b[k] + (uint64_t)c;
        s[k] = (uint32_t)r;
        c = (uint32_t)(r >> 32);
    }
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

template <int NW>
__global__ void __launch_bounds__(32) genasm_generic_kernel(const AlignParams P, const GenericGeom G)
{
    constexpr int NTW = 2 * NW;                 // text / pattern words of a full-width window (16 bases each)
    extern __shared__ __align__(16) uint32_t smem_all[];
    const int lane = threadIdx.x & 31;
    const int W = G.W, TBL = G.TBL, NWT = G.NWT;
    uint32_t *pm_s = smem_all + lane;                                   // + (code * NW + k) * 32
    uint32_t *tw_s = smem_all + 5 * NW * 32 + lane;                     // + word * 32
    uint32_t *tb_s = smem_all + 5 * NW * 32 + ((W + 15) / 16) * 32 + lane;  // + ((column * NWT + kk) * 2 + plane) * 32

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
        for (int i = W - 1; i >= 0; i--) {
            const uint32_t cw = tw_s[(i >> 4) * 32];
            uint32_t code = (cw >> ((i & 15) * 2)) & 3u;
            if (i >= n) code = 4u;
            uint32_t pm[NW], Ph[NW];
#pragma unroll
            for (int k = 0; k < NW; k++) pm[k] = pm_s[(code * NW + k) * 32];
            delta_column_any<NW>(Pv, Mv, pm, Ph);
            if (i <= TBL) {
#pragma unroll
                for (int kk = 0; kk < NW; kk++) {
                    if (kk < NWT) {
                        const int k = NW - 1 - kk;
                        uint32_t *p = tb_s + ((i * NWT + kk) * 2) * 32;
                        p[0] = Pv[k] | Ph[k];
                        p[32] = ~Pv[k] & (Ph[k] | pm[k]);
                    }
                }
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

        // ---- TB + RLE (src/genasm_cpu.cpp:290-409): op = 2A + B = 0 '=', 1 'X', 2 'I', 3 'D' ----
        const int jmax = m < TBL ? m : TBL;
        int i = 0, j = 0;
        uint32_t cur_op = 4u, cur_cnt = 0u, edits = 0u;
        auto flush = [&]() {
            nruns++;
            if (want_cigar) {
                if (out < out_end) *out++ = (uint8_t)((cur_op << 6) | cur_cnt);
                else overflow = true;
            }
        };
        while (j < jmax && i < TBL) {
            const uint32_t *p = tb_s + ((i * NWT + (j >> 5)) * 2) * 32;
            const uint32_t bit = 0x80000000u >> (j & 31);
            const uint32_t op = ((p[0] & bit) ? 2u : 0u) | ((p[32] & bit) ? 1u : 0u);
            if (op != 2u) i++;
            if (op != 3u) j++;
            if (op != 0u) edits++;
            if (op != cur_op) {
                if (cur_cnt) flush();
                cur_op = op;
                cur_cnt = 1u;
            } else {
                cur_cnt++;
            }
        }
        if (cur_cnt) flush();   // runs end with their window (quirk Q2)
        t_pos += (uint64_t)i;
        q_pos += (uint64_t)j;
        ed += edits;
        if (q_pos >= q_end) {
            P.edit[pair] = ed;
            P.ref_consumed[pair] = t_pos - t_begin;
            P.nruns[pair] = nruns;
            P.status[pair] = overflow ? 5 : 0;
            if (P.dc_entries) P.dc_entries[pair] = entries & (kWindowUnit - 1);
            if (P.windows) P.windows[pair] = (uint32_t)(entries >> 40);
            have = false;
        }
    }
}

}  // namespace sg
