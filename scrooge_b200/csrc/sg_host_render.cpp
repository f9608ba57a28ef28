// sg_host_render.cpp -- CIGAR text from packed runs on the host: "%d%c" per run (reference src/genasm_gpu.cu:881-888,
// src/genasm_cpu.cpp:387-403), 64 runs per step with AVX-512 VBMI/VBMI2 when the CPU has them.
//
// A packed run is one byte (op << 6) | count, count 1..63 (window configurations with W-O <= 63; longer runs use a
// continuation encoding and stay on the scalar path in sg_host_api.cu).  Its text is 2 characters (count < 10) or 3.
// The reference renders one alignment at a time through a stringstream (GPU path) or sprintf (CPU path); round 1 used a
// 256-entry table and one 4-byte store per run (~3 ns per run and thread: 130 ms for the 1.1 G runs of a 524 288 x 10 kbp
// call on 16 threads -- more than the alignment call itself).  Here 64 runs become three 64-byte vectors of
// (tens, ones, op) characters by byte permutes, the absent tens digits are squeezed out with VPCOMPRESSB, and the
// three pieces are stored with byte masks (nothing is written past an alignment's text: neighbouring alignments are
// rendered by other threads).  Compiled by g++ with per-function targets; the scalar path is the fallback.
#include <atomic>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <immintrin.h>

namespace {

struct RunTable {
    uint32_t text[256];
    uint8_t len[256];
    RunTable()
    {
        static const char ops[4] = {'=', 'X', 'I', 'D'};
        for (unsigned b = 0; b < 256; b++) {
            const unsigned c = b & 63u;
            char t[4] = {0, 0, 0, 0};
            unsigned l = 0;
            if (c >= 10) t[l++] = (char)('0' + c / 10);
            t[l++] = (char)('0' + c % 10);
            t[l++] = ops[b >> 6];
            memcpy(&text[b], t, 4);
            len[b] = (uint8_t)l;
        }
    }
};
const RunTable g_tab;

inline uint64_t len_scalar(const uint8_t *p, uint64_t cnt)
{
    uint64_t len = 0;
    for (uint64_t k = 0; k < cnt; k++) len += g_tab.len[p[k]];
    return len;
}

// exact: writes the text and nothing after it
inline char *render_scalar(const uint8_t *p, uint64_t cnt, char *o)
{
    if (!cnt) return o;
    for (uint64_t k = 0; k + 1 < cnt; k++) {   // a 4-byte store may spill 1-2 bytes into the next run's place: fine
        memcpy(o, &g_tab.text[p[k]], 4);
        o += g_tab.len[p[k]];
    }
    const uint8_t b = p[cnt - 1];
    memcpy(o, &g_tab.text[b], g_tab.len[b]);
    return o + g_tab.len[b];
}

#define SG_TGT __attribute__((target("avx512f,avx512bw,avx512vbmi,avx512vbmi2,bmi,bmi2,popcnt")))

SG_TGT uint64_t len_avx512(const uint8_t *p, uint64_t cnt)
{
    const __m512i m63 = _mm512_set1_epi8(63), nine = _mm512_set1_epi8(9);
    uint64_t three = 0, k = 0;
    for (; k + 64 <= cnt; k += 64) {
        const __m512i c = _mm512_and_si512(_mm512_loadu_si512((const void *)(p + k)), m63);
        three += (uint64_t)_mm_popcnt_u64(_mm512_cmpgt_epu8_mask(c, nine));
    }
    return 2 * k + three + len_scalar(p + k, cnt - k);
}

struct Consts {
    alignas(64) uint8_t tens[64], ones[64], ops[64];
    alignas(64) uint8_t idx_ab[3][64];   // output byte -> index into (tens | ones << 6) for vpermi2b
    alignas(64) uint8_t idx_c[3][64];    // output byte -> run index for the op character
    uint64_t is_c[3], is_a[3];           // output bytes that hold an op character / a tens digit
    int first_a[3];                      // run of the first tens position of each output vector
    Consts()
    {
        for (int c = 0; c < 64; c++) { tens[c] = (uint8_t)('0' + c / 10); ones[c] = (uint8_t)('0' + c % 10); ops[c] = (uint8_t)"=XID"[c & 3]; }
        for (int j = 0; j < 3; j++) {
            is_c[j] = is_a[j] = 0;
            first_a[j] = -1;
            for (int m = 0; m < 64; m++) {
                const int i = 64 * j + m, run = i / 3, which = i % 3;
                idx_ab[j][m] = (uint8_t)(which == 1 ? 64 + run : run);   // vpermi2b: bit 6 selects the second table
                idx_c[j][m] = (uint8_t)run;
                if (which == 2) is_c[j] |= 1ull << m;
                if (which == 0) { is_a[j] |= 1ull << m; if (first_a[j] < 0) first_a[j] = run; }
            }
        }
    }
};
const Consts g_c;

SG_TGT char *render_avx512(const uint8_t *p, uint64_t cnt, char *o)
{
    const __m512i m63 = _mm512_set1_epi8(63), nine = _mm512_set1_epi8(9);
    const __m512i t_tens = _mm512_load_si512((const void *)g_c.tens), t_ones = _mm512_load_si512((const void *)g_c.ones);
    const __m512i t_ops = _mm512_load_si512((const void *)g_c.ops);
    uint64_t k = 0;
    for (; k + 64 <= cnt; k += 64) {
        const __m512i b = _mm512_loadu_si512((const void *)(p + k));
        const __m512i c = _mm512_and_si512(b, m63);
        const __m512i A = _mm512_permutexvar_epi8(c, t_tens), B = _mm512_permutexvar_epi8(c, t_ones);
        const __m512i C = _mm512_permutexvar_epi8(_mm512_srli_epi16(_mm512_andnot_si512(m63, b), 6), t_ops);   // op = b >> 6 per byte
        const uint64_t has_tens = _mm512_cmpgt_epu8_mask(c, nine);
#pragma GCC unroll 3
        for (int j = 0; j < 3; j++) {
            __m512i v = _mm512_permutex2var_epi8(A, _mm512_load_si512((const void *)g_c.idx_ab[j]), B);
            v = _mm512_mask_permutexvar_epi8(v, g_c.is_c[j], _mm512_load_si512((const void *)g_c.idx_c[j]), C);
            // keep everything but the tens positions of runs with a one-digit count
            const uint64_t keep = ~g_c.is_a[j] | _pdep_u64(has_tens >> g_c.first_a[j], g_c.is_a[j]);
            const int n = (int)_mm_popcnt_u64(keep);
            _mm512_mask_storeu_epi8(o, _bzhi_u64(~0ull, (unsigned)n), _mm512_maskz_compress_epi8(keep, v));
            o += n;
        }
    }
    return render_scalar(p + k, cnt - k, o);
}

std::atomic<int> g_isa{-1};   // 1: AVX-512 VBMI2 path, 0: scalar (every thread that finds -1 computes the same value)

inline int isa()
{
    if (g_isa.load(std::memory_order_relaxed) < 0) {
        __builtin_cpu_init();
        int v = __builtin_cpu_supports("avx512vbmi2") && __builtin_cpu_supports("avx512vbmi") && __builtin_cpu_supports("avx512bw") &&
                __builtin_cpu_supports("bmi2");
        if (const char *e = getenv("SG_HOST_ISA")) if (atoi(e) < 2) v = 0;
        g_isa.store(v, std::memory_order_relaxed);
    }
    return g_isa.load(std::memory_order_relaxed);
}

}  // namespace

extern "C" {

// text length of cnt packed runs (counts 1..63)
uint64_t sg_host_runs_text_len(const uint8_t *runs, uint64_t cnt) { return isa() ? len_avx512(runs, cnt) : len_scalar(runs, cnt); }

// renders cnt packed runs at out (exactly sg_host_runs_text_len bytes are written); returns the end
char *sg_host_runs_render(const uint8_t *runs, uint64_t cnt, char *out) { return isa() ? render_avx512(runs, cnt, out) : render_scalar(runs, cnt, out); }

// The same into memory that will not be read again soon (the text blob of a whole batch: gigabytes): rendered into
// `scratch` first (cache resident; at least 3 * cnt + 192 bytes), then copied out with non-temporal stores for the whole
// 64-byte lines of the destination and plain stores for its edges -- a plain store to a line that is not in cache
// makes the core read it first, which for a 2.4 GB blob is 2.4 GB of DRAM traffic the host does not have to spare.
// The caller issues the store fence (sg_host_stream_fence) before other threads read the text.
__attribute__((target("avx512f"))) static char *stream_out(const char *s0, size_t len, char *out)
{
    // s0 has the same misalignment as out: whole lines of out are whole lines of s0
    const size_t mis = (uintptr_t)out & 63u;
    size_t i = mis ? (64 - mis < len ? 64 - mis : len) : 0;
    memcpy(out, s0, i);
    for (; i + 64 <= len; i += 64) _mm512_stream_si512((__m512i *)(out + i), _mm512_load_si512((const void *)(s0 + i)));
    memcpy(out + i, s0 + i, len - i);
    return out + len;
}

char *sg_host_runs_render_stream(const uint8_t *runs, uint64_t cnt, char *out, char *scratch)
{
    if (!isa()) return render_scalar(runs, cnt, out);
    char *base = (char *)(((uintptr_t)scratch + 63u) & ~(uintptr_t)63u);
    char *s0 = base + ((uintptr_t)out & 63u);
    char *send = render_avx512(runs, cnt, s0);
    return stream_out(s0, (size_t)(send - s0), out);
}

void sg_host_stream_fence(void) { _mm_sfence(); }

int sg_host_render_isa(void) { return isa(); }

}  // extern "C"
