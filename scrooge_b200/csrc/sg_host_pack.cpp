// sg_host_pack.cpp -- host-side ASCII -> 2 bit/base packing (same layout as pack_2bit_kernel: 16 bases per
// little-endian 32-bit word, base k in bits 2k+1:2k, A0 C1 G2 T3, case-insensitive), multi-threaded, AVX-512 when the
// CPU has it.  Used by the host API to shrink the PCIe upload 4x when the upload, not the GPU, is the bottleneck
// (SURVEY.md section 8f-2).  Compiled by g++ (function multiversioning), not nvcc.
#include <cstdint>
#include <cstring>
#include <algorithm>
#include <atomic>
#include <immintrin.h>
#include <omp.h>

namespace {

// scalar reference: also handles heads/tails of the vector paths.  Returns the first bad position or UINT64_MAX.
uint64_t pack_scalar(const uint8_t *a, uint64_t n, uint8_t *out /* n/4 bytes, n % 4 == 0 */)
{
    uint64_t bad = ~0ull;
    for (uint64_t i = 0; i < n; i += 4) {
        uint32_t byte = 0;
        for (int k = 0; k < 4; k++) {
            const uint32_t c = a[i + k] & 0xDFu;
            const uint32_t code = ((c >> 1) ^ (c >> 2)) & 3u;
            if (c != (uint32_t)"ACGT"[code] && bad == ~0ull) bad = i + k;
            byte |= code << (2 * k);
        }
        out[i / 4] = (uint8_t)byte;
    }
    return bad;
}

// 64 bases -> 16 bytes.  code = ((c >> 1) ^ (c >> 2)) & 3 after folding the case; the alphabet check maps the code back
// to its letter (byte shuffle) and compares.
__attribute__((target("avx512f,avx512bw"))) static inline __m128i pack64_avx512(const uint8_t *p, __mmask64 *bad)
{
    const __m512i m03 = _mm512_set1_epi8(0x03), mdf = _mm512_set1_epi8((char)0xDF);
    const __m512i letters = _mm512_set4_epi32(0, 0, 0, 0x54474341);        // per 128-bit lane: bytes 0..3 = "ACGT"
    const __m512i w14 = _mm512_set1_epi16(0x0401), w116 = _mm512_set1_epi32(0x00100001);
    const __m512i x = _mm512_loadu_si512((const void *)p);
    const __m512i u = _mm512_and_si512(x, mdf);
    const __m512i t = _mm512_and_si512(_mm512_xor_si512(_mm512_srli_epi16(u, 1), _mm512_srli_epi16(u, 2)), m03);
    *bad |= _mm512_cmpneq_epi8_mask(_mm512_shuffle_epi8(letters, t), u);
    const __m512i p16 = _mm512_maddubs_epi16(t, w14);   // c0 + 4*c1 per 16-bit lane
    const __m512i p32 = _mm512_madd_epi16(p16, w116);   // + 16*(c2 + 4*c3) per 32-bit lane
    return _mm512_cvtepi32_epi8(p32);
}

// A thread of this loop is bound by the latency of its own cache misses (about 6.5 GB/s of ASCII per thread with plain
// loads and stores, measured), not by the arithmetic: the input is prefetched 4 KB ahead into L2 and, when the output
// is 64-byte aligned, whole output lines are written with non-temporal stores (no read-for-ownership of a buffer that
// only the copy engine will read).  +20..40 % per thread on the Xeon hosts measured.
__attribute__((target("avx512f,avx512bw"))) uint64_t pack_avx512(const uint8_t *a, uint64_t n, uint8_t *out)
{
    // n is a multiple of 64
    uint64_t first_bad_block = ~0ull, i = 0;
    const bool nt = (((uintptr_t)out) & 63u) == 0;
    for (; i + 256 <= n; i += 256) {
        _mm_prefetch((const char *)(a + i + 4096), _MM_HINT_T1);
        _mm_prefetch((const char *)(a + i + 4096 + 64), _MM_HINT_T1);
        _mm_prefetch((const char *)(a + i + 4096 + 128), _MM_HINT_T1);
        _mm_prefetch((const char *)(a + i + 4096 + 192), _MM_HINT_T1);
        __mmask64 bad = 0;
        const __m128i r0 = pack64_avx512(a + i, &bad), r1 = pack64_avx512(a + i + 64, &bad);
        const __m128i r2 = pack64_avx512(a + i + 128, &bad), r3 = pack64_avx512(a + i + 192, &bad);
        __m512i z = _mm512_castsi128_si512(r0);
        z = _mm512_inserti32x4(z, r1, 1);
        z = _mm512_inserti32x4(z, r2, 2);
        z = _mm512_inserti32x4(z, r3, 3);
        if (nt) _mm512_stream_si512((__m512i *)(out + i / 4), z);
        else _mm512_storeu_si512((void *)(out + i / 4), z);
        if (bad && first_bad_block == ~0ull) first_bad_block = i;
    }
    for (; i < n; i += 64) {
        __mmask64 bad = 0;
        _mm_storeu_si128((__m128i *)(out + i / 4), pack64_avx512(a + i, &bad));
        if (bad && first_bad_block == ~0ull) first_bad_block = i;
    }
    if (nt) _mm_sfence();   // the copy engine reads this buffer next
    if (first_bad_block == ~0ull) return ~0ull;
    for (uint64_t k = first_bad_block; k < std::min(n, first_bad_block + 256); k++) {
        const uint8_t c = a[k] & 0xDF;
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return k;
    }
    return first_bad_block;
}

__attribute__((target("avx2"))) uint64_t pack_avx2(const uint8_t *a, uint64_t n, uint8_t *out)
{
    // n is a multiple of 32
    const __m256i m03 = _mm256_set1_epi8(0x03), mdf = _mm256_set1_epi8((char)0xDF);
    const __m256i letters = _mm256_set_epi32(0, 0, 0, 0x54474341, 0, 0, 0, 0x54474341);
    const __m256i w14 = _mm256_set1_epi16(0x0401), w116 = _mm256_set1_epi32(0x00100001);
    const __m256i gather = _mm256_set_epi8(-1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0,
                                           -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, -1, 12, 8, 4, 0);
    uint64_t first_bad_block = ~0ull;
    for (uint64_t i = 0; i < n; i += 32) {
        const __m256i x = _mm256_loadu_si256((const __m256i *)(a + i));
        const __m256i u = _mm256_and_si256(x, mdf);
        const __m256i t = _mm256_and_si256(_mm256_xor_si256(_mm256_srli_epi16(u, 1), _mm256_srli_epi16(u, 2)), m03);
        const int ok = _mm256_movemask_epi8(_mm256_cmpeq_epi8(_mm256_shuffle_epi8(letters, t), u));
        if (ok != -1 && first_bad_block == ~0ull) first_bad_block = i;
        const __m256i p32 = _mm256_madd_epi16(_mm256_maddubs_epi16(t, w14), w116);
        const __m256i g = _mm256_shuffle_epi8(p32, gather);  // 4 bytes at the bottom of each 128-bit lane
        uint32_t lo = (uint32_t)_mm256_extract_epi32(g, 0), hi = (uint32_t)_mm256_extract_epi32(g, 4);
        memcpy(out + i / 4, &lo, 4);
        memcpy(out + i / 4 + 4, &hi, 4);
    }
    if (first_bad_block == ~0ull) return ~0ull;
    for (uint64_t k = first_bad_block; k < first_bad_block + 32; k++) {
        const uint8_t c = a[k] & 0xDF;
        if (c != 'A' && c != 'C' && c != 'G' && c != 'T') return k;
    }
    return first_bad_block;
}

std::atomic<int> g_isa{-1};  // 2 avx512, 1 avx2, 0 scalar (every thread that finds -1 computes the same value)

}  // namespace

extern "C" {

// Packs n_bases ASCII characters into packed[0 .. ceil(n/16)) 32-bit words (+ the caller's padding words are left
// untouched).  Bases past n in the last word are zero.  Returns the smallest offending position or UINT64_MAX.
uint64_t sg_host_pack_2bit(const char *ascii, uint64_t n_bases, uint32_t *packed, int threads)
{
    if (g_isa.load(std::memory_order_relaxed) < 0) {
        __builtin_cpu_init();
        int isa = __builtin_cpu_supports("avx512bw") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
        if (const char *e = getenv("SG_HOST_ISA")) isa = std::min(isa, atoi(e));
        g_isa.store(isa, std::memory_order_relaxed);
    }
    const uint8_t *a = (const uint8_t *)ascii;
    uint8_t *out = (uint8_t *)packed;
    const uint64_t whole = n_bases & ~63ull;  // handled by the vector paths in 64-base blocks
    if (threads < 1) threads = omp_get_max_threads();
    const uint64_t nblk = whole / 64;
    uint64_t bad = ~0ull;
    // chunks of 16 K blocks (1 MiB of ASCII) keep every thread streaming through its own pages
    const uint64_t chunk = 16384;
    const long long nchunks = (long long)((nblk + chunk - 1) / chunk);
#pragma omp parallel for schedule(dynamic, 1) num_threads(threads) reduction(min : bad)
    for (long long c = 0; c < nchunks; c++) {
        const uint64_t b0 = (uint64_t)c * chunk, b1 = std::min(nblk, b0 + chunk);
        const uint64_t off = b0 * 64, len = (b1 - b0) * 64;
        uint64_t r;
        if (g_isa == 2) r = pack_avx512(a + off, len, out + off / 4);
        else if (g_isa == 1) r = pack_avx2(a + off, len, out + off / 4);
        else r = pack_scalar(a + off, len, out + off / 4);
        if (r != ~0ull) bad = std::min(bad, off + r);
    }
    if (whole < n_bases) {  // tail: pad to a whole number of 16-base words with 'A' (code 0)
        uint8_t tmp[64];
        const uint64_t rem = n_bases - whole, padded = (rem + 15) & ~15ull;
        memset(tmp, 'A', sizeof tmp);
        memcpy(tmp, a + whole, rem);
        const uint64_t r = pack_scalar(tmp, padded, out + whole / 4);
        if (r != ~0ull && r < rem) bad = std::min(bad, whole + r);
    }
    return bad;
}

// Single-threaded variant for callers that parallelise over strings themselves: packs one string into whole words
// starting at `packed` (bases past n_bases in the last word are zero).
uint64_t sg_host_pack_2bit_st(const char *ascii, uint64_t n_bases, uint32_t *packed)
{
    if (g_isa < 0) sg_host_pack_2bit("", 0, packed, 1);
    const uint8_t *a = (const uint8_t *)ascii;
    uint8_t *out = (uint8_t *)packed;
    const uint64_t whole = n_bases & ~63ull;
    uint64_t bad = ~0ull;
    if (whole) {
        if (g_isa == 2) bad = pack_avx512(a, whole, out);
        else if (g_isa == 1) bad = pack_avx2(a, whole, out);
        else bad = pack_scalar(a, whole, out);
    }
    if (whole < n_bases) {
        uint8_t tmp[64];
        const uint64_t rem = n_bases - whole, padded = (rem + 15) & ~15ull;
        memset(tmp, 'A', sizeof tmp);
        memcpy(tmp, a + whole, rem);
        const uint64_t r = pack_scalar(tmp, padded, out + whole / 4);
        if (r != ~0ull && r < rem) bad = std::min(bad, whole + r);
    }
    return bad;
}

int sg_host_pack_isa(void)
{
    if (g_isa < 0) {
        uint32_t w = 0;
        sg_host_pack_2bit("", 0, &w, 1);
    }
    return g_isa;
}

}  // extern "C"
