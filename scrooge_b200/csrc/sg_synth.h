// sg_synth.h -- deterministic synthetic pair generator, identical on host and device.
//
// Shapes follow SURVEY.md section 8(d): i.i.d. uniform text, read derived by walking the text with a
// per-base edit probability and a sub:ins:del ratio (PacBio-like 6:50:54 from the reference's
// DATASETS.md:51, Illumina-like 90:5:5), 64 slack bases appended to the text so that the last window is
// not text-starved.  A counter-based generator (SplitMix64 keyed by seed and pair index) replaces the
// survey's mt19937_64 suggestion so that a CUDA thread and a host thread produce the same pair.
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define SG_HD __host__ __device__ __forceinline__
#else
#define SG_HD inline
#endif

struct SgSynthParams {
    uint64_t seed;
    uint32_t read_len;
    uint32_t err_threshold;  // edit iff (draw >> 32) < err_threshold
    uint32_t w_sub, w_ins, w_del;
    uint32_t slack;
};

SG_HD uint64_t sg_splitmix64(uint64_t &state)
{
    uint64_t z = (state += 0x9E3779B97F4A7C15ull);
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

SG_HD uint64_t sg_synth_stride(uint32_t read_len, uint32_t slack)
{
    // deletions lengthen the text relative to the read; the generator stops deleting once the text is L/8 + 32 bases
    // ahead of the read (30 sigma above what the PacBio-like model does at 15 % error on 10 kbp), so a text never
    // exceeds L + L/8 + 32 bases + slack.  Rounded up to a multiple of 16 (whole packed words per row).
    return ((uint64_t)read_len + read_len / 8u + 32ull + slack + 15ull) & ~15ull;
}

// Generates pair `pair`; returns the text length.  text must hold sg_synth_stride() bytes, read read_len.
SG_HD uint64_t sg_synth_pair(const SgSynthParams &p, uint64_t pair, char *text, char *read)
{
    const char bases[4] = {'A', 'C', 'G', 'T'};
    uint64_t state = p.seed ^ (pair * 0xD1342543DE82EF95ull + 0x2545F4914F6CDD1Dull);
    sg_splitmix64(state);
    const uint32_t wsum = p.w_sub + p.w_ins + p.w_del;
    uint64_t tl = 0;
    uint32_t rl = 0;
    while (rl < p.read_len) {
        uint64_t r = sg_splitmix64(state);
        uint32_t tb = (uint32_t)(r & 3u);
        uint32_t ob = (uint32_t)((r >> 2) & 3u);
        bool edit = (uint32_t)(r >> 32) < p.err_threshold && wsum > 0;
        uint32_t pick = edit ? (uint32_t)((r >> 8) & 0xFFFFFFu) % wsum : 0u;
        // a deletion is only taken while tl < rl + L/8 + 32, so tl never exceeds L + L/8 + 32 (+ slack)
        if (edit && pick >= p.w_sub + p.w_ins && tl >= (uint64_t)rl + p.read_len / 8u + 32ull) edit = false;
        if (!edit) {
            text[tl++] = bases[tb];
            read[rl++] = bases[tb];
        } else {
            if (pick < p.w_sub) {  // substitution: one of the three other bases
                uint32_t alt = (tb + 1u + (uint32_t)((r >> 4) & 0xFu) % 3u) & 3u;
                text[tl++] = bases[tb];
                read[rl++] = bases[alt];
            } else if (pick < p.w_sub + p.w_ins) {  // insertion: read gains a base, text not consumed
                read[rl++] = bases[ob];
            } else {  // deletion: text base skipped
                text[tl++] = bases[tb];
            }
        }
    }
    for (uint32_t s = 0; s < p.slack; s++) {
        uint64_t r = sg_splitmix64(state);
        text[tl++] = bases[r & 3u];
    }
    return tl;
}

// i.i.d. uniform genome base number `i` of genome `seed` (stateless, so any thread can produce any stretch)
SG_HD char sg_synth_genome_base(uint64_t seed, uint64_t i)
{
    uint64_t st = seed ^ ((i >> 5) * 0x9FB21C651E98DF25ull);
    const uint64_t r = sg_splitmix64(st);  // 32 bases per draw
    const char bases[4] = {'A', 'C', 'G', 'T'};
    return bases[(r >> (2 * (i & 31))) & 3u];
}

// Read `idx` of a read-mapping workload (BASELINE.json configs[3]): sampled at a uniform position of the genome and
// mutated like sg_synth_pair.  Returns the true start position.  genome_len must exceed 2 * read_len + 64.
SG_HD uint64_t sg_synth_read_from_genome(const SgSynthParams &p, uint64_t idx, const char *genome, uint64_t genome_len, char *read)
{
    const char bases[4] = {'A', 'C', 'G', 'T'};
    uint64_t state = p.seed ^ (idx * 0xD1342543DE82EF95ull + 0x632BE59BD9B4E019ull);
    sg_splitmix64(state);
    const uint64_t span = genome_len - 2ull * p.read_len - 64ull;
    const uint64_t pos = sg_splitmix64(state) % span;
    const uint32_t wsum = p.w_sub + p.w_ins + p.w_del;
    uint64_t t = pos;
    uint32_t rl = 0;
    while (rl < p.read_len) {
        const uint64_t r = sg_splitmix64(state);
        const bool room = t + 1 < pos + 2ull * p.read_len;  // the walk never leaves [pos, pos + 2L)
        const char g = genome[t];
        uint32_t tb = g == 'A' ? 0u : (g == 'C' ? 1u : (g == 'G' ? 2u : 3u));
        bool edit = (uint32_t)(r >> 32) < p.err_threshold && wsum > 0;
        const uint32_t pick = edit ? (uint32_t)((r >> 8) & 0xFFFFFFu) % wsum : 0u;
        if (edit && pick >= p.w_sub + p.w_ins && !room) edit = false;
        if (!edit) {
            read[rl++] = bases[tb];
            if (room) t++;
        } else if (pick < p.w_sub) {
            read[rl++] = bases[(tb + 1u + (uint32_t)((r >> 4) & 0xFu) % 3u) & 3u];
            if (room) t++;
        } else if (pick < p.w_sub + p.w_ins) {
            read[rl++] = bases[(r >> 2) & 3u];
        } else {
            t++;
        }
    }
    return pos;
}
