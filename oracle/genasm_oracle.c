/*
 * genasm_oracle.c -- CPU restatement of Scrooge's windowed GenASM aligner.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is the parity checker for the CUDA path.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may build, load or call it.  The product library
 * (scrooge_b200/csrc) never links or calls anything in oracle/.
 *
 * Parity status: PINNED.  tests/test_oracle.py checks this restatement against
 *   (a) the reference's own known-answer vectors (src/tests.cu:236-246,
 *       src/tests.cu:275-284, src/library_example.cu:12-13),
 *   (b) golden outputs produced by the unmodified reference genasm_cpu.cpp
 *       compiled into oracle/_ref/ (the JSON files in tests/golden/, made by
 *       tests/golden/make_golden.py), and
 *   (c) when oracle/_ref/ is present, live differential runs on random pairs.
 *
 * Each function cites the reference lines (relative to /root/reference/) it
 * restates.  The restatement is written from the algorithm, in plain C with
 * run-time (W, O) instead of the reference's compile-time macros, and stores
 * full entries (SENE) without DENT; the reference's three storage toggles do
 * not change results.
 */
#define _POSIX_C_SOURCE 200809L
#include <stdint.h>
#include <stdlib.h>
#include <time.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define SGO_MAX_W 256
#define SGO_WORDS (SGO_MAX_W / 64)

/* one W-bit vector, W <= 256, as 64-bit words, least significant first; only the first ceil(W/64) words are
 * touched (the reference switches to arrays of 32-bit words above 64 bits, src/bitvector.hpp:42-47; the
 * arithmetic is the same) */
typedef struct { uint64_t w[SGO_WORDS]; } vec_t;

static inline vec_t v_fill(int nw, uint64_t x) { vec_t r; for (int k = 0; k < SGO_WORDS; k++) r.w[k] = k < nw ? x : 0; return r; }
static inline vec_t v_and(int nw, vec_t a, vec_t b) { for (int k = 0; k < nw; k++) a.w[k] &= b.w[k]; return a; }
static inline vec_t v_or(int nw, vec_t a, vec_t b) { for (int k = 0; k < nw; k++) a.w[k] |= b.w[k]; return a; }
static inline vec_t v_shl1(int nw, vec_t a)   /* towards higher bit indices, zero fill (src/bitvector.hpp:115-140) */
{
    for (int k = nw - 1; k > 0; k--) a.w[k] = (a.w[k] << 1) | (a.w[k - 1] >> 63);
    a.w[0] <<= 1;
    return a;
}
static inline int v_bit(vec_t a, int b) { return (int)((a.w[b >> 6] >> (b & 63)) & 1ull); }
/* the `bits` lowest bits set */
static inline vec_t v_low_mask(int nw, int bits)
{
    vec_t r = v_fill(nw, 0);
    for (int k = 0; k < nw; k++) {
        int t = bits - 64 * k;
        r.w[k] = t >= 64 ? ~0ull : (t <= 0 ? 0ull : ((1ull << t) - 1ull));
    }
    return r;
}

typedef struct {
    int W;        /* window size in characters; K == W (src/genasm_cpu.cpp:7-8, scripts/profile.py:29) */
    int O;        /* window overlap (src/genasm_cpu.cpp:9) */
    int tb_limit; /* W - O  (src/genasm_cpu.cpp:50) */
} sgo_cfg;

/* scratch for one worker thread: R[d][i] for d in [0,W], i in [0,W]
 * (src/genasm_cpu.cpp:71-78, SENE indexing COLUMNS*d + i) */
typedef struct {
    vec_t R[(SGO_MAX_W + 1) * (SGO_MAX_W + 1)];
} sgo_scratch;

typedef struct {
    uint64_t windows;    /* number of windows processed */
    uint64_t dc_entries; /* sum over windows of (d_w + 1) * (n_w + 1): the early-termination-minimal
                            number of R[d][i] entries (src/genasm_cpu.cpp:214-216,278-283) */
    uint64_t tb_steps;   /* number of traceback steps */
} sgo_stats;


/* ASCII -> base codes A0 C1 G2 T3, case-insensitive (src/genasm_cpu.cpp:462-493).
 * Returns -1 and the offending position through *bad_pos instead of assert(false). */
int sgo_ascii_to_codes(const char *ascii, size_t len, uint8_t *codes, size_t *bad_pos)
{
    for (size_t i = 0; i < len; i++) {
        switch (ascii[i]) {
            case 'A': case 'a': codes[i] = 0; break;
            case 'C': case 'c': codes[i] = 1; break;
            case 'G': case 'g': codes[i] = 2; break;
            case 'T': case 't': codes[i] = 3; break;
            default:
                if (bad_pos) *bad_pos = i;
                return -1;
        }
    }
    return 0;
}

/* 2-bit packing in the reference GPU layout: 4 bases per byte, base k of a byte in bits
 * 7-2k..6-2k, tail byte zero padded (src/genasm_gpu.cu:640-673; KAT src/tests.cu:583-606). */
int sgo_ascii_to_twobit_ref_layout(const char *ascii, size_t len, uint8_t *out)
{
    size_t nbytes = (len + 3) / 4;
    memset(out, 0, nbytes);
    for (size_t i = 0; i < len; i++) {
        uint8_t c;
        size_t bad;
        if (sgo_ascii_to_codes(ascii + i, 1, &c, &bad) != 0) return -1;
        out[i / 4] |= (uint8_t)(c << (6 - 2 * (i % 4)));
    }
    return 0;
}

/* Pattern bitmasks (src/genasm_cpu.cpp:178-198): bit b of masks[c] is 0 iff pattern[m-1-b] == c,
 * every other bit (including b >= m) is 1. */
static void pattern_masks(int nw, int m, const uint8_t *pattern, vec_t masks[4])
{
    masks[0] = masks[1] = masks[2] = masks[3] = v_fill(nw, ~0ull);
    for (int b = 0; b < m; b++) {
        masks[pattern[m - 1 - b]].w[b >> 6] &= ~(1ull << (b & 63));
    }
}

/* Distance calculation for one window (src/genasm_cpu.cpp:210-288), SENE + early termination.
 * Vectors are W-bit; they are held in ceil(W/64) 64-bit words and truncated to W bits after every shift so
 * that W = 32 behaves like the reference's 32-bit element type (src/bitvector.hpp:32-49,115-140); bits
 * at and above m are never examined and shifts only move bits upwards, so the truncation is neutral
 * for every other W as well.
 * Returns d_w, the smallest d for which bit m-1 of R[d][0] is zero. */
static int window_dc(const sgo_cfg *cfg, int n, const uint8_t *text, int m, const uint8_t *pattern,
                     vec_t *R)
{
    const int W = cfg->W;
    const int cols = W + 1;
    const int nw = (W + 63) / 64;
    const vec_t wmask = v_low_mask(nw, W);
    vec_t pm[4];
    pattern_masks(nw, m, pattern, pm);

    for (int d = 0; d <= W; d++) {
        for (int i = n; i >= 0; i--) {
            vec_t center;
            if (i == n) {
                /* boundary column: all ones shifted left by d (src/genasm_cpu.cpp:225-231,239-245) */
                vec_t low = v_low_mask(nw, d);
                center = wmask;
                for (int k = 0; k < nw; k++) center.w[k] &= ~low.w[k];
            } else {
                /* note: text[i] is only touched for i < n (quirk Q6) */
                vec_t right = R[cols * d + (i + 1)];
                vec_t mat = v_and(nw, v_or(nw, v_shl1(nw, right), pm[text[i]]), wmask);
                if (d == 0) {
                    center = mat; /* src/genasm_cpu.cpp:232-238 */
                } else {
                    vec_t top = R[cols * (d - 1) + i];
                    vec_t topright = R[cols * (d - 1) + (i + 1)];
                    vec_t sub = v_and(nw, v_shl1(nw, topright), wmask);
                    vec_t ins = v_and(nw, v_shl1(nw, top), wmask);
                    vec_t del = topright;
                    center = v_and(nw, v_and(nw, mat, sub), v_and(nw, ins, del)); /* src/genasm_cpu.cpp:246-252 */
                }
            }
            R[cols * d + i] = center;
            if (i == 0 && v_bit(center, m - 1) == 0) {
                return d; /* early termination, src/genasm_cpu.cpp:278-283 */
            }
        }
    }
    return -1; /* unreachable with K == W (quirk Q4) */
}

static char *emit_run(char *out, int count, char type)
{
    /* the reference prints "%d%c" per run (src/genasm_cpu.cpp:389,401) */
    if (count >= 100) *out++ = (char)('0' + count / 100);
    if (count >= 10) *out++ = (char)('0' + (count / 10) % 10);
    *out++ = (char)('0' + count % 10);
    *out++ = type;
    return out;
}

/* Traceback of one window (src/genasm_cpu.cpp:290-409), SENE bit tests, priority I > D > X > '='.
 * Runs are encoded per window and flushed at window end (quirk Q2). */
static int window_tb(const sgo_cfg *cfg, int n, int m, const vec_t *R, int d_w,
                     int *text_consumed, int *pattern_consumed, char **cigar, uint64_t *steps)
{
    const int cols = cfg->W + 1;
    int i = 0, j = 0, d = d_w;
    char cur_type = ' ';
    int cur_count = 0;

    while (j < m) {
        if (i >= cfg->tb_limit) break; /* src/genasm_cpu.cpp:309-310 */
        if (j >= cfg->tb_limit) break;

        int i_limit = i >= n;
        int d_limit = d == 0;
        int can_ins, can_del, can_sub;
        if (j < m - 1) {
            /* bit index of pattern position J is m-1-J (src/genasm_cpu.cpp:59,321-323) */
            can_ins = !d_limit && !v_bit(R[cols * (d - 1) + i], m - 1 - (j + 1));
            can_del = !d_limit && !i_limit && !v_bit(R[cols * (d - 1) + (i + 1)], m - 1 - j);
            can_sub = !d_limit && !i_limit && !v_bit(R[cols * (d - 1) + (i + 1)], m - 1 - (j + 1));
        } else {
            can_ins = !d_limit; /* src/genasm_cpu.cpp:336-343 */
            can_del = 0;
            can_sub = !d_limit && !i_limit;
        }

        char type;
        if (can_ins)      { j++; d--; type = 'I'; }
        else if (can_del) { i++; d--; type = 'D'; }
        else if (can_sub) { i++; j++; d--; type = 'X'; }
        else              { i++; j++; type = '='; }
        (*steps)++;

        if (type != cur_type) {
            if (cur_count > 0) *cigar = emit_run(*cigar, cur_count, cur_type);
            cur_type = type;
            cur_count = 1;
        } else {
            cur_count++;
        }
    }
    if (cur_count > 0) *cigar = emit_run(*cigar, cur_count, cur_type);

    *text_consumed = i;
    *pattern_consumed = j;
    return d_w - d; /* edits used, src/genasm_cpu.cpp:407 */
}

/* Window loop for one pair (src/genasm_cpu.cpp:411-438).  cigar must hold 4*read_len+1 bytes
 * (src/genasm_cpu.cpp:520).  Returns the edit distance; *ref_consumed = sum of text_consumed. */
int64_t sgo_align_codes(int W, int O, const uint8_t *ref, uint64_t ref_len, const uint8_t *read,
                        uint64_t read_len, char *cigar, uint64_t *ref_consumed, sgo_stats *stats,
                        sgo_scratch *scratch)
{
    sgo_cfg cfg = { W, O, W - O };
    uint64_t ref_idx = 0, read_idx = 0;
    int64_t edit_distance = 0;
    char *out = cigar;
    sgo_stats local = { 0, 0, 0 };

    while (read_idx < read_len) {
        int n = (int)((ref_len - ref_idx) < (uint64_t)W ? (ref_len - ref_idx) : (uint64_t)W);
        int m = (int)((read_len - read_idx) < (uint64_t)W ? (read_len - read_idx) : (uint64_t)W);
        int d_w = window_dc(&cfg, n, ref + ref_idx, m, read + read_idx, scratch->R);
        int tc = 0, pc = 0;
        int used = window_tb(&cfg, n, m, scratch->R, d_w, &tc, &pc, &out, &local.tb_steps);
        local.windows++;
        local.dc_entries += (uint64_t)(d_w + 1) * (uint64_t)(n + 1);
        edit_distance += used;
        ref_idx += (uint64_t)tc;
        read_idx += (uint64_t)pc;
    }
    *out = '\0';
    if (ref_consumed) *ref_consumed = ref_idx;
    if (stats) {
        stats->windows += local.windows;
        stats->dc_entries += local.dc_entries;
        stats->tb_steps += local.tb_steps;
    }
    return edit_distance;
}

/*
 * Batch entry points (flat blobs + offsets; offsets have n+1 entries).
 * Unstructured interface: src/genasm_cpu.cpp:557-609 (all N results are returned, quirk Q1).
 * cigar_blob receives the N NUL-terminated strings at cigar_off[p] = 4*query_off[p] + p.
 * Returns 0, or -(1+pair) if that pair holds a non-ACGT character (quirk Q9).
 */
int sgo_align_pairs(int W, int O, const char *text_blob, const uint64_t *text_off,
                    const char *query_blob, const uint64_t *query_off, uint64_t n_pairs, int threads,
                    int64_t *edit_out, uint64_t *ref_consumed_out, char *cigar_blob,
                    uint64_t *stats_out /* [3]: windows, dc_entries, tb_steps; may be NULL */,
                    int64_t *core_ns /* may be NULL */)
{
    if (W > SGO_MAX_W || W < 2 || O < 0 || O >= W) return -1000000;
    uint64_t tot_t = text_off[n_pairs], tot_q = query_off[n_pairs];
    uint8_t *tc = (uint8_t *)malloc(tot_t + 1), *qc = (uint8_t *)malloc(tot_q + 1);
    if (!tc || !qc) { free(tc); free(qc); return -1000001; }
    int err = 0;
    for (uint64_t p = 0; p < n_pairs && !err; p++) {
        size_t bad;
        if (sgo_ascii_to_codes(text_blob + text_off[p], text_off[p + 1] - text_off[p], tc + text_off[p], &bad) ||
            sgo_ascii_to_codes(query_blob + query_off[p], query_off[p + 1] - query_off[p], qc + query_off[p], &bad))
            err = -(int)(1 + p);
    }
    if (err) { free(tc); free(qc); return err; }
    if (threads < 1) threads = 1;
    uint64_t windows = 0, entries = 0, steps = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    /* OpenMP dynamic schedule over pairs, one scratch per thread (src/genasm_cpu.cpp:440-460) */
    #pragma omp parallel num_threads(threads) reduction(+:windows, entries, steps)
    {
        sgo_scratch *scratch = (sgo_scratch *)malloc(sizeof(sgo_scratch));
        sgo_stats st = { 0, 0, 0 };
        #pragma omp for schedule(dynamic)
        for (long long p = 0; p < (long long)n_pairs; p++) {
            uint64_t rc = 0;
            edit_out[p] = sgo_align_codes(W, O, tc + text_off[p], text_off[p + 1] - text_off[p],
                                          qc + query_off[p], query_off[p + 1] - query_off[p],
                                          cigar_blob + 4 * query_off[p] + (uint64_t)p, &rc, &st, scratch);
            if (ref_consumed_out) ref_consumed_out[p] = rc;
        }
        windows += st.windows; entries += st.dc_entries; steps += st.tb_steps;
        free(scratch);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (core_ns) *core_ns = (int64_t)(t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec);
    if (stats_out) { stats_out[0] = windows; stats_out[1] = entries; stats_out[2] = steps; }
    free(tc); free(qc);
    return 0;
}

/*
 * Read-mapping interface: src/genasm_cpu.cpp:495-555.  The text of candidate c is the genome suffix
 * starting at cand_start[c] (src/genasm_cpu.cpp:512-514); cand_read[c] indexes the read.
 * cigar_blob receives candidate c's string at cigar_off[c] (caller-provided, >= 4*len+1 apart).
 */
int sgo_align_candidates(int W, int O, const char *genome, uint64_t genome_len, const char *read_blob,
                         const uint64_t *read_off, uint64_t n_reads, const uint64_t *cand_start,
                         const uint32_t *cand_read, uint64_t n_cand, int threads, int64_t *edit_out,
                         uint64_t *ref_consumed_out, char *cigar_blob, const uint64_t *cigar_off,
                         uint64_t *stats_out, int64_t *core_ns)
{
    if (W > SGO_MAX_W || W < 2 || O < 0 || O >= W) return -1000000;
    uint8_t *gc = (uint8_t *)malloc(genome_len + 1), *rc8 = (uint8_t *)malloc(read_off[n_reads] + 1);
    if (!gc || !rc8) { free(gc); free(rc8); return -1000001; }
    size_t bad;
    if (sgo_ascii_to_codes(genome, genome_len, gc, &bad)) { free(gc); free(rc8); return -2000000; }
    for (uint64_t r = 0; r < n_reads; r++) {
        if (sgo_ascii_to_codes(read_blob + read_off[r], read_off[r + 1] - read_off[r], rc8 + read_off[r], &bad)) {
            free(gc); free(rc8); return -(int)(1 + r);
        }
    }
    for (uint64_t c = 0; c < n_cand; c++) {
        if (cand_start[c] > genome_len || cand_read[c] >= n_reads) { free(gc); free(rc8); return -3000000; }
    }
    if (threads < 1) threads = 1;
    uint64_t windows = 0, entries = 0, steps = 0;
    struct timespec t0, t1;
    clock_gettime(CLOCK_MONOTONIC, &t0);
    #pragma omp parallel num_threads(threads) reduction(+:windows, entries, steps)
    {
        sgo_scratch *scratch = (sgo_scratch *)malloc(sizeof(sgo_scratch));
        sgo_stats st = { 0, 0, 0 };
        #pragma omp for schedule(dynamic)
        for (long long c = 0; c < (long long)n_cand; c++) {
            uint64_t r = cand_read[c], consumed = 0;
            edit_out[c] = sgo_align_codes(W, O, gc + cand_start[c], genome_len - cand_start[c],
                                          rc8 + read_off[r], read_off[r + 1] - read_off[r],
                                          cigar_blob + cigar_off[c], &consumed, &st, scratch);
            if (ref_consumed_out) ref_consumed_out[c] = consumed;
        }
        windows += st.windows; entries += st.dc_entries; steps += st.tb_steps;
        free(scratch);
    }
    clock_gettime(CLOCK_MONOTONIC, &t1);
    if (core_ns) *core_ns = (int64_t)(t1.tv_sec - t0.tv_sec) * 1000000000ll + (t1.tv_nsec - t0.tv_nsec);
    if (stats_out) { stats_out[0] = windows; stats_out[1] = entries; stats_out[2] = steps; }
    free(gc); free(rc8);
    return 0;
}

/* CIGAR validator restating validateCigarString (src/tests.cu:27-169): format, coverage of the whole
 * read, staying inside the reference, '='/'X' agreeing with the bases, #edits == edit_distance.
 * Returns 0 when valid, otherwise a small positive reason code. */
int sgo_validate_cigar(const char *cigar, const char *ref, uint64_t ref_len, uint64_t ref_start,
                       const char *read, uint64_t read_len, int64_t edit_distance)
{
    uint64_t i = ref_start, j = 0;
    int64_t edits = 0;
    const char *p = cigar;
    while (*p) {
        if (*p < '0' || *p > '9') return 1; /* bad format (src/tests.cu:27-60) */
        uint64_t count = 0;
        while (*p >= '0' && *p <= '9') { count = count * 10 + (uint64_t)(*p - '0'); p++; }
        char type = *p++;
        if (count == 0) return 2;
        if (type == 'I') { j += count; edits += (int64_t)count; }
        else if (type == 'D') { i += count; edits += (int64_t)count; }
        else if (type == 'X' || type == '=' || type == 'M') {
            for (uint64_t e = 0; e < count; e++) {
                if (i >= ref_len || j >= read_len) return 3;
                char a = ref[i], b = read[j];
                if (a >= 'a') a = (char)(a - 32);
                if (b >= 'a') b = (char)(b - 32);
                if (type == 'X' && a == b) return 4;
                if (type == '=' && a != b) return 5;
                if (type == 'M' && a != b) edits++;
                i++; j++;
            }
            if (type == 'X') edits += (int64_t)count;
        } else return 6;
    }
    if (j != read_len) return 7; /* src/tests.cu:88-96 */
    if (i > ref_len) return 8;   /* src/tests.cu:98-101 */
    if (edits != edit_distance) return 9; /* src/tests.cu:163-166 */
    return 0;
}

int sgo_max_threads(void)
{
#ifdef _OPENMP
    return omp_get_num_procs();
#else
    return 1;
#endif
}
