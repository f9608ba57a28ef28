/*
 * ref_shim.cpp -- thin extern "C" door onto the UNMODIFIED reference CPU aligner.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/genasm_oracle.c header).  This translation unit is
 * compiled by oracle/Makefile together with /root/reference/src/genasm_cpu.cpp, read from where it
 * lies (nothing from the reference is copied into this repository); the result goes to
 * oracle/_ref/libscrooge_ref_w{64,32}.so, which is git-ignored but travels to the GPU box.
 *
 * It calls only the reference's public interface genasm_cpu::align_all (src/genasm_cpu.hpp:6-7).
 * Quirk Q1 (src/genasm_cpu.cpp:600-605): the unstructured overload copies out only even-indexed
 * pairs, so real pairs are interleaved with empty dummy pairs (zero windows, zero cost) and every
 * real result comes back.
 */
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include <omp.h>
#include "genasm_cpu.hpp"

extern "C" {

int ref_config_w(void)
{
#ifdef CLI_KNOBS
    return CLI_W;
#else
    return 64; /* in-file default, src/genasm_cpu.cpp:7 */
#endif
}

int ref_config_o(void)
{
#ifdef CLI_KNOBS
    return CLI_O;
#else
    return 33; /* src/genasm_cpu.cpp:9 */
#endif
}

int ref_max_threads(void) { return omp_get_num_procs(); }

/* cigar_blob: pair p's NUL-terminated CIGAR at 4*query_off[p] + p (same convention as the oracle). */
int ref_align_pairs(const char *text_blob, const uint64_t *text_off, const char *query_blob,
                    const uint64_t *query_off, uint64_t n_pairs, int threads, int64_t *edit_out,
                    char *cigar_blob, int64_t *core_ns)
{
    genasm_cpu::enabled_algorithm_log = false;
    std::vector<std::string> texts, queries;
    texts.reserve(2 * n_pairs);
    queries.reserve(2 * n_pairs);
    for (uint64_t p = 0; p < n_pairs; p++) {
        texts.emplace_back(text_blob + text_off[p], text_off[p + 1] - text_off[p]);
        queries.emplace_back(query_blob + query_off[p], query_off[p + 1] - query_off[p]);
        texts.emplace_back();   /* dummy pair at the odd index (Q1) */
        queries.emplace_back();
    }
    long long ns = 0;
    std::vector<Alignment_t> res = genasm_cpu::align_all(texts, queries, threads, &ns);
    if (res.size() != n_pairs) return -1;
    for (uint64_t p = 0; p < n_pairs; p++) {
        edit_out[p] = res[p].edit_distance;
        memcpy(cigar_blob + 4 * query_off[p] + p, res[p].cigar.c_str(), res[p].cigar.size() + 1);
    }
    if (core_ns) *core_ns = ns;
    return 0;
}

/* Read-mapping overload (src/genasm_cpu.hpp:6).  cand_* are grouped by read in read order:
 * read r owns candidates [cand_begin[r], cand_begin[r+1]).  Output c at cigar_off[c]. */
int ref_align_mapping(const char *genome, uint64_t genome_len, const char *read_blob,
                      const uint64_t *read_off, uint64_t n_reads, const uint64_t *cand_begin,
                      const uint64_t *cand_start, int threads, int64_t *edit_out, char *cigar_blob,
                      const uint64_t *cigar_off, int64_t *core_ns)
{
    genasm_cpu::enabled_algorithm_log = false;
    Genome_t g;
    g.content.assign(genome, genome_len);
    std::vector<Read_t> reads(n_reads);
    for (uint64_t r = 0; r < n_reads; r++) {
        reads[r].content.assign(read_blob + read_off[r], read_off[r + 1] - read_off[r]);
        for (uint64_t c = cand_begin[r]; c < cand_begin[r + 1]; c++) {
            CandidateLocation_t loc;
            loc.start_in_reference = (long long)cand_start[c];
            loc.start_in_chromosome = (long long)cand_start[c];
            loc.start_of_aligned_region = 0;
            loc.size_of_aligned_region = 0;
            loc.strand = true;
            reads[r].locations.push_back(loc);
        }
    }
    long long ns = 0;
    std::vector<Alignment_t> res = genasm_cpu::align_all(g, reads, threads, &ns);
    uint64_t n_cand = cand_begin[n_reads];
    if (res.size() != n_cand) return -1;
    for (uint64_t c = 0; c < n_cand; c++) {
        edit_out[c] = res[c].edit_distance;
        memcpy(cigar_blob + cigar_off[c], res[c].cigar.c_str(), res[c].cigar.size() + 1);
    }
    if (core_ns) *core_ns = ns;
    return 0;
}

} /* extern "C" */
