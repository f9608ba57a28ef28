/*
 * ref_gpu_shim.cpp -- thin extern "C" door onto the UNMODIFIED reference GPU aligner.
 *
 * TEST / MEASUREMENT INFRASTRUCTURE ONLY (see oracle/genasm_oracle.c header).  Compiled by
 * `make -C oracle refgpu` together with /root/reference/src/genasm_gpu.cu and util.cpp, read from
 * where they lie, for sm_100a (the reference Makefile:6,12 builds -arch=$NVCC_ARCH -rdc=true);
 * nothing from the reference is copied into this repository.  Output:
 * oracle/_ref/libscrooge_refgpu_<tag>.so -- git-ignored, travels to the GPU box.
 *
 * Purpose: a same-box comparison point: the reference's own CUDA kernel
 * on the B200, through its own public interface genasm_gpu::align_all (src/genasm_gpu.hpp:5-8), timed by
 * its own core_algorithm_ns (src/genasm_gpu.cu:940-948).  Never linked into or called from the product.
 */
#include <chrono>
#include <cstdint>
#include <cstring>
#include <string>
#include <vector>
#include "genasm_gpu.hpp"

extern "C" {

int refgpu_config_w(void)
{
#ifdef CLI_KNOBS
    return CLI_W;
#else
    return 64; /* in-file default, src/genasm_gpu.cu:7 */
#endif
}

/* cigar_blob: pair p's NUL-terminated CIGAR at 4*query_off[p] + p (same convention as ref_shim.cpp).
 * core_ns = the reference's own kernel time; total_ns = wall time of the whole align_all call. */
int refgpu_align_pairs(const char *text_blob, const uint64_t *text_off, const char *query_blob,
                       const uint64_t *query_off, uint64_t n_pairs, int64_t *edit_out, char *cigar_blob,
                       int64_t *core_ns, int64_t *total_ns)
{
    genasm_gpu::enabled_algorithm_log = false;
    std::vector<std::string> texts, queries;
    texts.reserve(n_pairs);
    queries.reserve(n_pairs);
    for (uint64_t p = 0; p < n_pairs; p++) {
        texts.emplace_back(text_blob + text_off[p], text_off[p + 1] - text_off[p]);
        queries.emplace_back(query_blob + query_off[p], query_off[p + 1] - query_off[p]);
    }
    long long ns = 0;
    auto t0 = std::chrono::steady_clock::now();
    std::vector<Alignment_t> res = genasm_gpu::align_all(texts, queries, &ns);
    auto t1 = std::chrono::steady_clock::now();
    if (res.size() != n_pairs) return -1;
    for (uint64_t p = 0; p < n_pairs; p++) {
        edit_out[p] = res[p].edit_distance;
        if (cigar_blob) memcpy(cigar_blob + 4 * query_off[p] + p, res[p].cigar.c_str(), res[p].cigar.size() + 1);
    }
    if (core_ns) *core_ns = ns;
    if (total_ns) *total_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
    return 0;
}

} /* extern "C" */
