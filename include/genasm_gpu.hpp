// genasm_gpu.hpp -- drop-in C++ interface of the B200-native aligner.
//
// Same namespace, names, argument order (texts first, then queries) and result order as the reference's
// GPU library header (reference src/genasm_gpu.hpp:5-8), implemented on top of the C ABI in
// scrooge_b200.h.  Differences a caller can observe:
//   * all visible GPUs of the box are used (SG_NUM_GPUS=<n> limits it; the reference is pinned to GPU 0,
//     src/genasm_gpu.cu:67) -- alignments are independent, results are identical for any GPU count;
//   * errors throw std::runtime_error instead of exit()/assert() (src/cuda_util.hpp:3-10,
//     src/genasm_gpu.cu:636,984);
//   * the window configuration is a run-time choice: SG_WINDOW=<W> and SG_OVERLAP=<O> (default W=64, O=min(W/2+1, W-1):
//     64/33 and 32/17 as the reference ships them; any 2 <= W <= 256, 0 <= O < W, W-O <= 128 is accepted) where
//     the reference needs a recompile with -DCLI_W/-DCLI_K/-DCLI_O (src/genasm_gpu.cu:1-63);
//   * align_all_ex additionally returns the consumed reference prefix of every alignment.
// The reference also exports a __global__ ascii_to_twobit_strings used only by its own unit test
// (src/genasm_gpu.hpp:9, src/tests.cu:626); the equivalent here is sg_dev_pack_2bit in scrooge_b200.h.
#pragma once

#include "util.hpp"

namespace genasm_gpu {
    extern bool enabled_algorithm_log;

    // read-mapping interface: one Alignment_t per (read, location), read-major then location order
    std::vector<Alignment_t> align_all(Genome_t &reference, std::vector<Read_t> &reads, long long *core_algorithm_ns = NULL);
    // unstructured interface: queries[i] against a prefix of texts[i]
    std::vector<Alignment_t> align_all(std::vector<std::string> &texts, std::vector<std::string> &queries, long long *core_algorithm_ns = NULL);

    // extensions
    struct Extra {
        std::vector<unsigned long long> ref_consumed;  // consumed text prefix per alignment = #(=,X,D)
        long long total_ns = 0;                        // whole call, wall clock
    };
    std::vector<Alignment_t> align_all_ex(Genome_t &reference, std::vector<Read_t> &reads, Extra &extra, long long *core_algorithm_ns = NULL);
    std::vector<Alignment_t> align_all_ex(std::vector<std::string> &texts, std::vector<std::string> &queries, Extra &extra, long long *core_algorithm_ns = NULL);
}
