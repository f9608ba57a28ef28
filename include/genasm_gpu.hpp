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
// (src/genasm_gpu.hpp:9, src/tests.cu:626,703); it is declared below for CUDA translation units and defined, as
// relocatable device code with the reference's byte layout, in scrooge_b200/lib/libscrooge_b200_rdc.a
// (scrooge_b200/csrc/sg_dropin_rdc.cu).  The aligner's own ingest is sg_dev_pack_2bit in scrooge_b200.h.
//
// Boundary types.  Inside the reference's tree this header is never seen: `#include "genasm_gpu.hpp"` in src/*.cu
// resolves to the reference's own file (same-directory rule) and libscrooge_b200.so defines exactly the symbols that
// file declares.  Elsewhere this header defers to the reference's util.hpp whenever one is
// reachable (it is then the single definition of Genome_t / Read_t / Alignment_t ..., whichever header a caller
// includes first) and brings its own layout-identical types (scrooge_types.hpp, scrooge_io.hpp) only when none is.
#pragma once

#if defined(SEED_FILE_MAF) && !defined(SCROOGE_B200_TYPES)
// the reference's util.hpp (src/util.hpp:8) has been included already: its types are the boundary types
#elif !defined(SCROOGE_B200_TYPES) && defined(__has_include)
#if __has_include("util.hpp")
#include "util.hpp"
#else
#include "scrooge_types.hpp"
#include "scrooge_io.hpp"
#endif
#else
#include "scrooge_types.hpp"
#include "scrooge_io.hpp"
#endif

namespace genasm_gpu {
    extern bool enabled_algorithm_log;

    // read-mapping interface: one Alignment_t per (read, location), read-major then location order
    std::vector<Alignment_t> align_all(Genome_t &reference, std::vector<Read_t> &reads, long long *core_algorithm_ns = NULL);
    // unstructured interface: queries[i] against a prefix of texts[i]
    std::vector<Alignment_t> align_all(std::vector<std::string> &texts, std::vector<std::string> &queries, long long *core_algorithm_ns = NULL);

#ifdef __CUDACC__
    // count strings: twobit_strings[i] receives ceil(string_lengths[i] / 4) bytes, base k of a byte in bits 7-2k:6-2k
    // (src/genasm_gpu.cu:640-681).  Needs -rdc=true and libscrooge_b200_rdc.a on the link line.
    __global__ void ascii_to_twobit_strings(int count, long long *string_lengths, char **ascii_strings, char **twobit_strings);
#endif

    // extensions
    struct Extra {
        std::vector<unsigned long long> ref_consumed;  // consumed text prefix per alignment = #(=,X,D)
        long long total_ns = 0;                        // whole call, wall clock
    };
    std::vector<Alignment_t> align_all_ex(Genome_t &reference, std::vector<Read_t> &reads, Extra &extra, long long *core_algorithm_ns = NULL);
    std::vector<Alignment_t> align_all_ex(std::vector<std::string> &texts, std::vector<std::string> &queries, Extra &extra, long long *core_algorithm_ns = NULL);
}
