// genasm_gpu.cpp -- the reference's two C++ library overloads (src/genasm_gpu.hpp:7-8) on top of the C ABI.
//
// Host side of the drop-in: flattens the caller's containers into blobs + offsets, calls sg_align_pairs /
// sg_set_reference + sg_align_candidates, renders the packed runs to CIGAR text with all host threads
// (the reference renders serially through a stringstream per alignment, src/genasm_gpu.cu:881-888,
// 1049-1053 -- 20x its kernel time in its own README transcript, README.md:103-108).
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <iostream>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/genasm_gpu.hpp"
#include "../../include/scrooge_b200.h"

namespace genasm_gpu {

bool enabled_algorithm_log = true;

namespace {

struct Holder {
    sg_ctx *ctx = nullptr;
    int W = 0, O = 0, n_gpus = 0;
    // the genome resident on the GPUs: (address, length, content hash) of the Genome_t::content it was packed from
    const char *genome_ptr = nullptr;
    uint64_t genome_len = 0, genome_hash = 0;
    bool have_genome = false;
    ~Holder() { if (ctx) sg_ctx_destroy(ctx); }
};

// 64-bit content hash of a byte range, all host threads (a 3 Gbp genome at memory speed: ~30 ms, against ~100 ms per GPU
// for uploading and packing it again).  Every byte takes part: a caller that edits the genome in place gets a new upload.
uint64_t content_hash(const char *p, uint64_t n)
{
    const uint64_t blk = 1ull << 20, nblk = (n + blk - 1) / blk;
    uint64_t h = 0x9E3779B97F4A7C15ull ^ n;
#pragma omp parallel for schedule(static) reduction(^ : h)
    for (long long b = 0; b < (long long)nblk; b++) {
        const uint64_t a0 = (uint64_t)b * blk, a1 = std::min(n, a0 + blk);
        uint64_t x[4] = {0x243F6A8885A308D3ull + (uint64_t)b, 0x13198A2E03707344ull, 0xA4093822299F31D0ull, 0x082EFA98EC4E6C89ull};
        uint64_t i = a0;
        for (; i + 32 <= a1; i += 32) {
            uint64_t w[4];
            memcpy(w, p + i, 32);
            for (int k = 0; k < 4; k++) { x[k] = (x[k] ^ w[k]) * 0xFF51AFD7ED558CCDull; x[k] ^= x[k] >> 29; }
        }
        for (; i < a1; i++) { x[0] = (x[0] ^ (uint8_t)p[i]) * 0xFF51AFD7ED558CCDull; x[0] ^= x[0] >> 29; }
        uint64_t m = x[0] ^ (x[1] * 3) ^ (x[2] * 5) ^ (x[3] * 7);
        m = (m ^ (m >> 33)) * 0xC4CEB9FE1A85EC53ull;
        h ^= m + 0x9E3779B97F4A7C15ull * (uint64_t)(b + 1);
    }
    return h;
}

std::mutex g_mutex;  // calls are serialised per process (the reference is not re-entrant either)
Holder g_holder;

int env_int(const char *name, int dflt)
{
    const char *v = std::getenv(name);
    return v && *v ? std::atoi(v) : dflt;
}

sg_ctx *context()
{
    const int W = env_int("SG_WINDOW", 64);
    const int O = env_int("SG_OVERLAP", sg_default_overlap(W));
    const int n = env_int("SG_NUM_GPUS", 0);
    if (g_holder.ctx && (g_holder.W != W || g_holder.O != O || g_holder.n_gpus != n)) {
        sg_ctx_destroy(g_holder.ctx);
        g_holder.ctx = nullptr;
        g_holder.have_genome = false;
    }
    if (!g_holder.ctx) {
        if (sg_ctx_create_wo(&g_holder.ctx, nullptr, n, W, O) != SG_OK)
            throw std::runtime_error(std::string("scrooge_b200: ") + sg_last_error());
        g_holder.W = W;
        g_holder.O = O;
        g_holder.n_gpus = n;
    }
    return g_holder.ctx;
}

void check(int rc)
{
    if (rc != SG_OK) throw std::runtime_error(std::string("scrooge_b200: ") + sg_last_error());
}

std::vector<Alignment_t> collect(sg_result *res, Extra *extra, long long *core_algorithm_ns)
{
    const auto t_begin = std::chrono::steady_clock::now();
    const uint64_t n = sg_result_count(res);
    std::vector<Alignment_t> out(n);
    const int64_t *ed = sg_result_edit_distances(res);
    const uint64_t *rc = sg_result_ref_consumed(res);
    if (extra) extra->ref_consumed.resize(n);
    // Every CIGAR is rendered into a per-thread scratch that stays in cache (3 characters per run at most) and copied
    // into its string from there: the string's memory is written once (resize-then-render would zero-fill it first),
    // all host threads at once.
    const uint64_t *ro = sg_result_run_offsets(res);
#pragma omp parallel
    {
        std::vector<char> scratch(4096);
#pragma omp for schedule(dynamic, 64)
        for (long long i = 0; i < (long long)n; i++) {
            const uint64_t runs = ro ? ro[i + 1] - ro[i] : 0;
            if (runs) {
                if (scratch.size() < 4 * runs + 8) scratch.resize(4 * runs + 8);
                const int64_t len = sg_result_render_cigar(res, (uint64_t)i, scratch.data(), scratch.size());
                if (len > 0) out[i].cigar.assign(scratch.data(), (size_t)len);
            }
            out[i].edit_distance = ed[i];
            if (extra) extra->ref_consumed[i] = rc[i];
        }
    }
    const long long ns = sg_result_kernel_ns(res);
    if (core_algorithm_ns) *core_algorithm_ns = ns;
    if (extra) extra->total_ns = sg_result_total_ns(res);
    if (enabled_algorithm_log && ns > 0)  // same line the reference prints (src/genasm_gpu.cu:950-951)
        std::cerr << "core algorithm ran at " << (long long)((double)n * 1e9 / (double)ns) << " aligns/second" << std::endl;
    if (std::getenv("SG_DEBUG"))
        fprintf(stderr, "[sg] align_all: C ABI call %.1f ms, rendering %llu CIGAR strings %.1f ms\n", sg_result_total_ns(res) / 1e6,
                (unsigned long long)n, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_begin).count() * 1e3);
    sg_result_free(res);
    return out;
}

void views(const std::vector<std::string> &v, std::vector<const char *> &ptr, std::vector<uint64_t> &len)
{
    ptr.resize(v.size());
    len.resize(v.size());
    for (size_t i = 0; i < v.size(); i++) { ptr[i] = v[i].data(); len[i] = v[i].size(); }
}

std::vector<Alignment_t> pairs_impl(std::vector<std::string> &texts, std::vector<std::string> &queries, Extra *extra,
                                    long long *core_algorithm_ns)
{
    if (texts.size() != queries.size())  // the reference asserts (src/genasm_gpu.cu:984)
        throw std::runtime_error("scrooge_b200: texts and queries differ in size");
    std::lock_guard<std::mutex> lock(g_mutex);
    sg_ctx *ctx = context();
    if (enabled_algorithm_log) std::cerr << "Preparing data..." << std::endl;
    // the strings are handed over where they lie (pointer + length); nothing is flattened or copied on this side
    std::vector<const char *> tptr, qptr;
    std::vector<uint64_t> tlen, qlen;
    views(texts, tptr, tlen);
    views(queries, qptr, qlen);
    sg_result *res = nullptr;
    check(sg_align_pairs_v(ctx, tptr.data(), tlen.data(), qptr.data(), qlen.data(), texts.size(), 0, &res));
    return collect(res, extra, core_algorithm_ns);
}

std::vector<Alignment_t> mapping_impl(Genome_t &reference, std::vector<Read_t> &reads, Extra *extra, long long *core_algorithm_ns)
{
    std::lock_guard<std::mutex> lock(g_mutex);
    sg_ctx *ctx = context();
    if (enabled_algorithm_log) std::cerr << "Preparing data..." << std::endl;
    // the packed genome stays on the GPUs across calls: upload only when address, length or content changed
    // (the reference converts and uploads it on every call, src/genasm_gpu.cu:692-748,903)
    {
        const char *gp = reference.content.data();
        const uint64_t gl = reference.content.size(), gh = content_hash(gp, gl);
        if (!(g_holder.have_genome && g_holder.genome_ptr == gp && g_holder.genome_len == gl && g_holder.genome_hash == gh)) {
            g_holder.have_genome = false;
            check(sg_set_reference(ctx, gp, gl));
            g_holder.genome_ptr = gp; g_holder.genome_len = gl; g_holder.genome_hash = gh;
            g_holder.have_genome = true;
        }
    }
    std::vector<const char *> rptr(reads.size());
    std::vector<uint64_t> rlen(reads.size());
    uint64_t n_cand = 0;
    for (size_t r = 0; r < reads.size(); r++) { rptr[r] = reads[r].content.data(); rlen[r] = reads[r].content.size(); n_cand += reads[r].locations.size(); }
    std::vector<uint64_t> cstart;
    std::vector<uint32_t> cread;
    cstart.reserve(n_cand);
    cread.reserve(n_cand);
    for (size_t r = 0; r < reads.size(); r++) {
        for (const CandidateLocation_t &loc : reads[r].locations) {  // read-major, then location order (src/genasm_gpu.cu:961-967)
            if (loc.start_in_reference < 0) throw std::runtime_error("scrooge_b200: negative start_in_reference");
            cstart.push_back((uint64_t)loc.start_in_reference);
            cread.push_back((uint32_t)r);
        }
    }
    sg_result *res = nullptr;
    check(sg_align_candidates_v(ctx, rptr.data(), rlen.data(), reads.size(), cstart.data(), cread.data(), n_cand, 0, &res));
    return collect(res, extra, core_algorithm_ns);
}

}  // namespace

std::vector<Alignment_t> align_all(Genome_t &reference, std::vector<Read_t> &reads, long long *core_algorithm_ns)
{
    return mapping_impl(reference, reads, nullptr, core_algorithm_ns);
}

std::vector<Alignment_t> align_all(std::vector<std::string> &texts, std::vector<std::string> &queries, long long *core_algorithm_ns)
{
    return pairs_impl(texts, queries, nullptr, core_algorithm_ns);
}

std::vector<Alignment_t> align_all_ex(Genome_t &reference, std::vector<Read_t> &reads, Extra &extra, long long *core_algorithm_ns)
{
    return mapping_impl(reference, reads, &extra, core_algorithm_ns);
}

std::vector<Alignment_t> align_all_ex(std::vector<std::string> &texts, std::vector<std::string> &queries, Extra &extra,
                                      long long *core_algorithm_ns)
{
    return pairs_impl(texts, queries, &extra, core_algorithm_ns);
}

}  // namespace genasm_gpu
