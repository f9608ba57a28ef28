// sg_host_api.cu -- layer 1 of include/scrooge_b200.h: host buffers in, host results out.
//
// Replaces the host side of the reference GPU library (src/genasm_gpu.cu:692-1065): cudaMallocManaged
// blobs, per-string descriptor loops, one synchronous kernel over everything and a linked-list walk per
// alignment become
//   * a host-side scatter over the context's GPUs: alignments are independent
//     (src/genasm_cpu.cpp:451-455), so each GPU gets a contiguous share balanced by query bases and there
//     is no inter-GPU exchange of any kind.  In mapping mode every GPU holds its own packed reference;
//   * per GPU, a three-slot software pipeline over sub-batches: while batch k is being aligned, batch
//     k+1's ASCII is on its way over PCIe (one contiguous copy, packed to 2 bit/base on the device) and
//     batch k-1's distances and compacted CIGAR runs are on their way back into pinned host memory;
//   * descriptors derived on the device from the offset arrays, a run slab with per-alignment capacity
//     2*|query|+8 (reference: 2*|query| entries, src/genasm_gpu.cu:995-1001) compacted on the device.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <chrono>
#include <algorithm>
#include <memory>
#include <omp.h>
#include <unistd.h>
#include <cuda_runtime.h>

#include "../../include/scrooge_b200.h"
#include "sg_internal.h"

namespace sg {

// ---- device / pinned buffers that only grow ----------------------------------------------------------
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return SG_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) {
            cudaGetLastError();
            want = bytes;
            e = cudaMalloc(&p, want);
            if (e != cudaSuccess) {
                cudaGetLastError();
                p = nullptr;
                return fail(SG_ERR_OOM, "cudaMalloc failed for " + std::to_string(bytes) + " bytes");
            }
        }
        cap = want;
        return SG_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

struct PinBuf {
    void *p = nullptr;
    size_t cap = 0;
    int reserve(size_t bytes)
    {
        if (bytes <= cap) return SG_OK;
        if (p) cudaFreeHost(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            cudaGetLastError();
            p = nullptr;
            return fail(SG_ERR_OOM, "cudaMallocHost failed for " + std::to_string(bytes) + " bytes");
        }
        cap = want;
        return SG_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// Pinned blocks for result pieces are expensive to create (page pinning) and cheap to reuse: a process-wide
// pool hands them to results and takes them back in sg_result_free.
struct PinnedPool {
    struct Block { uint8_t *p; size_t cap; };
    std::mutex mu;
    std::vector<Block> free_blocks;
    size_t cached = 0;
    // how much pinned memory the pool keeps for reuse: a quarter of the machine's RAM, at most 64 GB (a read-mapping
    // call over 8 M candidates returns 19 GB of runs; re-pinning that much costs seconds), SG_PINNED_CACHE_GB overrides
    const size_t kMaxCached = [] {
        if (const char *v = std::getenv("SG_PINNED_CACHE_GB")) return (size_t)std::max(0ll, std::atoll(v)) << 30;
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
        const size_t ram = pages > 0 && psz > 0 ? (size_t)pages * (size_t)psz : (size_t)64 << 30;
        return std::min<size_t>(ram / 4, (size_t)64 << 30);
    }();

    int acquire(size_t bytes, Block *out)
    {
        {
            std::lock_guard<std::mutex> g(mu);
            size_t best = free_blocks.size();
            for (size_t k = 0; k < free_blocks.size(); k++)
                if (free_blocks[k].cap >= bytes && (best == free_blocks.size() || free_blocks[k].cap < free_blocks[best].cap)) best = k;
            if (best != free_blocks.size()) {
                *out = free_blocks[best];
                cached -= out->cap;
                free_blocks.erase(free_blocks.begin() + best);
                return SG_OK;
            }
        }
        size_t want = std::max<size_t>(bytes + bytes / 16, 4096);
        void *p = nullptr;
        if (cudaMallocHost(&p, want) != cudaSuccess) {
            cudaGetLastError();
            trim(0);
            if (cudaMallocHost(&p, want) != cudaSuccess) {
                cudaGetLastError();
                return fail(SG_ERR_OOM, "cudaMallocHost failed for " + std::to_string(want) + " bytes");
            }
        }
        out->p = (uint8_t *)p;
        out->cap = want;
        return SG_OK;
    }
    void release(Block b)
    {
        if (!b.p) return;
        std::lock_guard<std::mutex> g(mu);
        if (cached + b.cap > kMaxCached) { cudaFreeHost(b.p); return; }
        free_blocks.push_back(b);
        cached += b.cap;
    }
    void trim(size_t keep)
    {
        std::lock_guard<std::mutex> g(mu);
        while (!free_blocks.empty() && cached > keep) {
            cached -= free_blocks.back().cap;
            cudaFreeHost(free_blocks.back().p);
            free_blocks.pop_back();
        }
    }
};
static PinnedPool g_pool;

// SG_DEBUG=1: where the host time of a call goes (printed by run_all)
struct HostTimes { double pack = 0, wait = 0, desc = 0, copy_out = 0; };
static HostTimes g_ht;
static const bool g_debug = std::getenv("SG_DEBUG") != nullptr;
struct ScopedT {
    double &acc; std::chrono::steady_clock::time_point t0;
    explicit ScopedT(double &a) : acc(a), t0(std::chrono::steady_clock::now()) {}
    ~ScopedT() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// Per-alignment host loops (descriptor arithmetic): serial up to a quarter of a million alignments -- a sub-batch of long reads --
// and split over plain threads (131 072 alignments each) above that (batches of short reads).  Deliberately not OpenMP: an OpenMP
// team that fits the cores spin-waits after its region and delays the CUDA calls that follow.
template <class F> void parallel_for(uint64_t n, int threads, F &&fn)
{
    const int nt = (int)std::min<uint64_t>((uint64_t)std::max(1, threads), n >> 17);
    if (nt <= 1) { fn((uint64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++) th.emplace_back([&fn, n, t, nt]() { fn(n * (uint64_t)t / nt, n * (uint64_t)(t + 1) / nt); });
    for (auto &x : th) x.join();
}

extern "C" uint64_t sg_host_pack_2bit_st(const char *ascii, uint64_t n_bases, uint32_t *packed);

constexpr int kMaxSlots = 8;
constexpr int kMaxDmaDepth = 8;

// One pipeline stage's worth of buffers: a sub-batch lives in a slot from upload to download.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_mid = nullptr, ev_end = nullptr;
    cudaEvent_t ev_dma[kMaxDmaDepth] = {};   // adaptive ingest: one per ASCII chunk copy in flight
    DevBuf ascii_t, ascii_q, packed_t, packed_q, desc, slab, counter, edit, refc, nruns, status, run_off, scan_tmp, runs, bad;
    PinBuf h_small, h_status, h_stage_t, h_stage_q, h_desc;
    PinnedPool::Block piece{nullptr, 0};
    // the batch in flight
    bool busy = false, mid_done = false;
    uint64_t a0 = 0, a1 = 0, total_runs = 0;
    uint64_t bad_bias[2] = {0, 0};   // hybrid ingest: the device-packed tail of a blob starts at this base
    int create()
    {
        SG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        SG_CUDA(cudaEventCreate(&ev_k0));
        SG_CUDA(cudaEventCreate(&ev_k1));
        SG_CUDA(cudaEventCreateWithFlags(&ev_mid, cudaEventDisableTiming));
        SG_CUDA(cudaEventCreateWithFlags(&ev_end, cudaEventDisableTiming));
        for (cudaEvent_t &e : ev_dma) SG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming | cudaEventBlockingSync));
        return SG_OK;
    }
    void destroy()
    {
        for (DevBuf *b : {&ascii_t, &ascii_q, &packed_t, &packed_q, &desc, &slab, &counter, &edit, &refc, &nruns, &status, &run_off,
                          &scan_tmp, &runs, &bad})
            b->release();
        for (PinBuf *b : {&h_small, &h_status, &h_stage_t, &h_stage_q, &h_desc}) b->release();
        g_pool.release(piece);
        piece = {nullptr, 0};
        if (ev_k0) cudaEventDestroy(ev_k0);
        if (ev_k1) cudaEventDestroy(ev_k1);
        if (ev_mid) cudaEventDestroy(ev_mid);
        if (ev_end) cudaEventDestroy(ev_end);
        for (cudaEvent_t e : ev_dma) if (e) cudaEventDestroy(e);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct Device {
    int id = 0;
    Slot slots[kMaxSlots];
    int n_slots = 3;
    DevBuf genome;  // packed reference, resident across calls
    uint64_t genome_len = 0;
    bool has_genome = false;
};

}  // namespace sg

using namespace sg;

struct sg_ctx {
    int W = 64;
    int O = 33;
    std::vector<Device> devs;
    // sub-batch rule: at least batch_bytes of ASCII AND at least min_batch_units alignments (one alignment
    // occupies one lane for its whole life -- 11 ms for a 10 kbp read with every lane busy -- so a launch needs an
    // alignment per lane to fill the device), but never more than max_batch_bytes (per-slot buffers)
    uint64_t batch_bytes = 256ull << 20;
    uint64_t max_batch_bytes = 3ull << 30;
    uint64_t min_batch_units = 65536;
    // ingest policy: pack to 2 bit/base on the host (all host threads, AVX-512) and upload a quarter of the bytes, or
    // upload ASCII and pack on the device.  Host packing wins when the upload is the bottleneck and there are enough
    // host threads per GPU (measured on the B200 box: 101 GB/s of ASCII on 16 threads vs 47 GB/s over PCIe).
    bool host_pack = false;
    int host_threads = 1;
    // hybrid ingest (blob inputs with host packing): this fraction of every blob's tail goes over PCIe as ASCII while
    // the host threads pack the head -- the copy engine and the packer run at the same time, and the device packs its
    // share in a few hundred microseconds.  Balance: f_ascii = 1 - 1 / (Rpcie / Rhost + 0.75) ~ 0.25 at 47 GB/s of PCIe
    // and the 79 GB/s the 16 host threads reach while the copy engine reads the same memory (101 GB/s alone).
    double ascii_frac = 0.25;
    // adaptive ingest (the default for blob inputs): the blob is cut into chunks; the host threads pack chunks from the
    // front (each packed chunk is uploaded as soon as it is done) while a feeder thread keeps `dma_depth` ASCII chunk
    // copies from the back in flight; the device packs whatever arrived as ASCII.  Whoever is faster takes more: no
    // tuned fraction, and it degrades gracefully when several ranks share the host's threads.
    bool adaptive = true;
    uint64_t chunk_bytes = 8ull << 20;
    int dma_depth = 4;
    uint64_t ascii_min_bytes = 8ull << 20;   // blobs smaller than this are not split
    std::mutex mu;  // calls on one context are serialised
};

struct sg_result {
    uint64_t n = 0;
    bool has_cigar = false;
    bool wide_runs = false;   // W - O > 63: run bytes with count 0 continue into the next byte (SG_RUN_COUNT)
    // distances, consumed prefixes and run offsets (n+1) live in ONE pinned block from the pool: the device-to-host
    // copies of every sub-batch land at their final place, nothing is copied or zero-filled on the host
    PinnedPool::Block store{nullptr, 0};
    int64_t *edit = nullptr;
    uint64_t *refc = nullptr;
    uint64_t *run_off = nullptr;  // n+1
    // packed runs, one pinned piece per processed sub-batch, in alignment order
    std::vector<PinnedPool::Block> pieces;
    std::vector<uint64_t> piece_first;   // first alignment of each piece
    std::vector<uint64_t> piece_run0;    // global run offset of each piece's first run
    std::vector<uint64_t> piece_runs;    // runs in each piece
    std::vector<uint8_t> flat;           // lazily flattened view for sg_result_runs
    int64_t kernel_ns = 0, total_ns = 0;
    ~sg_result() { for (auto &b : pieces) g_pool.release(b); g_pool.release(store); }
};

namespace {

struct ShardOut {
    int rc = SG_OK;
    std::string err;
    double kernel_ms = 0;
    std::vector<PinnedPool::Block> pieces;
    std::vector<uint64_t> piece_first, piece_runs;
    ~ShardOut() { for (auto &b : pieces) g_pool.release(b); }
};

// n strings, given either as one blob + n+1 offsets or as n pointers + n lengths (no flattening needed)
struct Strings {
    const char *blob = nullptr; const uint64_t *off = nullptr;
    const char *const *ptr = nullptr; const uint64_t *len = nullptr;
    const char *data(uint64_t i) const { return ptr ? ptr[i] : blob + off[i]; }
    uint64_t size(uint64_t i) const { return ptr ? len[i] : off[i + 1] - off[i]; }
};

// What differs between the two interfaces: where a sub-batch's texts and queries come from.
struct Workload {
    bool mapping = false;
    Strings text, query;   // pairs: text p / query p.  mapping: `query` holds the reads, the text is the resident genome
    const uint64_t *cand_start = nullptr; const uint32_t *cand_read = nullptr;
    uint32_t flags = 0;
};

#define R(x) do { int rc_ = (x); if (rc_) return rc_; } while (0)

int bad_base_error(const char *what, const char *unit, uint64_t index, uint64_t pos)
{
    // names the offender like the reference's assert would have stopped on it (src/genasm_gpu.cu:636)
    return fail(SG_ERR_BAD_BASE, std::string("non-ACGT character in ") + what + " of " + unit + " " + std::to_string(index) + " at position " +
                                     std::to_string(pos));
}

// Adaptive ingest of one blob range (see sg_ctx::adaptive).  All copies go to `st` in issue order.
int upload_blob_adaptive(sg_ctx *ctx, int dev_id, cudaStream_t st, cudaEvent_t *ev_dma, const char *src, uint64_t nbytes, DevBuf &d_ascii,
                         DevBuf &d_packed, PinBuf &h_stage, uint64_t *d_bad, uint64_t *bad_pos, uint64_t *split_out)
{
    const uint64_t words = sg_packed_words(nbytes);
    R(d_packed.reserve(words * 4));
    R(h_stage.reserve(words * 4 + 64));
    R(d_ascii.reserve(nbytes + 64));
    uint32_t *hp = h_stage.as<uint32_t>();
    const uint64_t C = ctx->chunk_bytes;   // a multiple of 256: chunks start on whole packed words and whole output cache lines
    const long long nch = (long long)((nbytes + C - 1) / C);
    const int depth = std::min(kMaxDmaDepth, std::max(1, ctx->dma_depth));
    const int packers = std::max(1, ctx->host_threads);
    std::mutex mu;
    long long front = 0, back = nch;   // host threads take chunk `front++`, the feeder chunk `--back`
    auto take_front = [&]() -> long long { std::lock_guard<std::mutex> g(mu); return front < back ? front++ : -1; };
    auto take_back = [&]() -> long long { std::lock_guard<std::mutex> g(mu); return back > front ? --back : -1; };
    uint64_t bad = ~0ull;
    int cuda_rc = 0;   // shared; only ever set to 1
    ScopedT t_pack(g_ht.pack);
#pragma omp parallel num_threads(packers + 1) reduction(min : bad)
    {
        // the team may be smaller than asked for: the feeder role exists only when there is a second thread
        const int tid = omp_get_thread_num(), team = omp_get_num_threads();
        bool ok = cudaSetDevice(dev_id) == cudaSuccess;
        if (ok && team > 1 && tid == 0) {
            int issued = 0;
            while (true) {
                if (issued >= depth && cudaEventSynchronize(ev_dma[issued % depth]) != cudaSuccess) { ok = false; break; }
                const long long c = take_back();
                if (c < 0) break;
                const uint64_t off = (uint64_t)c * C, len = std::min(C, nbytes - off);
                if (cudaMemcpyAsync(d_ascii.as<char>() + off, src + off, len, cudaMemcpyHostToDevice, st) != cudaSuccess ||
                    cudaEventRecord(ev_dma[issued % depth], st) != cudaSuccess) { ok = false; break; }
                issued++;
            }
        } else if (ok) {
            while (true) {
                const long long c = take_front();
                if (c < 0) break;
                const uint64_t off = (uint64_t)c * C, len = std::min(C, nbytes - off);
                const uint64_t r = sg_host_pack_2bit_st(src + off, len, hp + off / 16);
                if (r != ~0ull) { bad = std::min(bad, off + r); break; }
                if (cudaMemcpyAsync(d_packed.as<uint32_t>() + off / 16, hp + off / 16, ((len + 15) / 16) * 4, cudaMemcpyHostToDevice, st) !=
                    cudaSuccess) { ok = false; break; }
            }
        }
        if (!ok) {
#pragma omp atomic write
            cuda_rc = 1;
        }
    }
    if (cuda_rc) { cudaGetLastError(); return fail(SG_ERR_CUDA, "adaptive ingest: a CUDA call failed"); }
    *bad_pos = bad;
    if (bad != ~0ull) return SG_OK;   // the caller names the offender
    const uint64_t split = std::min(nbytes, (uint64_t)back * C);   // [0, split) packed by the host, [split, nbytes) arrived as ASCII
    *split_out = split;
    if (split < nbytes) return sg_dev_pack_2bit(d_ascii.as<char>() + split, nbytes - split, d_packed.as<uint32_t>() + split / 16, d_bad, st);
    const uint64_t used = (nbytes + 15) / 16;
    SG_CUDA(cudaMemsetAsync(d_packed.as<uint32_t>() + used, 0, (words - used) * 4, st));  // padding words the aligner may read
    return SG_OK;
}

// Strings [i0, i1) -> packed words on the device; start[k] receives the first base of string i0+k in the packed blob.
//   host_pack: packed by the host threads straight into pinned staging (a blob as one stream, separate strings each
//              at a word boundary), a quarter of the bytes cross PCIe;
//   else:      ASCII crosses PCIe (a blob straight from the caller's memory, separate strings gathered into pinned
//              staging first) and pack_2bit_kernel packs it; an offending base is then reported through d_bad.
int upload_strings(sg_ctx *ctx, int dev_id, cudaStream_t st, cudaEvent_t *ev_dma, const Strings &S, uint64_t i0, uint64_t i1, const char *what, const char *unit,
                   DevBuf &d_ascii, DevBuf &d_packed, PinBuf &h_stage, uint64_t *d_bad, uint64_t *start, uint64_t *bad_bias)
{
    *bad_bias = 0;
    const uint64_t n = i1 - i0;
    const int threads = ctx->host_threads;
    if (S.blob) {
        const uint64_t base = S.off[i0], nbytes = S.off[i1] - base;
        parallel_for(n, threads, [&](uint64_t k0, uint64_t k1) { for (uint64_t k = k0; k < k1; k++) start[k] = S.off[i0 + k] - base; });
        const uint64_t words = sg_packed_words(nbytes);
        if (ctx->adaptive && nbytes >= ctx->ascii_min_bytes && nbytes >= 256) {
            uint64_t bad = ~0ull, split = nbytes;
            R(upload_blob_adaptive(ctx, dev_id, st, ev_dma, S.blob + base, nbytes, d_ascii, d_packed, h_stage, d_bad, &bad, &split));
            if (bad != ~0ull) {
                const uint64_t i = (uint64_t)(std::upper_bound(S.off + i0, S.off + i1 + 1, base + bad) - S.off) - 1;
                return bad_base_error(what, unit, i, base + bad - S.off[i]);
            }
            *bad_bias = split < nbytes ? split : 0;
            return SG_OK;
        }
        R(d_packed.reserve(words * 4));
        if (!ctx->host_pack) {
            R(d_ascii.reserve(nbytes + 64));
            SG_CUDA(cudaMemcpyAsync(d_ascii.p, S.blob + base, nbytes, cudaMemcpyHostToDevice, st));
            return sg_dev_pack_2bit(d_ascii.as<char>(), nbytes, d_packed.as<uint32_t>(), d_bad, st);
        }
        R(h_stage.reserve(words * 4));
        uint32_t *hp = h_stage.as<uint32_t>();
        // hybrid: bases [split, nbytes) travel as ASCII (the copy is queued first, so the copy engine works while the
        // host packs [0, split)) and are packed on the device into the words that follow the host's
        uint64_t split = nbytes;
        if (ctx->ascii_frac > 0 && nbytes >= ctx->ascii_min_bytes && nbytes >= 128) split = (uint64_t)((double)nbytes * (1.0 - ctx->ascii_frac)) & ~63ull;
        if (split < nbytes) {
            R(d_ascii.reserve(nbytes - split + 64));
            SG_CUDA(cudaMemcpyAsync(d_ascii.p, S.blob + base + split, nbytes - split, cudaMemcpyHostToDevice, st));
        }
        const uint64_t used = (split + 15) / 16;
        uint64_t bad;
        { ScopedT t(g_ht.pack); bad = sg_host_pack_2bit(S.blob + base, split, hp, threads); }
        if (bad != ~0ull) {
            const uint64_t i = (uint64_t)(std::upper_bound(S.off + i0, S.off + i1 + 1, base + bad) - S.off) - 1;
            return bad_base_error(what, unit, i, base + bad - S.off[i]);
        }
        if (split < nbytes) {
            SG_CUDA(cudaMemcpyAsync(d_packed.p, hp, used * 4, cudaMemcpyHostToDevice, st));
            *bad_bias = split;
            return sg_dev_pack_2bit(d_ascii.as<char>(), nbytes - split, d_packed.as<uint32_t>() + used, d_bad, st);
        }
        memset(hp + used, 0, (words - used) * 4);  // padding words the aligner may read
        SG_CUDA(cudaMemcpyAsync(d_packed.p, hp, words * 4, cudaMemcpyHostToDevice, st));
        return SG_OK;
    }
    if (ctx->host_pack) {
        uint64_t w = 0;
        for (uint64_t k = 0; k < n; k++) { start[k] = w * 16; w += (S.len[i0 + k] + 15) / 16; }
        const uint64_t words = w + 8;
        R(d_packed.reserve(words * 4));
        R(h_stage.reserve(words * 4));
        uint32_t *hp = h_stage.as<uint32_t>();
        uint64_t bad_k = ~0ull;
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads) reduction(min : bad_k)
        for (long long k = 0; k < (long long)n; k++)
            if (sg_host_pack_2bit_st(S.ptr[i0 + k], S.len[i0 + k], hp + start[k] / 16) != ~0ull) bad_k = std::min<uint64_t>(bad_k, (uint64_t)k);
        if (bad_k != ~0ull) {
            std::vector<uint32_t> tmp((S.len[i0 + bad_k] + 15) / 16 + 1);
            return bad_base_error(what, unit, i0 + bad_k, sg_host_pack_2bit_st(S.ptr[i0 + bad_k], S.len[i0 + bad_k], tmp.data()));
        }
        memset(hp + w, 0, 8 * 4);
        SG_CUDA(cudaMemcpyAsync(d_packed.p, hp, words * 4, cudaMemcpyHostToDevice, st));
        return SG_OK;
    }
    uint64_t bytes = 0;
    for (uint64_t k = 0; k < n; k++) { start[k] = bytes; bytes += S.len[i0 + k]; }
    R(h_stage.reserve(bytes + 64));
    R(d_ascii.reserve(bytes + 64));
    R(d_packed.reserve(sg_packed_words(bytes) * 4));
    char *ha = h_stage.as<char>();
#pragma omp parallel for schedule(dynamic, 16) num_threads(threads)
    for (long long k = 0; k < (long long)n; k++)
        if (S.len[i0 + k]) memcpy(ha + start[k], S.ptr[i0 + k], S.len[i0 + k]);
    SG_CUDA(cudaMemcpyAsync(d_ascii.p, ha, bytes, cudaMemcpyHostToDevice, st));
    return sg_dev_pack_2bit(d_ascii.as<char>(), bytes, d_packed.as<uint32_t>(), d_bad, st);
}

// stage A: uploads, ingest, descriptors, alignment kernel, run-count scan; ends with ev_mid
int stage_a(sg_ctx *ctx, Device &d, Slot &s, const Workload &w, uint64_t a0, uint64_t a1, sg_result *res)
{
    const uint64_t n = a1 - a0;
    const bool want_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
    cudaStream_t st = s.stream;
    s.a0 = a0; s.a1 = a1; s.busy = true; s.mid_done = false; s.total_runs = 0;
    R(s.desc.reserve((5 * n + 1) * 8)); R(s.h_desc.reserve((5 * n + 1) * 8));
    R(s.counter.reserve(8)); R(s.edit.reserve(n * 8)); R(s.refc.reserve(n * 8));
    R(s.nruns.reserve(n * 4)); R(s.status.reserve(n)); R(s.run_off.reserve((n + 1) * 8));
    R(s.scan_tmp.reserve(sg_scan_tmp_bytes(n))); R(s.bad.reserve(16)); R(s.h_small.reserve(64));
    R(s.h_status.reserve(n));
    SG_CUDA(cudaMemsetAsync(s.bad.p, 0xFF, 16, st));
    // descriptors are built on the host: [tstart | tlen | qstart | qlen | slab_off (n+1)]
    uint64_t *h_tstart = s.h_desc.as<uint64_t>(), *h_tlen = h_tstart + n, *h_qstart = h_tlen + n, *h_qlen = h_qstart + n, *h_slab = h_qlen + n;
    uint64_t *d_tstart = s.desc.as<uint64_t>(), *d_tlen = d_tstart + n, *d_qstart = d_tlen + n, *d_qlen = d_qstart + n, *d_slab = d_qlen + n;
    const uint32_t *d_text;
    if (!w.mapping) {
        R(upload_strings(ctx, d.id, st, s.ev_dma, w.text, a0, a1, "text", "pair", s.ascii_t, s.packed_t, s.h_stage_t, s.bad.as<uint64_t>(), h_tstart, &s.bad_bias[0]));
        R(upload_strings(ctx, d.id, st, s.ev_dma, w.query, a0, a1, "query", "pair", s.ascii_q, s.packed_q, s.h_stage_q, s.bad.as<uint64_t>() + 1, h_qstart, &s.bad_bias[1]));
        parallel_for(n, ctx->host_threads, [&](uint64_t k0, uint64_t k1) {
            for (uint64_t k = k0; k < k1; k++) { h_tlen[k] = w.text.size(a0 + k); h_qlen[k] = w.query.size(a0 + k); }
        });
        d_text = s.packed_t.as<uint32_t>();
    } else {
        // reads referenced by this sub-batch: the contiguous index range [r0, r1] (candidates arrive read-major, so the
        // range is tight; each read is uploaded and packed once and shared by its candidates, cf. reference
        // twobit_reads, src/genasm_gpu.cu:784-796)
        uint32_t r0 = w.cand_read[a0], r1 = w.cand_read[a0];
        for (uint64_t c = a0; c < a1; c++) { r0 = std::min(r0, w.cand_read[c]); r1 = std::max(r1, w.cand_read[c]); }
        std::vector<uint64_t> rstart((uint64_t)r1 - r0 + 1);
        R(upload_strings(ctx, d.id, st, s.ev_dma, w.query, r0, (uint64_t)r1 + 1, "content", "read", s.ascii_q, s.packed_q, s.h_stage_q,
                         s.bad.as<uint64_t>() + 1, rstart.data(), &s.bad_bias[1]));
        for (uint64_t k = 0; k < n; k++) {
            const uint64_t cs = w.cand_start[a0 + k];
            const uint32_t r = w.cand_read[a0 + k];
            h_tstart[k] = cs;
            h_tlen[k] = d.genome_len - cs;  // the text runs to the end of the genome (src/genasm_cpu.cpp:512-514)
            h_qstart[k] = rstart[r - r0];
            h_qlen[k] = w.query.size(r);
        }
        d_text = d.genome.as<uint32_t>();
    }
    ScopedT t_desc(g_ht.desc);
    uint64_t slab_bytes = 0;
    if (!w.mapping && w.query.blob) {   // capacity 2*|query|+8 per alignment: the prefix sum is a difference of offsets
        const uint64_t *qo = w.query.off + a0;
        parallel_for(n + 1, ctx->host_threads, [&](uint64_t k0, uint64_t k1) { for (uint64_t k = k0; k < k1; k++) h_slab[k] = 2ull * (qo[k] - qo[0]) + 8ull * k; });
        slab_bytes = h_slab[n];
    } else {
        for (uint64_t k = 0; k < n; k++) { h_slab[k] = slab_bytes; slab_bytes += 2ull * h_qlen[k] + 8ull; }
        h_slab[n] = slab_bytes;
    }
    SG_CUDA(cudaMemcpyAsync(s.desc.p, s.h_desc.p, (5 * n + 1) * 8, cudaMemcpyHostToDevice, st));
    if (want_cigar) R(s.slab.reserve(slab_bytes + 16));
    SG_CUDA(cudaEventRecord(s.ev_k0, st));
    R(sg_dev_align_wo(ctx->W, ctx->O, d_text, d_tstart, d_tlen, s.packed_q.as<uint32_t>(), d_qstart, d_qlen, n, w.flags, s.slab.as<uint8_t>(), d_slab,
                   s.counter.as<uint64_t>(), s.edit.as<int64_t>(), s.refc.as<uint64_t>(), s.nruns.as<uint32_t>(), s.status.as<uint8_t>(),
                   nullptr, nullptr, st));
    SG_CUDA(cudaEventRecord(s.ev_k1, st));
    uint64_t *h = s.h_small.as<uint64_t>();
    SG_CUDA(cudaMemcpyAsync(h, s.bad.p, 16, cudaMemcpyDeviceToHost, st));
    if (want_cigar) {
        R(sg_dev_scan_runs(s.nruns.as<uint32_t>(), n, s.run_off.as<uint64_t>(), s.scan_tmp.p, st));
        SG_CUDA(cudaMemcpyAsync(h + 2, s.run_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    }
    SG_CUDA(cudaMemcpyAsync(res->edit + a0, s.edit.p, n * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(res->refc + a0, s.refc.p, n * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(s.h_status.p, s.status.p, n, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaEventRecord(s.ev_mid, st));
    return SG_OK;
}

// stage B: once the run total is known, gather the runs and send everything home; ends with ev_end
int stage_b(Slot &s, const Workload &w, sg_result *res)
{
    if (!s.busy || s.mid_done) return SG_OK;
    const uint64_t n = s.a1 - s.a0;
    const bool want_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
    cudaStream_t st = s.stream;
    { ScopedT t(g_ht.wait); SG_CUDA(cudaEventSynchronize(s.ev_mid)); }
    s.mid_done = true;
    const uint64_t *h = s.h_small.as<uint64_t>();
    if (h[0] != ~0ull || h[1] != ~0ull) {
        // device ingest found an offending base at position h[] of the uploaded ASCII: find its string via the starts
        const bool in_text = h[0] != ~0ull;
        const uint64_t pos = in_text ? h[0] + s.bad_bias[0] : h[1] + s.bad_bias[1];
        const uint64_t *hd = s.h_desc.as<uint64_t>();
        if (!w.mapping) {
            const uint64_t *start = in_text ? hd : hd + 2 * n;
            const uint64_t k = (uint64_t)(std::upper_bound(start, start + n, pos) - start) - 1;
            return bad_base_error(in_text ? "text" : "query", "pair", s.a0 + k, pos - start[k]);
        }
        uint64_t best = 0, best_start = 0;  // candidates of one read share a start: pick the read with the largest start <= pos
        for (uint64_t k = 0; k < n; k++)
            if (hd[2 * n + k] <= pos && hd[2 * n + k] >= best_start) { best_start = hd[2 * n + k]; best = w.cand_read[s.a0 + k]; }
        return bad_base_error("content", "read", best, pos - best_start);
    }
    if (want_cigar) {
        s.total_runs = h[2];
        R(s.runs.reserve(s.total_runs + 16));
        g_pool.release(s.piece);
        s.piece = {nullptr, 0};
        R(g_pool.acquire(s.total_runs + 16, &s.piece));
        // the mean number of runs per alignment of this sub-batch is known by now: short alignments are gathered by four lanes
        R(sg_dev_gather_runs_sized(s.slab.as<uint8_t>(), s.desc.as<uint64_t>() + 4 * n, s.nruns.as<uint32_t>(), s.run_off.as<uint64_t>(), n,
                                   s.runs.as<uint8_t>(), std::max<uint64_t>(1, (s.total_runs + n - 1) / n), st));
        SG_CUDA(cudaMemcpyAsync(s.piece.p, s.runs.p, s.total_runs, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(res->run_off + s.a0, s.run_off.p, n * 8, cudaMemcpyDeviceToHost, st));  // sub-batch-local, rebased in finalize()
    }
    SG_CUDA(cudaEventRecord(s.ev_end, st));
    return SG_OK;
}

// stage C: results of the slot's sub-batch into the caller-visible result; frees the slot
int stage_c(Slot &s, const Workload &w, sg_result *res, ShardOut &so)
{
    if (!s.busy) return SG_OK;
    const uint64_t n = s.a1 - s.a0;
    const bool want_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
    R(stage_b(s, w, res));
    { ScopedT t(g_ht.wait); SG_CUDA(cudaEventSynchronize(s.ev_end)); }
    s.busy = false;
    ScopedT t_out(g_ht.copy_out);
    const uint8_t *status = s.h_status.as<uint8_t>();
    if (const void *hit = n ? memchr(status, SG_ERR_CIGAR_OVERFLOW, n) : nullptr)
        return fail(SG_ERR_CIGAR_OVERFLOW, "alignment " + std::to_string(s.a0 + (uint64_t)((const uint8_t *)hit - status)) + " exceeded its run capacity");
    float ms = 0;
    SG_CUDA(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    so.kernel_ms += ms;
    if (want_cigar) {
        so.pieces.push_back(s.piece);
        so.piece_first.push_back(s.a0);
        so.piece_runs.push_back(s.total_runs);
        s.piece = {nullptr, 0};
    }
    return SG_OK;
}

// The pipeline over one device's share [c0, c1): weight prefix `woff` decides the sub-batch cuts.
void run_shard(sg_ctx *ctx, Device &d, const Workload &w, const uint64_t *woff, uint64_t per_unit_extra, uint64_t c0, uint64_t c1,
               sg_result *res, ShardOut &so)
{
    auto bail = [&](int rc) {
        so.rc = rc;
        so.err = g_last_error;
        for (Slot &s : d.slots) {  // let the device drain before the buffers are reused
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.busy = false;
        }
        cudaGetLastError();
    };
    if (cudaSetDevice(d.id) != cudaSuccess) { cudaGetLastError(); fail(SG_ERR_CUDA, "cudaSetDevice failed"); bail(SG_ERR_CUDA); return; }
    std::vector<uint64_t> cuts{c0};
    while (cuts.back() < c1) {
        // the sub-batch [a, b): b grows while the batch stays under max_batch_bytes and is either still small in bytes or
        // still short of min_batch_units alignments.  Both stop conditions are monotone in b: binary search for the
        // first b in (a, c1) that stops (a batch always takes at least one alignment).
        const uint64_t a = cuts.back();
        auto stops = [&](uint64_t b) {
            const uint64_t bytes = (woff[b + 1] - woff[a]) + per_unit_extra * (b + 1 - a);
            return bytes > ctx->max_batch_bytes || (bytes > ctx->batch_bytes && b - a >= ctx->min_batch_units);
        };
        uint64_t lo = a + 1, hi = c1;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (stops(mid)) hi = mid; else lo = mid + 1;
        }
        cuts.push_back(lo);
    }
    const int nb = (int)cuts.size() - 1;
    const int kSlots = d.n_slots;
    for (int k = 0; k < nb; k++) {
        Slot &s = d.slots[k % kSlots];
        int rc = stage_c(s, w, res, so);                       // frees the slot used by batch k - kSlots
        if (!rc) rc = stage_a(ctx, d, s, w, cuts[k], cuts[k + 1], res);
        if (!rc && k >= 1) rc = stage_b(d.slots[(k - 1) % kSlots], w, res);
        if (rc) { bail(rc); return; }
    }
    for (int k = std::max(0, nb - kSlots); k < nb; k++) {
        int rc = stage_c(d.slots[k % kSlots], w, res, so);
        if (rc) { bail(rc); return; }
    }
}

// splits [0,n) into contiguous parts with about equal weight, weight prefix given by off (n+1 entries)
std::vector<uint64_t> split_by_weight(const uint64_t *off, uint64_t n, int parts)
{
    std::vector<uint64_t> cut(parts + 1, n);
    cut[0] = 0;
    const uint64_t total = off[n] - off[0] + n;  // +1 per alignment so that empty queries still spread
    for (int k = 1; k < parts; k++) {
        const uint64_t target = total / parts * k;
        uint64_t lo = cut[k - 1], hi = n;
        while (lo < hi) {
            uint64_t mid = (lo + hi) / 2;
            if (off[mid] - off[0] + mid < target) lo = mid + 1; else hi = mid;
        }
        cut[k] = lo;
    }
    return cut;
}

void finalize(sg_result *res, std::vector<ShardOut> &shards)
{
    double kms = 0;
    for (ShardOut &so : shards) kms = std::max(kms, so.kernel_ms);
    res->kernel_ns = (int64_t)(kms * 1e6);
    if (!res->has_cigar) return;
    // pieces are in alignment order once the shards are concatenated; rebase per-piece offsets to global ones
    uint64_t run0 = 0;
    for (ShardOut &so : shards) {
        for (size_t k = 0; k < so.pieces.size(); k++) {
            res->piece_first.push_back(so.piece_first[k]);
            res->piece_run0.push_back(run0);
            res->piece_runs.push_back(so.piece_runs[k]);
            run0 += so.piece_runs[k];
            res->pieces.push_back(so.pieces[k]);
        }
        so.pieces.clear();
    }
    for (size_t k = 0; k < res->pieces.size(); k++) {
        const uint64_t a0 = res->piece_first[k];
        const uint64_t a1 = k + 1 < res->pieces.size() ? res->piece_first[k + 1] : res->n;
        const uint64_t base = res->piece_run0[k];
        if (base) {
            uint64_t *ro = res->run_off + a0;
            parallel_for(a1 - a0, (int)std::thread::hardware_concurrency(), [&](uint64_t k0, uint64_t k1) { for (uint64_t k = k0; k < k1; k++) ro[k] += base; });
        }
    }
    res->run_off[res->n] = run0;
}

const uint8_t *runs_of(const sg_result *r, uint64_t idx, uint64_t *count)
{
    *count = r->run_off[idx + 1] - r->run_off[idx];
    if (*count == 0) return nullptr;
    size_t k = (size_t)(std::upper_bound(r->piece_first.begin(), r->piece_first.end(), idx) - r->piece_first.begin()) - 1;
    return r->pieces[k].p + (r->run_off[idx] - r->piece_run0[k]);
}

int run_all(sg_ctx *ctx, const Workload &w, const uint64_t *woff, uint64_t per_unit_extra, uint64_t n, sg_result **out)
{
    auto t_begin = std::chrono::steady_clock::now();
    std::unique_ptr<sg_result> res(new sg_result);
    res->n = n;
    res->has_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
    res->wide_runs = ctx->W - ctx->O > 63;
    if (int rc = g_pool.acquire((3 * n + 2) * 8, &res->store)) return rc;
    res->edit = reinterpret_cast<int64_t *>(res->store.p);
    res->refc = reinterpret_cast<uint64_t *>(res->store.p) + n;
    res->run_off = reinterpret_cast<uint64_t *>(res->store.p) + 2 * n;
    res->run_off[0] = 0;
    res->run_off[n] = 0;
    const int nd = (int)ctx->devs.size();
    std::vector<ShardOut> shards(nd);
    if (n) {
        const std::vector<uint64_t> cut = split_by_weight(woff, n, nd);
        if (nd == 1) {
            run_shard(ctx, ctx->devs[0], w, woff, per_unit_extra, cut[0], cut[1], res.get(), shards[0]);
        } else {
            std::vector<std::thread> th;
            for (int k = 0; k < nd; k++)
                th.emplace_back([&, k]() { run_shard(ctx, ctx->devs[k], w, woff, per_unit_extra, cut[k], cut[k + 1], res.get(), shards[k]); });
            for (auto &t : th) t.join();
        }
        for (ShardOut &so : shards)
            if (so.rc) return fail(so.rc, so.err);
    }
    finalize(res.get(), shards);
    res->total_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
    if (g_debug) {
        fprintf(stderr, "[sg] call %.1f ms: host pack %.1f, waits %.1f, results out %.1f, descriptors %.1f\n", res->total_ns / 1e6,
                g_ht.pack * 1e3, g_ht.wait * 1e3, g_ht.copy_out * 1e3, g_ht.desc * 1e3);
        g_ht = HostTimes();
    }
    *out = res.release();
    return SG_OK;
}

}  // namespace

extern "C" {

int sg_ctx_create(sg_ctx **out, const int *device_ids, int n_devices, int W)
{
    if (W != 64 && W != 32) return fail(SG_ERR_BAD_ARG, "W must be 64 (O=33) or 32 (O=17); sg_ctx_create_wo takes any window");
    return sg_ctx_create_wo(out, device_ids, n_devices, W, sg_default_overlap(W));
}

int sg_ctx_window(const sg_ctx *ctx) { return ctx ? ctx->W : 0; }
int sg_ctx_overlap(const sg_ctx *ctx) { return ctx ? ctx->O : 0; }

int sg_ctx_create_wo(sg_ctx **out, const int *device_ids, int n_devices, int W, int O)
{
    if (!out) return fail(SG_ERR_BAD_ARG, "sg_ctx_create: null out");
    if (W < 2 || W > 256 || O < 0 || O >= W || W - O > 128)
        return fail(SG_ERR_BAD_ARG, "window configuration out of range: need 2 <= W <= 256, 0 <= O < W, W - O <= 128");
    const int avail = sg_device_count();
    if (avail == 0) return fail(SG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (n_devices <= 0) n_devices = avail;
    std::unique_ptr<sg_ctx, void (*)(sg_ctx *)> ctx(new sg_ctx, sg_ctx_destroy);
    ctx->W = W;
    ctx->O = O;
    ctx->devs.resize(n_devices);
    ctx->min_batch_units = 0;
    if (const char *v = std::getenv("SG_BATCH_MB")) {  // tuning knobs for experiments
        const long mb = std::atol(v);
        if (mb > 0) ctx->batch_bytes = (uint64_t)mb << 20;
    }
    {
        // host threads this context may use for packing: SG_HOST_THREADS, else all hardware threads shared by its GPUs
        int hw = (int)std::thread::hardware_concurrency();
        if (const char *v = std::getenv("SG_HOST_THREADS")) hw = std::max(1, std::atoi(v));
        ctx->host_threads = std::max(1, hw / n_devices);
        ctx->host_pack = ctx->host_threads >= 10;   // ~7-10 GB/s of ASCII per thread against ~47 GB/s of PCIe per GPU
        if (const char *v = std::getenv("SG_HOST_PACK")) ctx->host_pack = std::atoi(v) != 0;
        if (const char *v = std::getenv("SG_ASCII_MIN_BYTES")) ctx->ascii_min_bytes = (uint64_t)std::max(0ll, std::atoll(v));
        if (const char *v = std::getenv("SG_ASCII_PCT")) ctx->ascii_frac = std::min(100, std::max(0, std::atoi(v))) / 100.0;
        // SG_INGEST=adaptive (default) | fixed: the pre-adaptive policies (host packing with a fixed ASCII share when the
        // GPU has >= 10 host threads, else ASCII upload + device packing); SG_HOST_PACK / SG_ASCII_PCT imply fixed
        ctx->adaptive = !std::getenv("SG_HOST_PACK") && !std::getenv("SG_ASCII_PCT");
        if (const char *v = std::getenv("SG_INGEST")) ctx->adaptive = std::string(v) == "adaptive";
        if (const char *v = std::getenv("SG_CHUNK_KB")) ctx->chunk_bytes = std::max<uint64_t>(256, ((uint64_t)std::max(1ll, std::atoll(v)) << 10) & ~255ull);
        if (const char *v = std::getenv("SG_DMA_DEPTH")) ctx->dma_depth = std::min(kMaxDmaDepth, std::max(1, std::atoi(v)));
    }
    if (const char *v = std::getenv("SG_MAX_BATCH_MB")) {
        const long mb = std::atol(v);
        if (mb > 0) ctx->max_batch_bytes = (uint64_t)mb << 20;
    }
    for (int k = 0; k < n_devices; k++) {
        Device &d = ctx->devs[k];
        d.id = device_ids ? device_ids[k] : k;
        if (d.id < 0 || d.id >= avail) return fail(SG_ERR_BAD_ARG, "device id out of range");
        SG_CUDA(cudaSetDevice(d.id));
        if (const char *v = std::getenv("SG_SLOTS")) d.n_slots = std::min(kMaxSlots, std::max(2, std::atoi(v)));
        for (int q = 0; q < d.n_slots; q++) R(d.slots[q].create());
        int wps = 0, sms = 0;
        R(sg_dev_align_geometry_wo(W, O, &wps, nullptr, &sms));
        // one alignment per resident lane fills the device (104 192 lanes on a B200 at W=64)
        ctx->min_batch_units = std::max<uint64_t>(ctx->min_batch_units, 32ull * (uint64_t)wps * (uint64_t)sms);
    }
    if (const char *v = std::getenv("SG_MIN_BATCH_UNITS")) ctx->min_batch_units = (uint64_t)std::max(1ll, std::atoll(v));
    *out = ctx.release();
    return SG_OK;
}

void sg_ctx_destroy(sg_ctx *ctx)
{
    if (!ctx) return;
    for (Device &d : ctx->devs) {
        cudaSetDevice(d.id);
        for (Slot &s : d.slots) s.destroy();
        d.genome.release();
    }
    delete ctx;
}

int sg_ctx_num_devices(const sg_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

static int align_pairs_common(sg_ctx *ctx, const Strings &text, const Strings &query, uint64_t n_pairs, uint32_t flags, sg_result **out)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    Workload w;
    w.text = text; w.query = query; w.flags = flags;
    // sub-batches are cut by uploaded bytes (text + query), shards by the same weight
    std::unique_ptr<uint64_t[]> woff(new uint64_t[n_pairs + 1]);
    if (text.blob && query.blob) {   // a prefix sum of sizes is a difference of offsets: no serial pass over the pairs
        const uint64_t t0 = text.off[0], q0 = query.off[0];
        parallel_for(n_pairs + 1, ctx->host_threads, [&](uint64_t p0, uint64_t p1) {
            for (uint64_t p = p0; p < p1; p++) woff[p] = (text.off[p] - t0) + (query.off[p] - q0);
        });
    } else {
        woff[0] = 0;
        for (uint64_t p = 0; p < n_pairs; p++) woff[p + 1] = woff[p] + text.size(p) + query.size(p);
    }
    return run_all(ctx, w, woff.get(), 48, n_pairs, out);
}

int sg_align_pairs(sg_ctx *ctx, const char *text_blob, const uint64_t *text_off, const char *query_blob,
                   const uint64_t *query_off, uint64_t n_pairs, uint32_t flags, sg_result **out)
{
    if (!ctx || !out || !text_off || !query_off) return fail(SG_ERR_BAD_ARG, "sg_align_pairs: null argument");
    Strings t, q;
    t.blob = text_blob; t.off = text_off;
    q.blob = query_blob; q.off = query_off;
    return align_pairs_common(ctx, t, q, n_pairs, flags, out);
}

int sg_align_pairs_v(sg_ctx *ctx, const char *const *texts, const uint64_t *text_len, const char *const *queries,
                     const uint64_t *query_len, uint64_t n_pairs, uint32_t flags, sg_result **out)
{
    if (!ctx || !out || (n_pairs && (!texts || !text_len || !queries || !query_len)))
        return fail(SG_ERR_BAD_ARG, "sg_align_pairs_v: null argument");
    Strings t, q;
    t.ptr = texts; t.len = text_len;
    q.ptr = queries; q.len = query_len;
    static const char *const none[1] = {nullptr};
    static const uint64_t zero[1] = {0};
    if (!n_pairs) { t.ptr = q.ptr = none; t.len = q.len = zero; }
    return align_pairs_common(ctx, t, q, n_pairs, flags, out);
}

int sg_set_reference(sg_ctx *ctx, const char *genome_ascii, uint64_t genome_len)
{
    if (!ctx || (!genome_ascii && genome_len)) return fail(SG_ERR_BAD_ARG, "sg_set_reference: null argument");
    std::lock_guard<std::mutex> lock(ctx->mu);
    const int nd = (int)ctx->devs.size();
    std::vector<int> rcs(nd, SG_OK);
    std::vector<std::string> errs(nd);
    auto work = [&](int k) -> int {
        Device &d = ctx->devs[k];
        Slot &s = d.slots[0];
        SG_CUDA(cudaSetDevice(d.id));
        d.has_genome = false;
        R(d.genome.reserve(sg_packed_words(genome_len) * 4 + 64));
        R(s.bad.reserve(16));
        R(s.h_small.reserve(64));
        SG_CUDA(cudaMemsetAsync(s.bad.p, 0xFF, 16, s.stream));
        // upload in 256 Mbase pieces (a multiple of 16 bases, so every piece packs to whole words), alternating
        // between two staging buffers so that the copy of piece k+1 can overlap the packing of piece k
        const uint64_t piece = 256ull << 20;
        DevBuf *stage[2] = {&s.ascii_t, &s.ascii_q};
        R(stage[0]->reserve(std::min<uint64_t>(piece, genome_len) + 64));
        if (genome_len > piece) R(stage[1]->reserve(std::min<uint64_t>(piece, genome_len - piece) + 64));
        uint64_t pos = 0, first_bad = ~0ull;
        uint64_t *h = s.h_small.as<uint64_t>();
        int k2 = 0;
        do {
            const uint64_t len = std::min<uint64_t>(piece, genome_len - pos);
            SG_CUDA(cudaMemcpyAsync(stage[k2]->p, genome_ascii + pos, len, cudaMemcpyHostToDevice, s.stream));
            R(sg_dev_pack_2bit(stage[k2]->as<char>(), len, d.genome.as<uint32_t>() + pos / 16, s.bad.as<uint64_t>(), s.stream));
            SG_CUDA(cudaMemcpyAsync(h, s.bad.p, 8, cudaMemcpyDeviceToHost, s.stream));
            SG_CUDA(cudaStreamSynchronize(s.stream));
            if (h[0] != ~0ull) { first_bad = pos + h[0]; break; }
            pos += len;
            k2 ^= 1;
        } while (pos < genome_len);
        if (first_bad != ~0ull) return fail(SG_ERR_BAD_BASE, "non-ACGT character in reference at position " + std::to_string(first_bad));
        d.genome_len = genome_len;
        d.has_genome = true;
        return SG_OK;
    };
    std::vector<std::thread> th;
    for (int k = 0; k < nd; k++) th.emplace_back([&, k]() { rcs[k] = work(k); if (rcs[k]) errs[k] = g_last_error; });
    for (auto &t : th) t.join();
    for (int k = 0; k < nd; k++) if (rcs[k]) return fail(rcs[k], errs[k]);
    return SG_OK;
}

static int align_candidates_common(sg_ctx *ctx, const Strings &reads, uint64_t n_reads, const uint64_t *cand_start, const uint32_t *cand_read,
                                   uint64_t n_cand, uint32_t flags, sg_result **out)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    for (Device &d : ctx->devs)
        if (!d.has_genome) return fail(SG_ERR_NO_REFERENCE, "sg_align_candidates: call sg_set_reference first");
    const uint64_t genome_len = ctx->devs[0].genome_len;
    std::vector<uint64_t> woff(n_cand + 1, 0);  // weight of a candidate = its read's length
    for (uint64_t c = 0; c < n_cand; c++) {
        if (cand_read[c] >= n_reads) return fail(SG_ERR_BAD_ARG, "candidate " + std::to_string(c) + ": read index out of range");
        if (cand_start[c] > genome_len) return fail(SG_ERR_BAD_ARG, "candidate " + std::to_string(c) + ": start beyond the reference");
        woff[c + 1] = woff[c] + reads.size(cand_read[c]);
    }
    Workload w;
    w.mapping = true;
    w.query = reads; w.cand_start = cand_start; w.cand_read = cand_read; w.flags = flags;
    return run_all(ctx, w, woff.data(), 64, n_cand, out);
}

int sg_align_candidates(sg_ctx *ctx, const char *read_blob, const uint64_t *read_off, uint64_t n_reads,
                        const uint64_t *cand_start, const uint32_t *cand_read, uint64_t n_cand, uint32_t flags,
                        sg_result **out)
{
    if (!ctx || !out || !read_off || (n_cand && (!cand_start || !cand_read)))
        return fail(SG_ERR_BAD_ARG, "sg_align_candidates: null argument");
    Strings r;
    r.blob = read_blob; r.off = read_off;
    return align_candidates_common(ctx, r, n_reads, cand_start, cand_read, n_cand, flags, out);
}

int sg_align_candidates_v(sg_ctx *ctx, const char *const *reads, const uint64_t *read_len, uint64_t n_reads,
                          const uint64_t *cand_start, const uint32_t *cand_read, uint64_t n_cand, uint32_t flags,
                          sg_result **out)
{
    if (!ctx || !out || (n_reads && (!reads || !read_len)) || (n_cand && (!cand_start || !cand_read)))
        return fail(SG_ERR_BAD_ARG, "sg_align_candidates_v: null argument");
    Strings r;
    r.ptr = reads; r.len = read_len;
    return align_candidates_common(ctx, r, n_reads, cand_start, cand_read, n_cand, flags, out);
}

uint64_t sg_result_count(const sg_result *r) { return r ? r->n : 0; }
const int64_t *sg_result_edit_distances(const sg_result *r) { return r ? r->edit : nullptr; }
const uint64_t *sg_result_ref_consumed(const sg_result *r) { return r ? r->refc : nullptr; }
const uint64_t *sg_result_run_offsets(const sg_result *r) { return r && r->has_cigar ? r->run_off : nullptr; }

const uint8_t *sg_result_runs(const sg_result *r)
{
    if (!r || !r->has_cigar) return nullptr;
    if (r->pieces.size() == 1) return r->pieces[0].p;
    sg_result *m = const_cast<sg_result *>(r);
    if (m->flat.empty() && r->run_off[r->n]) {
        m->flat.resize(r->run_off[r->n]);
        for (size_t k = 0; k < r->pieces.size(); k++) memcpy(m->flat.data() + r->piece_run0[k], r->pieces[k].p, r->piece_runs[k]);
    }
    return m->flat.data();
}

int64_t sg_result_kernel_ns(const sg_result *r) { return r ? r->kernel_ns : 0; }
int64_t sg_result_total_ns(const sg_result *r) { return r ? r->total_ns : 0; }

// "%d%c" of one packed run (reference src/genasm_gpu.cu:881-888) from a 256-entry table: the text of run byte b padded
// to 4 characters, and its length (2 or 3; count 0 never occurs).  A run is rendered with one 4-byte store.
struct RunText {
    uint32_t text[256];
    uint8_t len[256];
    RunText()
    {
        static const char ops[4] = {'=', 'X', 'I', 'D'};
        for (unsigned b = 0; b < 256; b++) {
            const unsigned c = SG_RUN_COUNT(b);
            char t[4] = {0, 0, 0, 0};
            unsigned l = 0;
            if (c >= 10) t[l++] = (char)('0' + c / 10);
            t[l++] = (char)('0' + c % 10);
            t[l++] = ops[SG_RUN_OP(b)];
            memcpy(&text[b], t, 4);
            len[b] = (uint8_t)l;
        }
    }
};
static const RunText g_run_text;

static inline uint64_t runs_text_len(const uint8_t *p, uint64_t cnt)
{
    uint64_t len = 0;
    for (uint64_t k = 0; k < cnt; k++) len += g_run_text.len[p[k]];
    return len;
}

// renders cnt runs at o; the caller guarantees room for the text (+ nothing else): the last run is written byte by byte
static inline char *runs_render(const uint8_t *p, uint64_t cnt, char *o)
{
    if (!cnt) return o;
    for (uint64_t k = 0; k + 1 < cnt; k++) {   // a 4-byte store may spill 1-2 bytes into the next run's place: fine
        memcpy(o, &g_run_text.text[p[k]], 4);
        o += g_run_text.len[p[k]];
    }
    const uint8_t b = p[cnt - 1];
    memcpy(o, &g_run_text.text[b], g_run_text.len[b]);
    return o + g_run_text.len[b];
}

// Window configurations with W - O > 63: a run longer than 63 arrives as bytes with count 0 ("63 more of this op") followed
// by the byte with the rest.  fn(count, op) is called once per run, the pieces summed.
extern "C++" {
template <class F> static inline void for_each_wide_run(const uint8_t *p, uint64_t cnt, F &&fn)
{
    unsigned carry = 0;
    for (uint64_t k = 0; k < cnt; k++) {
        const unsigned c = SG_RUN_COUNT(p[k]);
        if (c == 0) { carry += 63; continue; }
        fn(carry + c, SG_RUN_OP(p[k]));
        carry = 0;
    }
}

static inline uint64_t wide_text_len(const uint8_t *p, uint64_t cnt)
{
    uint64_t len = 0;
    for_each_wide_run(p, cnt, [&](unsigned c, unsigned) { len += c >= 100 ? 4 : (c >= 10 ? 3 : 2); });
    return len;
}

static inline char *wide_render(const uint8_t *p, uint64_t cnt, char *o)
{
    static const char ops[4] = {'=', 'X', 'I', 'D'};
    for_each_wide_run(p, cnt, [&](unsigned c, unsigned op) {
        if (c >= 100) *o++ = (char)('0' + c / 100);
        if (c >= 10) *o++ = (char)('0' + (c / 10) % 10);
        *o++ = (char)('0' + c % 10);
        *o++ = ops[op];
    });
    return o;
}
}  // extern "C++"

uint64_t sg_result_cigar_len(const sg_result *r, uint64_t idx)
{
    if (!r || !r->has_cigar || idx >= r->n) return 0;
    uint64_t cnt;
    const uint8_t *p = runs_of(r, idx, &cnt);
    return r->wide_runs ? wide_text_len(p, cnt) : runs_text_len(p, cnt);
}

int64_t sg_result_render_cigar(const sg_result *r, uint64_t idx, char *buf, uint64_t cap)
{
    if (!r || !r->has_cigar || idx >= r->n || !buf) return -1;
    uint64_t cnt;
    const uint8_t *p = runs_of(r, idx, &cnt);
    // 3 characters per run at most: when the buffer is that large no length pass is needed
    // (a run of several bytes renders to at most 4 characters: the bound holds for them too)
    const uint64_t len = 3 * cnt + 1 <= cap ? 0 : (r->wide_runs ? wide_text_len(p, cnt) : runs_text_len(p, cnt));
    if (len + 1 > cap) return -1;
    char *end = r->wide_runs ? wide_render(p, cnt, buf) : runs_render(p, cnt, buf);
    *end = '\0';
    return (int64_t)(end - buf);
}

int64_t sg_result_entries(const sg_result *r, uint64_t idx, sg_cigar_entry *out, uint64_t cap)
{
    if (!r || !r->has_cigar || idx >= r->n || !out) return -1;
    static const char ops[4] = {'=', 'X', 'I', 'D'};
    uint64_t cnt;
    const uint8_t *p = runs_of(r, idx, &cnt);
    if (cnt > cap) return -1;
    if (r->wide_runs) {   // runs of up to 127 still fit the reference's uint8 count (src/util.hpp:43-46)
        uint64_t m = 0;
        for_each_wide_run(p, cnt, [&](unsigned c, unsigned op) { out[m].edit_count = (uint8_t)c; out[m].edit_type = ops[op]; m++; });
        return (int64_t)m;
    }
    for (uint64_t k = 0; k < cnt; k++) {
        out[k].edit_count = (uint8_t)SG_RUN_COUNT(p[k]);
        out[k].edit_type = ops[SG_RUN_OP(p[k])];
    }
    return (int64_t)cnt;
}

uint64_t sg_result_render_all(const sg_result *r, char *blob, uint64_t blob_cap, uint64_t *text_off, int threads)
{
    // "%d%c" per run (reference src/genasm_gpu.cu:881-888), all alignments, all host threads: two passes over the
    // packed runs (lengths, then characters).  The reference renders one alignment at a time through a stringstream.
    if (!r || !text_off) return 0;
    const uint64_t n = r->n;
    if (threads < 1) threads = (int)std::max(1u, std::thread::hardware_concurrency());
    text_off[0] = 0;
    if (!r->has_cigar) {
        for (uint64_t a = 0; a < n; a++) text_off[a + 1] = 0;
        return 0;
    }
    auto parallel = [&](auto &&fn) {
        const int nt = (int)std::min<uint64_t>((uint64_t)threads, std::max<uint64_t>(1, n / 1024));
        if (nt <= 1) { fn(0, n); return; }
        std::vector<std::thread> th;
        for (int t = 0; t < nt; t++) th.emplace_back([&, t]() { fn(n * t / nt, n * (t + 1) / nt); });
        for (auto &x : th) x.join();
    };
    parallel([&](uint64_t a0, uint64_t a1) {
        for (uint64_t a = a0; a < a1; a++) {
            uint64_t cnt;
            const uint8_t *p = runs_of(r, a, &cnt);
            text_off[a + 1] = r->wide_runs ? wide_text_len(p, cnt) : runs_text_len(p, cnt);
        }
    });
    for (uint64_t a = 0; a < n; a++) text_off[a + 1] += text_off[a];
    const uint64_t total = text_off[n];
    if (!blob || blob_cap < total) return total;  // sizes only: call again with a big enough blob
    parallel([&](uint64_t a0, uint64_t a1) {
        for (uint64_t a = a0; a < a1; a++) {
            uint64_t cnt;
            const uint8_t *p = runs_of(r, a, &cnt);
            if (r->wide_runs) wide_render(p, cnt, blob + text_off[a]);
            else runs_render(p, cnt, blob + text_off[a]);
        }
    });
    return total;
}

void sg_result_free(sg_result *r) { delete r; }

void *sg_host_alloc(uint64_t bytes)
{
    void *p = nullptr;
    if (cudaMallocHost(&p, bytes ? bytes : 1) != cudaSuccess) {
        cudaGetLastError();
        fail(SG_ERR_OOM, "cudaMallocHost failed for " + std::to_string(bytes) + " bytes");
        return nullptr;
    }
    return p;
}

void sg_host_free(void *p)
{
    if (p) cudaFreeHost(p);
}

}  // extern "C"
