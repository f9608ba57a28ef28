// sg_host_api.cu -- layer 1 of include/scrooge_b200.h: host buffers in, host results out.
//
// Replaces the host side of the reference GPU library (src/genasm_gpu.cu:692-1065): cudaMallocManaged
// blobs, per-string descriptor loops, one synchronous kernel over everything and a linked-list walk per
// alignment become
//   * a call cut into sub-batches that form ONE queue; every GPU of the context has a worker thread that takes
//     sub-batches from it (alignments are independent, src/genasm_cpu.cpp:451-455: no inter-GPU exchange of any
//     kind; a GPU that finishes early takes more, like the reference's atomic pair counter across thread blocks,
//     src/genasm_gpu.cu:602-622).  In mapping mode every GPU holds its own packed reference;
//   * per GPU, a five-slot software pipeline: while sub-batch k is being aligned, k+1 is on its way over PCIe and
//     the distances and compacted CIGAR runs of k-1 are on their way back into pinned host memory.  The worker never
//     blocks on the device while there is something to upload;
//   * ingest shared between the GPU's packer threads (a persistent team bound to the GPU's CPUs: chunks packed to
//     2 bit/base on the host, a quarter of the bytes cross PCIe) and its copy engine (chunks cross as ASCII, the device
//     packs them), as many packers as the measured ingest rate says.  EVERY host-to-device copy of a GPU goes through
//     one stream: a flooded copy-engine channel starves the host-to-device copies of other streams (Device::h2d);
//   * descriptors built on the host, a run slab with per-alignment capacity 2*|query|+8 (reference: 2*|query| entries,
//     src/genasm_gpu.cu:995-1001) compacted on the device; launches of mixed lengths handed out longest first.
// The end-to-end path is bound by the host's memory system; sg_result_stats and SG_TRACE=1 say
// where a call's time went.
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
#include <atomic>
#include <omp.h>
#include <unistd.h>
#include <cuda_runtime.h>

#include "../../include/scrooge_b200.h"
#include "sg_internal.h"
#include "sg_host_threads.h"

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
    // How much page-locked memory the pool keeps for reuse after the results that used it were freed.  Page-locked memory
    // cannot be swapped and every process has its own pool, so the cache follows what the recent calls actually used:
    // limit = max(base, peak of the bytes handed out at the same time over the recent calls), never more than a quarter
    // of the machine's RAM; the peak halves with every call that does not reach it, so one large call does not pin its
    // footprint for the life of the process.  base = min(RAM/32, 4 GB).  Re-pinning is what the cache avoids: ~0.3 s per
    // GB, i.e. more than the whole call for a read-mapping batch that returns 4 GB of runs.  SG_PINNED_CACHE_GB=<n> fixes
    // the limit (0 = keep nothing); sg_trim_host_cache() and the destruction of the last context release everything.
    size_t ram_bytes() const
    {
        const long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
        return pages > 0 && psz > 0 ? (size_t)pages * (size_t)psz : (size_t)64 << 30;
    }
    const long long kFixedCap = [] { const char *v = std::getenv("SG_PINNED_CACHE_GB"); return v ? std::max(0ll, std::atoll(v)) << 30 : -1ll; }();
    const size_t kBaseCap = std::min<size_t>(ram_bytes() / 32, (size_t)4 << 30);
    const size_t kHardCap = ram_bytes() / 4;
    size_t live = 0, peak_live = 0;   // bytes handed out now / their recent peak (under mu)
    size_t limit() const { return kFixedCap >= 0 ? (size_t)kFixedCap : std::min(kHardCap, std::max(kBaseCap, peak_live)); }
    void call_done()   // once per call: let the peak decay towards what is in use
    {
        std::lock_guard<std::mutex> g(mu);
        peak_live = std::max(live, peak_live / 2);
    }

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
                live += out->cap;
                peak_live = std::max(peak_live, live);
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
        std::lock_guard<std::mutex> g(mu);
        live += want;
        peak_live = std::max(peak_live, live);
        return SG_OK;
    }
    void release(Block b)
    {
        if (!b.p) return;
        std::lock_guard<std::mutex> g(mu);
        live -= std::min(live, b.cap);
        if (cached + b.cap > limit()) { cudaFreeHost(b.p); return; }
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

// Where the host time and the PCIe bytes of a call go, per GPU of the context (one writer: the GPU's own worker thread;
// the packer threads add their share through atomics).  Merged into sg_call_stats by run_all.
struct ShardStats {
    double upload = 0;      // wall time of the ingest phase: host packing + queuing of the copies, all sub-batches
    double wait = 0;        // blocked on the device (alignment kernel, compaction, copies back)
    double desc = 0;        // descriptor arithmetic
    double out = 0;         // results into the caller-visible block
    std::atomic<uint64_t> pack_thread_ns{0};   // summed over the packer threads: time inside the packing loop
    std::atomic<uint64_t> h2d_packed{0};       // bytes that crossed PCIe packed by the host (2 bit/base)
    uint64_t h2d_ascii = 0, h2d_other = 0, d2h = 0;
    uint32_t sub_batches = 0;
    uint32_t packers_in_use = 0;   // packer threads the ingest tuner used for the last sub-batch
};
static const bool g_debug = std::getenv("SG_DEBUG") != nullptr;
// SG_TRACE=1: one line per pipeline event on stderr (milliseconds since the call began) -- where a call's time goes
static const bool g_trace = std::getenv("SG_TRACE") != nullptr;
static std::chrono::steady_clock::time_point g_trace_t0;
static void trace(int dev, uint64_t first, uint64_t count, const char *what, double extra = -1)
{
    if (!g_trace) return;
    const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - g_trace_t0).count();
    if (extra >= 0) fprintf(stderr, "[sg-trace] %9.3f ms  gpu %d  batch %llu+%llu  %s %.3f\n", ms, dev, (unsigned long long)first, (unsigned long long)count, what, extra);
    else fprintf(stderr, "[sg-trace] %9.3f ms  gpu %d  batch %llu+%llu  %s\n", ms, dev, (unsigned long long)first, (unsigned long long)count, what);
}
struct ScopedT {
    double &acc; std::chrono::steady_clock::time_point t0;
    explicit ScopedT(double &a) : acc(a), t0(std::chrono::steady_clock::now()) {}
    ~ScopedT() { acc += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

// Per-alignment host loops (descriptor arithmetic): serial up to a quarter of a million alignments -- a sub-batch of long reads --
// and split over plain threads (131 072 alignments each) above that (batches of short reads).  Deliberately not OpenMP: an OpenMP
// team that fits the cores spin-waits after its region and delays the CUDA calls that follow.  The threads run on every CPU
// the process may use: a new thread inherits the affinity of its creator, and the creator is a GPU's worker, which is
// bound to ONE CPU for the duration of a call -- without the reset all of them would share that CPU.
static const std::vector<int> &process_cpus()
{
    static const std::vector<int> cpus = allowed_cpus();   // first use is in sg_ctx_create, before any worker is bound
    return cpus;
}

template <class F> void parallel_for(uint64_t n, int threads, F &&fn)
{
    const int nt = (int)std::min<uint64_t>((uint64_t)std::max(1, threads), n >> 17);
    if (nt <= 1) { fn((uint64_t)0, n); return; }
    std::vector<std::thread> th;
    for (int t = 0; t < nt; t++)
        th.emplace_back([&fn, n, t, nt]() {
            bind_this_thread(process_cpus());
            fn(n * (uint64_t)t / nt, n * (uint64_t)(t + 1) / nt);
        });
    for (auto &x : th) x.join();
}

extern "C" uint64_t sg_host_pack_2bit_st(const char *ascii, uint64_t n_bases, uint32_t *packed);

constexpr int kMaxSlots = 8;
constexpr int kMaxDmaDepth = 16;

// One pipeline stage's worth of buffers: a sub-batch lives in a slot from upload to download.
struct Slot {
    cudaStream_t stream = nullptr;
    cudaEvent_t ev_copied = nullptr;   // recorded in the device's host-to-device stream after the sub-batch's last copy
    cudaEvent_t ev_k0 = nullptr, ev_k1 = nullptr, ev_mid = nullptr, ev_end = nullptr;
    cudaEvent_t ev_dma[kMaxDmaDepth] = {};   // adaptive ingest: one per ASCII chunk copy in flight
    DevBuf ascii_t, ascii_q, packed_t, packed_q, desc, slab, counter, edit, refc, nruns, status, run_off, scan_tmp, runs, bad, order;
    PinBuf h_small, h_status, h_stage_t, h_stage_q, h_desc, h_order;
    PinnedPool::Block piece{nullptr, 0};
    // the batch in flight
    bool busy = false, mid_done = false;
    uint64_t a0 = 0, a1 = 0, total_runs = 0;
    uint64_t bad_bias[2] = {0, 0};   // hybrid ingest: the device-packed tail of a blob starts at this base
    int create()
    {
        SG_CUDA(cudaStreamCreateWithFlags(&stream, cudaStreamNonBlocking));
        SG_CUDA(cudaEventCreateWithFlags(&ev_copied, cudaEventDisableTiming));
        SG_CUDA(cudaEventCreate(&ev_k0));
        SG_CUDA(cudaEventCreate(&ev_k1));
        SG_CUDA(cudaEventCreateWithFlags(&ev_mid, cudaEventDisableTiming));
        SG_CUDA(cudaEventCreateWithFlags(&ev_end, cudaEventDisableTiming));
        // polled by the feeder of the adaptive ingest (a blocking wait costs an interrupt and a wake-up per chunk, which a
        // virtual machine turns into gaps on the PCIe link)
        for (cudaEvent_t &e : ev_dma) SG_CUDA(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        return SG_OK;
    }
    void destroy()
    {
        for (DevBuf *b : {&ascii_t, &ascii_q, &packed_t, &packed_q, &desc, &slab, &counter, &edit, &refc, &nruns, &status, &run_off,
                          &scan_tmp, &runs, &bad, &order})
            b->release();
        for (PinBuf *b : {&h_small, &h_status, &h_stage_t, &h_stage_q, &h_desc, &h_order}) b->release();
        g_pool.release(piece);
        piece = {nullptr, 0};
        if (ev_k0) cudaEventDestroy(ev_k0);
        if (ev_k1) cudaEventDestroy(ev_k1);
        if (ev_mid) cudaEventDestroy(ev_mid);
        if (ev_end) cudaEventDestroy(ev_end);
        for (cudaEvent_t e : ev_dma) if (e) cudaEventDestroy(e);
        if (ev_copied) cudaEventDestroy(ev_copied);
        if (stream) cudaStreamDestroy(stream);
    }
};

struct Device {
    int id = 0;
    // EVERY host-to-device copy of this GPU goes through this one stream, whatever sub-batch it belongs to.  The copy
    // engine serves the channel of a stream that keeps it busy and does not switch away while that channel has work: with
    // the ASCII chunk copies of sub-batch k+1 back to back in one stream, a host-to-device copy queued in ANOTHER stream
    // (the descriptors or the last packed chunks of sub-batch k, in front of its kernels) was not executed until the
    // host stopped feeding -- measured with tools/ce_probe.cu: "H2D 8 MB in B, then kernel in B" is not ready after 400 ms
    // of flooding stream A, while the same copy queued in A with an event for B's kernel is ready after 6.4 ms
    // (profiles/r02_copy_engine_probe.txt).  Kernels and device-to-host copies of other streams are not affected.  One
    // stream = one FIFO: sub-batches hand over to their own (compute) stream through ev_copied.
    cudaStream_t h2d = nullptr;
    IngestTuner tuner;
    Slot slots[kMaxSlots];
    int n_slots = 5;
    DevBuf genome;  // packed reference, resident across calls
    DevBuf piece_bad;   // sg_set_reference: one offending-base word per uploaded piece
    uint64_t genome_len = 0;
    bool has_genome = false;
    std::vector<int> cpus;   // the CPUs this GPU's host threads run on (sg_host_threads.h)
    // The worker (the feeder of the adaptive ingest: it keeps the copy engine's queue full and polls for finished kernels)
    // gets a CPU of its own, the packers share the rest: with packers on every CPU the feeder was descheduled for
    // milliseconds at a time and the PCIe link ran dry (13 GB/s of ASCII chunk copies where 24 GB/s fit).
    std::vector<int> worker_cpus, packer_cpus;
    // set by the worker for the duration of a call: starts the second stage of earlier sub-batches whose kernels are done.
    // The feeder calls it between chunk copies, so that a sub-batch's runs are on their way back while the next one is
    // still being uploaded (and its slot is free when the pipeline comes round to it).
    std::function<int()> idle_poll;
    ThreadTeam team;         // its packer threads: created with the context, asleep between jobs
};

// Devices own threads and streams: neither copied nor moved, so a fixed array instead of a std::vector.
struct DeviceList {
    std::unique_ptr<Device[]> p;
    size_t n = 0;
    void resize(size_t k) { p.reset(new Device[k]); n = k; }
    size_t size() const { return n; }
    Device &operator[](size_t k) { return p[k]; }
    const Device &operator[](size_t k) const { return p[k]; }
    Device *begin() { return p.get(); }
    Device *end() { return p.get() + n; }
    const Device *begin() const { return p.get(); }
    const Device *end() const { return p.get() + n; }
};

// The same split over the GPU's packer team: its threads exist, are bound to the GPU's CPUs and are idle whenever the
// worker runs descriptor loops (creating and joining plain threads cost ~0.3 ms per loop, five loops per sub-batch:
// a quarter of a 10 M x 150 bp call).
template <class F> void team_for(Device &d, uint64_t n, F &&fn)
{
    const int nt = d.team.size();
    if (nt <= 1 || n < (1u << 16)) { fn((uint64_t)0, n); return; }
    std::function<void(int)> job = [&](int t) { fn(n * (uint64_t)t / (uint64_t)nt, n * (uint64_t)(t + 1) / (uint64_t)nt); };
    d.team.launch(job);
    d.team.wait();
}

}  // namespace sg

using namespace sg;

struct sg_ctx {
    int W = 64;
    int O = 33;
    DeviceList devs;
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
    int dma_depth = 12;
    uint64_t ascii_min_bytes = 8ull << 20;   // blobs smaller than this are not split
    std::mutex mu;  // calls on one context are serialised
    bool taper = true;           // SG_TAPER=0: the last sub-batch of a call is not cut finer
    bool longest_first = true;   // SG_LONGEST_FIRST=0: launches take their alignments in input order
    int emit = 1;                // how the kernel stores runs: 1 = as whole words (SG_FLAG_RUN_WORDS), 0 = as bytes (SG_EMIT=bytes)
    bool counted = false;   // fully created (sg_ctx_destroy also cleans up after a failed creation)
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
    sg_call_stats stats;
    ~sg_result() { for (auto &b : pieces) g_pool.release(b); g_pool.release(store); }
};

namespace {

struct ShardOut {
    int rc = SG_OK;
    std::string err;
    double kernel_ms = 0;
    ShardStats stats;
    std::vector<PinnedPool::Block> pieces;
    std::vector<uint64_t> piece_first, piece_runs;
    ~ShardOut() { for (auto &b : pieces) g_pool.release(b); }
};

// n strings, given either as one blob + n+1 offsets or as n pointers + n lengths (no flattening needed)
struct Strings {
    const char *blob = nullptr; const uint64_t *off = nullptr;
    const char *const *ptr = nullptr; const uint64_t *len = nullptr;
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

// One blob range of a sub-batch on its way to the device.
struct Segment {
    const char *src = nullptr; uint64_t nbytes = 0;
    DevBuf *d_ascii = nullptr, *d_packed = nullptr; PinBuf *h_stage = nullptr; uint64_t *d_bad = nullptr;
    long long chunk0 = 0, nch = 0;   // its chunks in the sub-batch's chunk list
    uint64_t split = 0;              // out: [0, split) was packed by the host, [split, nbytes) crossed as ASCII
};

// Adaptive ingest of a sub-batch's blobs (see sg_ctx::adaptive): the blobs are cut into chunks, one list over all
// segments.  The GPU's packer threads take chunks from the FRONT (pack to 2 bit/base into pinned staging, upload the
// packed chunk right away); the calling thread -- the GPU's worker -- is the feeder: it keeps `dma_depth` ASCII chunk
// copies from the BACK in flight, and the device packs whatever arrived as ASCII.  Whoever is faster takes more.
// host_threads == 0: everything crosses as ASCII; dma_depth == 0: everything is packed by the host.
int upload_adaptive(sg_ctx *ctx, Device &d, cudaStream_t st, cudaEvent_t ev_copied, cudaEvent_t *ev_dma, Segment *seg, int nseg,
                    ShardStats &cs, uint64_t *bad_pos, int *bad_seg)
{
    cudaStream_t hst = d.h2d;   // every host-to-device copy of the GPU: see Device::h2d
    const uint64_t C = ctx->chunk_bytes;   // a multiple of 256: chunks start on whole packed words and whole output cache lines
    long long nch = 0;
    for (int k = 0; k < nseg; k++) {
        Segment &g = seg[k];
        const uint64_t words = sg_packed_words(g.nbytes);
        R(g.d_packed->reserve(words * 4));
        if (d.team.size() > 0) R(g.h_stage->reserve(words * 4 + 64));
        if (ctx->dma_depth > 0 || d.team.size() == 0) R(g.d_ascii->reserve(g.nbytes + 64));
        g.chunk0 = nch;
        g.nch = (long long)((g.nbytes + C - 1) / C);
        nch += g.nch;
    }
    const int depth = d.team.size() == 0 ? std::max(1, std::min(kMaxDmaDepth, ctx->dma_depth)) : std::min(kMaxDmaDepth, ctx->dma_depth);
    auto locate = [&](long long c, uint64_t *off, uint64_t *len) -> Segment & {
        int k = 0;
        while (k + 1 < nseg && c >= seg[k + 1].chunk0) k++;
        *off = (uint64_t)(c - seg[k].chunk0) * C;
        *len = std::min(C, seg[k].nbytes - *off);
        return seg[k];
    };
    std::mutex mu;
    long long front = 0, back = nch;   // packers take chunk `front++`, the feeder chunk `--back`
    auto take_front = [&]() -> long long { std::lock_guard<std::mutex> g(mu); return front < back ? front++ : -1; };
    auto take_back = [&]() -> long long { std::lock_guard<std::mutex> g(mu); return back > front ? --back : -1; };
    std::atomic<uint64_t> bad{~0ull};      // (segment << 56) | position: the smallest wins
    std::atomic<int> cuda_rc{0};
    uint64_t total_bytes = 0;
    for (int k = 0; k < nseg; k++) total_bytes += seg[k].nbytes;
    const int active = ctx->dma_depth > 0 ? d.tuner.packers(d.team.size(), total_bytes) : d.team.size();
    cs.packers_in_use = (uint32_t)std::max(0, active);
    const auto t_ingest = std::chrono::steady_clock::now();
    std::function<void(int)> job = [&](int tid) {
        if (tid >= active) return;
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t sent = 0;
        while (true) {
            const long long c = take_front();
            if (c < 0) break;
            uint64_t off, len;
            Segment &g = locate(c, &off, &len);
            uint32_t *hp = g.h_stage->as<uint32_t>() + off / 16;
            const uint64_t r = sg_host_pack_2bit_st(g.src + off, len, hp);
            if (r != ~0ull) {
                const uint64_t key = ((uint64_t)(&g - seg) << 56) | (off + r);
                uint64_t cur = bad.load();
                while (key < cur && !bad.compare_exchange_weak(cur, key)) {}
                break;
            }
            const uint64_t bytes = ((len + 15) / 16) * 4;
            if (cudaMemcpyAsync(g.d_packed->as<uint32_t>() + off / 16, hp, bytes, cudaMemcpyHostToDevice, hst) != cudaSuccess) { cuda_rc = 1; break; }
            sent += bytes;
        }
        cs.h2d_packed += sent;
        cs.pack_thread_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    };
    if (d.team.size() > 0) d.team.launch(job);
    if (depth > 0) {
        // The feeder claims a chunk for the copy engine only while that does not take work the packers would finish
        // sooner: a claimed chunk waits behind the copies already in flight (~0.3 ms each) while `active` packers get
        // through ~one chunk per millisecond each.  With 2 GB sub-batches the rule never binds; with the 256 MB
        // sub-batches of short reads a feeder that queued its 12 copies up front left the packers idle two thirds of
        // the time (10 M x 150 bp: ingest 53 ms with 275 of 795 packer thread-ms busy).
        const long long guard = std::max<long long>(1, active / 4);
        long long issued = 0, completed = 0;
        int polls = 0;
        auto pause = [] { for (int spin = 0; spin < 64; spin++) __builtin_ia32_pause(); };
        while (true) {
            while (completed < issued) {   // retire finished copies (events are reused modulo depth, oldest first)
                const cudaError_t qe = cudaEventQuery(ev_dma[completed % depth]);
                if (qe == cudaErrorNotReady) break;
                if (qe != cudaSuccess) { cuda_rc = 1; break; }
                completed++;
            }
            if (cuda_rc) break;
            const long long inflight = issued - completed;
            long long remaining;
            { std::lock_guard<std::mutex> g(mu); remaining = back - front; }
            if (remaining <= 0) break;
            if (inflight >= depth || (active > 0 && inflight > 0 && remaining < guard * (inflight + 1))) {
                if (d.idle_poll && (++polls & 63) == 0 && d.idle_poll() != SG_OK) { cuda_rc = 2; break; }
                pause();
                continue;
            }
            const long long c = take_back();
            if (c < 0) break;
            uint64_t off, len;
            Segment &g = locate(c, &off, &len);
            if (cudaMemcpyAsync(g.d_ascii->as<char>() + off, g.src + off, len, cudaMemcpyHostToDevice, hst) != cudaSuccess ||
                cudaEventRecord(ev_dma[issued % depth], hst) != cudaSuccess) { cuda_rc = 1; break; }
            cs.h2d_ascii += len;
            issued++;
            if (d.idle_poll && (issued & 7) == 0 && d.idle_poll() != SG_OK) { cuda_rc = 2; break; }
        }
    }
    if (d.team.size() > 0) d.team.wait();
    // the sub-batch's own stream (the device packs its ASCII part there) continues once its last chunk has arrived
    if (cudaEventRecord(ev_copied, hst) != cudaSuccess || cudaStreamWaitEvent(st, ev_copied, 0) != cudaSuccess) cuda_rc = 1;
    d.tuner.report(total_bytes, std::chrono::duration<double>(std::chrono::steady_clock::now() - t_ingest).count());
    if (cuda_rc == 2) return SG_ERR_CUDA;   // a second stage failed inside the poll: its message stands
    if (cuda_rc) { cudaGetLastError(); return fail(SG_ERR_CUDA, "adaptive ingest: a CUDA call failed"); }
    if (bad.load() != ~0ull) {   // the caller names the offender
        *bad_seg = (int)(bad.load() >> 56);
        *bad_pos = bad.load() & ((1ull << 56) - 1);
        return SG_OK;
    }
    *bad_pos = ~0ull;
    for (int k = 0; k < nseg; k++) {   // chunks [0, back) were packed by the host, [back, nch) arrived as ASCII
        Segment &g = seg[k];
        const long long host_chunks = std::min(g.nch, std::max(0ll, back - g.chunk0));
        g.split = std::min(g.nbytes, (uint64_t)host_chunks * C);
        if (g.split < g.nbytes) {
            R(sg_dev_pack_2bit_ex(g.d_ascii->as<char>() + g.split, g.nbytes - g.split, g.d_packed->as<uint32_t>() + g.split / 16, g.d_bad, SG_PACK_SIDE, st));
        } else {
            const uint64_t words = sg_packed_words(g.nbytes), used = (g.nbytes + 15) / 16;
            SG_CUDA(cudaMemsetAsync(g.d_packed->as<uint32_t>() + used, 0, (words - used) * 4, st));  // padding words the aligner may read
        }
    }
    return SG_OK;
}

// Separate strings [i0, i1) (pointer + length each, e.g. the std::strings of the C++ drop-in) -> packed words on the device,
// every string at a word boundary; start[k] receives the first base of string i0+k in the packed blob.  The strings live in
// pageable memory wherever the caller allocated them, so the copy engine cannot fetch them: the packer threads pack them
// where they lie (reading a string once costs the same as gathering it into pinned staging would) in chunks of strings, and
// every packed chunk is uploaded as soon as it is done.
int upload_separate(sg_ctx *ctx, Device &d, cudaStream_t st, const Strings &S, uint64_t i0, uint64_t i1, const char *what, const char *unit,
                    DevBuf &d_packed, PinBuf &h_stage, uint64_t *start, ShardStats &cs)
{
    const uint64_t n = i1 - i0;
    uint64_t w = 0;
    for (uint64_t k = 0; k < n; k++) { start[k] = w * 16; w += (S.len[i0 + k] + 15) / 16; }
    const uint64_t words = w + 8;
    R(d_packed.reserve(words * 4));
    R(h_stage.reserve(words * 4));
    uint32_t *hp = h_stage.as<uint32_t>();
    memset(hp + w, 0, 8 * 4);   // padding words the aligner may read
    // chunks of strings worth about chunk_bytes of ASCII each
    std::vector<uint64_t> cut{0};
    {
        uint64_t acc = 0;
        for (uint64_t k = 0; k < n; k++) {
            acc += S.len[i0 + k];
            if (acc >= ctx->chunk_bytes) { cut.push_back(k + 1); acc = 0; }
        }
        if (cut.back() != n) cut.push_back(n);
    }
    const long long nch = (long long)cut.size() - 1;
    std::atomic<long long> next{0};
    std::atomic<uint64_t> bad_k{~0ull};
    std::atomic<int> cuda_rc{0};
    std::function<void(int)> job = [&](int) {
        const auto t0 = std::chrono::steady_clock::now();
        uint64_t sent = 0;
        while (true) {
            const long long c = next.fetch_add(1);
            if (c >= nch) break;
            const uint64_t k0 = cut[c], k1 = cut[c + 1];
            bool ok = true;
            for (uint64_t k = k0; k < k1 && ok; k++)
                if (sg_host_pack_2bit_st(S.ptr[i0 + k], S.len[i0 + k], hp + start[k] / 16) != ~0ull) {
                    uint64_t cur = bad_k.load();
                    while (k < cur && !bad_k.compare_exchange_weak(cur, k)) {}
                    ok = false;
                }
            if (!ok) break;
            const uint64_t w0 = start[k0] / 16, w1 = k1 < n ? start[k1] / 16 : words;
            if (cudaMemcpyAsync(d_packed.as<uint32_t>() + w0, hp + w0, (w1 - w0) * 4, cudaMemcpyHostToDevice, d.h2d) != cudaSuccess) { cuda_rc = 1; break; }
            sent += (w1 - w0) * 4;
        }
        cs.h2d_packed += sent;
        cs.pack_thread_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count();
    };
    if (d.team.size() > 0) { d.team.launch(job); d.team.wait(); }
    else job(0);
    if (cuda_rc) { cudaGetLastError(); return fail(SG_ERR_CUDA, "ingest: a CUDA call failed"); }
    if (bad_k.load() != ~0ull) {
        const uint64_t k = bad_k.load();
        std::vector<uint32_t> tmp((S.len[i0 + k] + 15) / 16 + 1);
        return bad_base_error(what, unit, i0 + k, sg_host_pack_2bit_st(S.ptr[i0 + k], S.len[i0 + k], tmp.data()));
    }
    return SG_OK;
}

// Blob strings [i0, i1) of up to two inputs -> packed words on the device.
//   adaptive (default): upload_adaptive over all segments at once;
//   fixed policies (SG_INGEST=fixed, A/B experiments): host_pack = the packer threads pack the head of every blob and
//   ascii_frac of its tail crosses as ASCII; else everything crosses as ASCII and pack_2bit_kernel packs it.
// An offending base found on the device side is reported later through d_bad (+ bad_bias).
struct BlobIn { const Strings *S; uint64_t i0, i1; const char *what, *unit; DevBuf *d_ascii, *d_packed; PinBuf *h_stage; uint64_t *d_bad, *start, *bad_bias; };

int upload_blobs(sg_ctx *ctx, Device &d, cudaStream_t st, cudaEvent_t ev_copied, cudaEvent_t *ev_dma, BlobIn *in, int nin, ShardStats &cs)
{
    cudaStream_t hst = d.h2d;
    auto handover = [&]() -> int {   // st continues after everything queued in hst so far
        SG_CUDA(cudaEventRecord(ev_copied, hst));
        SG_CUDA(cudaStreamWaitEvent(st, ev_copied, 0));
        return SG_OK;
    };
    Segment seg[2];
    int nseg = 0, seg_of[2] = {-1, -1};
    for (int q = 0; q < nin; q++) {
        BlobIn &b = in[q];
        const Strings &S = *b.S;
        const uint64_t n = b.i1 - b.i0, base = S.off[b.i0], nbytes = S.off[b.i1] - base;
        *b.bad_bias = 0;
        team_for(d, n, [&](uint64_t k0, uint64_t k1) { for (uint64_t k = k0; k < k1; k++) b.start[k] = S.off[b.i0 + k] - base; });
        if (ctx->adaptive && nbytes >= ctx->ascii_min_bytes && nbytes >= 256) {
            Segment &g = seg[nseg];
            g.src = S.blob + base; g.nbytes = nbytes; g.d_ascii = b.d_ascii; g.d_packed = b.d_packed; g.h_stage = b.h_stage; g.d_bad = b.d_bad;
            seg_of[nseg++] = q;
            continue;
        }
        const uint64_t words = sg_packed_words(nbytes);
        R(b.d_packed->reserve(words * 4));
        if (!ctx->host_pack || d.team.size() == 0) {
            R(b.d_ascii->reserve(nbytes + 64));
            SG_CUDA(cudaMemcpyAsync(b.d_ascii->p, S.blob + base, nbytes, cudaMemcpyHostToDevice, hst));
            cs.h2d_ascii += nbytes;
            R(handover());
            R(sg_dev_pack_2bit(b.d_ascii->as<char>(), nbytes, b.d_packed->as<uint32_t>(), b.d_bad, st));
            continue;
        }
        R(b.h_stage->reserve(words * 4));
        uint32_t *hp = b.h_stage->as<uint32_t>();
        // hybrid: bases [split, nbytes) travel as ASCII (the copy is queued first, so the copy engine works while the
        // host packs [0, split)) and are packed on the device into the words that follow the host's
        uint64_t split = nbytes;
        if (ctx->ascii_frac > 0 && nbytes >= ctx->ascii_min_bytes && nbytes >= 128) split = (uint64_t)((double)nbytes * (1.0 - ctx->ascii_frac)) & ~63ull;
        if (split < nbytes) {
            R(b.d_ascii->reserve(nbytes - split + 64));
            SG_CUDA(cudaMemcpyAsync(b.d_ascii->p, S.blob + base + split, nbytes - split, cudaMemcpyHostToDevice, hst));
            cs.h2d_ascii += nbytes - split;
        }
        const uint64_t used = (split + 15) / 16;
        uint64_t bad;
        {
            const auto t0 = std::chrono::steady_clock::now();
            bad = sg_host_pack_2bit(S.blob + base, split, hp, std::max(1, ctx->host_threads));
            cs.pack_thread_ns += (uint64_t)std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t0).count() *
                                 (uint64_t)std::max(1, ctx->host_threads);
        }
        if (bad != ~0ull) {
            const uint64_t i = (uint64_t)(std::upper_bound(S.off + b.i0, S.off + b.i1 + 1, base + bad) - S.off) - 1;
            return bad_base_error(b.what, b.unit, i, base + bad - S.off[i]);
        }
        if (split < nbytes) {
            SG_CUDA(cudaMemcpyAsync(b.d_packed->p, hp, used * 4, cudaMemcpyHostToDevice, hst));
            cs.h2d_packed += used * 4;
            *b.bad_bias = split;
            R(handover());
            R(sg_dev_pack_2bit(b.d_ascii->as<char>(), nbytes - split, b.d_packed->as<uint32_t>() + used, b.d_bad, st));
            continue;
        }
        memset(hp + used, 0, (words - used) * 4);  // padding words the aligner may read
        SG_CUDA(cudaMemcpyAsync(b.d_packed->p, hp, words * 4, cudaMemcpyHostToDevice, hst));
        cs.h2d_packed += words * 4;
    }
    if (nseg) {
        uint64_t bad = ~0ull;
        int bad_seg = 0;
        R(upload_adaptive(ctx, d, st, ev_copied, ev_dma, seg, nseg, cs, &bad, &bad_seg));
        if (bad != ~0ull) {
            const BlobIn &b = in[seg_of[bad_seg]];
            const Strings &S = *b.S;
            const uint64_t base = S.off[b.i0];
            const uint64_t i = (uint64_t)(std::upper_bound(S.off + b.i0, S.off + b.i1 + 1, base + bad) - S.off) - 1;
            return bad_base_error(b.what, b.unit, i, base + bad - S.off[i]);
        }
        for (int k = 0; k < nseg; k++) *in[seg_of[k]].bad_bias = seg[k].split < seg[k].nbytes ? seg[k].split : 0;
    }
    return SG_OK;
}

// stage A: uploads, ingest, descriptors, alignment kernel, run-count scan; ends with ev_mid
int stage_a(sg_ctx *ctx, Device &d, Slot &s, const Workload &w, uint64_t a0, uint64_t a1, sg_result *res, ShardStats &cs)
{
    const uint64_t n = a1 - a0;
    const bool want_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
    cudaStream_t st = s.stream;
    s.a0 = a0; s.a1 = a1; s.busy = true; s.mid_done = false; s.total_runs = 0;
    cs.sub_batches++;
    trace(d.id, a0, n, "upload begins");
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
    s.bad_bias[0] = s.bad_bias[1] = 0;
    if (!w.mapping) {
        {
            ScopedT t_up(cs.upload);
            if (w.text.off) {
                BlobIn in[2] = {{&w.text, a0, a1, "text", "pair", &s.ascii_t, &s.packed_t, &s.h_stage_t, s.bad.as<uint64_t>(), h_tstart, &s.bad_bias[0]},
                                {&w.query, a0, a1, "query", "pair", &s.ascii_q, &s.packed_q, &s.h_stage_q, s.bad.as<uint64_t>() + 1, h_qstart, &s.bad_bias[1]}};
                R(upload_blobs(ctx, d, st, s.ev_copied, s.ev_dma, in, 2, cs));
                trace(d.id, a0, n, "upload queued");
            } else {
                R(upload_separate(ctx, d, st, w.text, a0, a1, "text", "pair", s.packed_t, s.h_stage_t, h_tstart, cs));
                R(upload_separate(ctx, d, st, w.query, a0, a1, "query", "pair", s.packed_q, s.h_stage_q, h_qstart, cs));
            }
        }
        ScopedT t_desc(cs.desc);
        team_for(d, n, [&](uint64_t k0, uint64_t k1) {
            for (uint64_t k = k0; k < k1; k++) { h_tlen[k] = w.text.size(a0 + k); h_qlen[k] = w.query.size(a0 + k); }
        });
        d_text = s.packed_t.as<uint32_t>();
    } else {
        // reads referenced by this sub-batch: the contiguous index range [r0, r1] (candidates arrive read-major, so the
        // range is tight; each read is uploaded and packed once and shared by its candidates, cf. reference
        // twobit_reads, src/genasm_gpu.cu:784-796)
        uint32_t r0 = w.cand_read[a0], r1 = w.cand_read[a0];
        {
            ScopedT t_desc(cs.desc);
            std::mutex mu;
            team_for(d, n, [&](uint64_t k0, uint64_t k1) {
                if (k0 >= k1) return;
                uint32_t lo = w.cand_read[a0 + k0], hi = lo;
                for (uint64_t c = a0 + k0; c < a0 + k1; c++) { lo = std::min(lo, w.cand_read[c]); hi = std::max(hi, w.cand_read[c]); }
                std::lock_guard<std::mutex> g(mu);
                r0 = std::min(r0, lo); r1 = std::max(r1, hi);
            });
        }
        std::vector<uint64_t> rstart((uint64_t)r1 - r0 + 1);
        {
            ScopedT t_up(cs.upload);
            if (w.query.off) {
                BlobIn in[1] = {{&w.query, r0, (uint64_t)r1 + 1, "content", "read", &s.ascii_q, &s.packed_q, &s.h_stage_q, s.bad.as<uint64_t>() + 1,
                                 rstart.data(), &s.bad_bias[1]}};
                R(upload_blobs(ctx, d, st, s.ev_copied, s.ev_dma, in, 1, cs));
            } else {
                R(upload_separate(ctx, d, st, w.query, r0, (uint64_t)r1 + 1, "content", "read", s.packed_q, s.h_stage_q, rstart.data(), cs));
            }
        }
        ScopedT t_desc(cs.desc);
        team_for(d, n, [&](uint64_t k0, uint64_t k1) {
            for (uint64_t k = k0; k < k1; k++) {
                const uint64_t cstart = w.cand_start[a0 + k];
                const uint32_t r = w.cand_read[a0 + k];
                h_tstart[k] = cstart;
                h_tlen[k] = d.genome_len - cstart;  // the text runs to the end of the genome (src/genasm_cpu.cpp:512-514)
                h_qstart[k] = rstart[r - r0];
                h_qlen[k] = w.query.size(r);
            }
        });
        d_text = d.genome.as<uint32_t>();
    }
    uint64_t slab_bytes = 0;
    // The kernel stores an alignment's runs as whole 32-bit words (SG_FLAG_RUN_WORDS) unless SG_EMIT=bytes: a quarter of the
    // store instructions and L2 sector writes.  Measured with apps/sg_variant_ab on a B200 (profiles/r02_variant_ab.jsonl):
    // 10 kbp pairs 11.87 -> 11.65 ms per 303 104, candidate lists with 7 of 8 loci spurious (~20 runs per window: what the
    // reference's callers produce by handing over every seed hit, src/tests.cu:335-409) 18.04 -> 12.74 ms, 150 bp reads
    // 1.674 -> 1.613 ms per 4 M at 32/17 and unchanged at 64/33.  Slots therefore start and end on 4-byte boundaries.
    const bool run_words = want_cigar && ctx->emit != 0;
    {
        ScopedT t_desc(cs.desc);
        if (!w.mapping && w.query.off) {
            // the prefix sum of the capacities is a difference of query offsets (slab_offset_blob, sg_host_threads.h)
            const uint64_t *qo = w.query.off + a0;
            team_for(d, n + 1, [&](uint64_t k0, uint64_t k1) { for (uint64_t k = k0; k < k1; k++) h_slab[k] = slab_offset_blob(qo[k] - qo[0], k, run_words); });
            slab_bytes = h_slab[n];
        } else {
            for (uint64_t k = 0; k < n; k++) { h_slab[k] = slab_bytes; slab_bytes += slab_capacity(h_qlen[k], run_words); }
            h_slab[n] = slab_bytes;
        }
    }
    SG_CUDA(cudaMemcpyAsync(s.desc.p, s.h_desc.p, (5 * n + 1) * 8, cudaMemcpyHostToDevice, d.h2d));
    cs.h2d_other += (5 * n + 1) * 8;
    // Mixed lengths in one launch: the kernel's queue hands the alignments out longest first (the reference's callers sort
    // their reads by descending length before the call for the same reason, src/tests.cu:377), results stay in input order.
    const uint32_t *d_order = nullptr;
    if (ctx->longest_first && n > 1 && n <= 0xFFFFFFFFull) {
        ScopedT t_desc(cs.desc);
        uint64_t lo = h_qlen[0], hi = h_qlen[0];
        {
            std::mutex mm;
            team_for(d, n, [&](uint64_t k0, uint64_t k1) {
                if (k0 >= k1) return;
                uint64_t l = h_qlen[k0], h = l;
                for (uint64_t k = k0 + 1; k < k1; k++) { l = std::min(l, h_qlen[k]); h = std::max(h, h_qlen[k]); }
                std::lock_guard<std::mutex> g(mm);
                lo = std::min(lo, l); hi = std::max(hi, h);
            });
        }
        if (hi > lo + lo / 4 + 64) {
            R(s.order.reserve(n * 4)); R(s.h_order.reserve(n * 4));
            uint32_t *ho = s.h_order.as<uint32_t>();
            for (uint64_t k = 0; k < n; k++) ho[k] = (uint32_t)k;
            std::stable_sort(ho, ho + n, [&](uint32_t a, uint32_t b) { return h_qlen[a] > h_qlen[b]; });
            SG_CUDA(cudaMemcpyAsync(s.order.p, ho, n * 4, cudaMemcpyHostToDevice, d.h2d));
            cs.h2d_other += n * 4;
            d_order = s.order.as<uint32_t>();
        }
    }
    if (want_cigar) R(s.slab.reserve(slab_bytes + 16));
    // everything this sub-batch needs from the host is queued in the GPU's host-to-device stream: hand over to its own
    SG_CUDA(cudaEventRecord(s.ev_copied, d.h2d));
    SG_CUDA(cudaStreamWaitEvent(st, s.ev_copied, 0));
    SG_CUDA(cudaEventRecord(s.ev_k0, st));
    R(sg_dev_align_ordered(ctx->W, ctx->O, d_text, d_tstart, d_tlen, s.packed_q.as<uint32_t>(), d_qstart, d_qlen, n,
                           w.flags | (run_words ? SG_FLAG_RUN_WORDS : 0u), s.slab.as<uint8_t>(), d_slab,
                           s.counter.as<uint64_t>(), s.edit.as<int64_t>(), s.refc.as<uint64_t>(), s.nruns.as<uint32_t>(), s.status.as<uint8_t>(),
                           nullptr, nullptr, d_order, st));
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
    cs.d2h += 17 * n + 24;
    SG_CUDA(cudaEventRecord(s.ev_mid, st));
    trace(d.id, a0, n, "first stage queued");
    return SG_OK;
}

// stage B: once the run total is known, gather the runs and send everything home; ends with ev_end
int stage_b(Slot &s, const Workload &w, sg_result *res, ShardStats &cs)
{
    if (!s.busy || s.mid_done) return SG_OK;
    const uint64_t n = s.a1 - s.a0;
    const bool want_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
    cudaStream_t st = s.stream;
    trace(-1, s.a0, n, "second stage: waiting for the kernel");
    { ScopedT t(cs.wait); SG_CUDA(cudaEventSynchronize(s.ev_mid)); }
    trace(-1, s.a0, n, "second stage: kernel done");
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
    // an alignment that ran out of slab capacity: stop here, before anything is gathered from the slab
    if (const void *hit = n ? memchr(s.h_status.p, SG_ERR_CIGAR_OVERFLOW, n) : nullptr)
        return fail(SG_ERR_CIGAR_OVERFLOW, "alignment " + std::to_string(s.a0 + (uint64_t)((const uint8_t *)hit - s.h_status.as<uint8_t>())) +
                                               " exceeded its run capacity");
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
        cs.d2h += s.total_runs + n * 8;
    }
    SG_CUDA(cudaEventRecord(s.ev_end, st));
    trace(-1, s.a0, n, "second stage queued, runs", (double)s.total_runs);
    return SG_OK;
}

// stage C: results of the slot's sub-batch into the caller-visible result; frees the slot
int stage_c(Slot &s, const Workload &w, sg_result *res, ShardOut &so)
{
    if (!s.busy) return SG_OK;
    const bool want_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
    R(stage_b(s, w, res, so.stats));
    trace(-1, s.a0, s.a1 - s.a0, "results: waiting for the copies back");
    { ScopedT t(so.stats.wait); SG_CUDA(cudaEventSynchronize(s.ev_end)); }
    s.busy = false;
    ScopedT t_out(so.stats.out);
    float ms = 0;
    SG_CUDA(cudaEventElapsedTime(&ms, s.ev_k0, s.ev_k1));
    so.kernel_ms += ms;
    trace(-1, s.a0, s.a1 - s.a0, "results in place; kernel ms", ms);
    if (want_cigar) {
        so.pieces.push_back(s.piece);
        so.piece_first.push_back(s.a0);
        so.piece_runs.push_back(s.total_runs);
        s.piece = {nullptr, 0};
    }
    return SG_OK;
}

// Sub-batch cuts of the alignment range [c0, c1): a sub-batch grows while it stays under max_batch_bytes and is either
// still small in bytes or still short of min_batch_units alignments.  Both stop conditions are monotone in the end:
// binary search for the first end in (a, c1) that stops (a batch always takes at least one alignment).
std::vector<uint64_t> sub_batch_cuts(uint64_t batch_bytes, uint64_t max_batch_bytes, uint64_t min_batch_units, const uint64_t *woff,
                                     uint64_t per_unit_extra, uint64_t c0, uint64_t c1)
{
    std::vector<uint64_t> cuts{c0};
    while (cuts.back() < c1) {
        const uint64_t a = cuts.back();
        auto stops = [&](uint64_t b) {
            const uint64_t bytes = (woff[b + 1] - woff[a]) + per_unit_extra * (b + 1 - a);
            return bytes > max_batch_bytes || (bytes > batch_bytes && b - a >= min_batch_units);
        };
        uint64_t lo = a + 1, hi = c1;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (stops(mid)) hi = mid; else lo = mid + 1;
        }
        cuts.push_back(lo);
    }
    return cuts;
}

// The end of a call is exposed: after the last upload nothing overlaps the last sub-batch's kernel, compaction and copy
// back.  So the tail of the call is cut finer -- the last full sub-batch is replaced by pieces of 1/2, 1/4, 1/8, 1/8 of
// its weight -- and the exposed remainder shrinks eightfold (measured on 524 288 x 10 kbp pairs: 21 ms of a 110 ms call
// were spent waiting for the last sub-batch).  Small launches do not fill the device, but at that point there is nothing
// else for it to do.
void taper_tail(std::vector<uint64_t> &cuts, const uint64_t *woff, uint64_t per_unit_extra, uint64_t min_piece_bytes)
{
    if (cuts.size() < 2) return;
    size_t k = cuts.size() - 2;   // the last sub-batch [cuts[k], cuts[k+1])
    auto weight = [&](uint64_t a, uint64_t b) { return (woff[b] - woff[a]) + per_unit_extra * (b - a); };
    // a short last sub-batch is joined with the one before it, so that the taper always works on a full one
    if (k > 0 && weight(cuts[k], cuts[k + 1]) * 2 < weight(cuts[k - 1], cuts[k])) { cuts.erase(cuts.begin() + (long)k); k--; }
    const uint64_t a = cuts[k], b = cuts[k + 1], total = weight(a, b);
    if (total < 8 * min_piece_bytes || b - a < 16) return;
    std::vector<uint64_t> inner;
    for (uint64_t num : {4ull, 6ull, 7ull}) {   // boundaries at 1/2, 3/4, 7/8 of the weight
        const uint64_t target = total / 8 * num;
        uint64_t lo = a + 1, hi = b - 1;
        while (lo < hi) {
            const uint64_t mid = lo + (hi - lo) / 2;
            if (weight(a, mid) < target) lo = mid + 1; else hi = mid;
        }
        if ((inner.empty() ? a : inner.back()) < lo && lo < b) inner.push_back(lo);
    }
    cuts.insert(cuts.begin() + (long)k + 1, inner.begin(), inner.end());
}

// One GPU's worker: a three-slot pipeline over the sub-batches it takes from the call's shared queue.  The queue is
// the list of sub-batch cuts of the whole call in order; a GPU that finishes early simply takes more of them (dynamic
// balance across GPUs, like the reference's atomic pair counter does across thread blocks, src/genasm_gpu.cu:602-622).
struct BatchQueue {
    std::vector<uint64_t> cuts;
    std::atomic<size_t> next{0};
    // the largest sub-batch of the call: alignments, uploaded text / query bytes, slab bytes (0 = unknown: grow on demand)
    uint64_t max_n = 0, max_text = 0, max_query = 0, max_slab = 0;
    bool blobs = false, mapping = false;
};

int presize_slot(sg_ctx *ctx, Device &d, Slot &s, const BatchQueue &q)
{
    const uint64_t n = q.max_n;
    if (!n) return SG_OK;
    R(s.desc.reserve((5 * n + 1) * 8)); R(s.h_desc.reserve((5 * n + 1) * 8));
    R(s.counter.reserve(8)); R(s.edit.reserve(n * 8)); R(s.refc.reserve(n * 8));
    R(s.nruns.reserve(n * 4)); R(s.status.reserve(n)); R(s.run_off.reserve((n + 1) * 8));
    R(s.scan_tmp.reserve(sg_scan_tmp_bytes(n))); R(s.bad.reserve(16)); R(s.h_small.reserve(64));
    R(s.h_status.reserve(n));
    if (q.max_slab) R(s.slab.reserve(q.max_slab + 16));
    if (!q.blobs) return SG_OK;
    const bool stage = d.team.size() > 0, ascii = ctx->dma_depth > 0 || d.team.size() == 0 || !ctx->adaptive;
    if (!q.mapping && q.max_text) {
        R(s.packed_t.reserve(sg_packed_words(q.max_text) * 4));
        if (stage) R(s.h_stage_t.reserve(sg_packed_words(q.max_text) * 4 + 64));
        if (ascii) R(s.ascii_t.reserve(q.max_text + 64));
    }
    if (q.max_query) {
        R(s.packed_q.reserve(sg_packed_words(q.max_query) * 4));
        if (stage) R(s.h_stage_q.reserve(sg_packed_words(q.max_query) * 4 + 64));
        if (ascii) R(s.ascii_q.reserve(q.max_query + 64));
    }
    return SG_OK;
}

void run_shard(sg_ctx *ctx, Device &d, const Workload &w, BatchQueue &q, sg_result *res, ShardOut &so)
{
    auto bail = [&](int rc) {
        so.rc = rc;
        so.err = g_last_error;
        if (d.h2d) cudaStreamSynchronize(d.h2d);
        for (Slot &s : d.slots) {  // let the device drain before the buffers are reused
            if (s.stream) cudaStreamSynchronize(s.stream);
            s.busy = false;
        }
        cudaGetLastError();
    };
    if (cudaSetDevice(d.id) != cudaSuccess) { cudaGetLastError(); fail(SG_ERR_CUDA, "cudaSetDevice failed"); bail(SG_ERR_CUDA); return; }
    const size_t nb = q.cuts.size() - 1;
    const int kSlots = d.n_slots;
    // Which sub-batch lands in which slot of which GPU changes from call to call (one queue, several takers), and a slot
    // whose buffers have to grow pays for it with cudaFree / cudaMallocHost in the middle of the pipeline (page-locking
    // the staging of a 2 GB sub-batch: ~0.2 s).  So every slot this call can use is sized for the call's largest
    // sub-batch up front; buffers only ever grow, a second call of the same shape allocates nothing.
    {
        const int used = (int)std::min<size_t>((size_t)kSlots, nb);
        for (int k = 0; k < used; k++) {
            const int rc = presize_slot(ctx, d, d.slots[k], q);
            if (rc) { bail(rc); return; }
        }
    }
    int k = 0;   // sub-batches this GPU has taken
    // The worker never blocks on the device while there is something to upload: a sub-batch's second stage (it needs the
    // run total from the device) is started as soon as a poll finds its kernel finished -- after an upload, or when its
    // slot is needed again -- and only the end of the call waits.  (Blocking here after every upload cost nothing with
    // 2 GB sub-batches, whose kernel is long done when the next upload ends, but it stalled the feeder and the packers
    // behind the kernels of the small sub-batches at the tapered end of a call.)
    auto poll_second_stages = [&](int upto) -> int {
        for (int j = std::max(0, k - kSlots + 1); j <= upto; j++) {   // never the slot being filled (j == k only after stage_a(k))
            Slot &sj = d.slots[j % kSlots];
            if (!sj.busy || sj.mid_done) continue;
            const cudaError_t qe = cudaEventQuery(sj.ev_mid);
            if (qe == cudaErrorNotReady) break;   // second stages start in order: their copies back land in order
            if (qe != cudaSuccess) { cudaGetLastError(); return fail(SG_ERR_CUDA, "cudaEventQuery failed"); }
            const int rc = stage_b(sj, w, res, so.stats);
            if (rc) return rc;
        }
        return SG_OK;
    };
    int poll_rc = SG_OK;
    d.idle_poll = [&]() -> int {   // between chunk copies of sub-batch k: the sub-batches before it
        const int rc = poll_second_stages(k - 1);
        if (rc) poll_rc = rc;
        return rc;
    };
    while (true) {
        const size_t b = q.next.fetch_add(1);
        if (b >= nb) break;
        Slot &s = d.slots[k % kSlots];
        int rc = stage_c(s, w, res, so);                       // frees the slot used by this GPU's batch k - kSlots
        if (!rc) rc = stage_a(ctx, d, s, w, q.cuts[b], q.cuts[b + 1], res, so.stats);
        if (poll_rc) rc = poll_rc;   // a second stage failed inside the feeder's poll (e.g. the device found a bad base in an
                                     // earlier sub-batch): its code and message are the call's
        if (!rc) rc = poll_second_stages(k);
        if (rc) { d.idle_poll = nullptr; bail(rc); return; }
        k++;
    }
    d.idle_poll = nullptr;
    for (int j = std::max(0, k - kSlots); j < k; j++) {       // the end of the call: second stages first, all of them ...
        int rc = stage_b(d.slots[j % kSlots], w, res, so.stats);
        if (rc) { bail(rc); return; }
    }
    for (int j = std::max(0, k - kSlots); j < k; j++) {       // ... then the results
        int rc = stage_c(d.slots[j % kSlots], w, res, so);
        if (rc) { bail(rc); return; }
    }
}

void finalize(sg_result *res, std::vector<ShardOut> &shards)
{
    double kms = 0;
    for (ShardOut &so : shards) kms = std::max(kms, so.kernel_ms);
    res->kernel_ns = (int64_t)(kms * 1e6);
    if (!res->has_cigar) return;
    // every processed sub-batch left one piece; with several GPUs taking sub-batches from one queue the pieces of a
    // shard are ordered but interleaved with the other shards': sort all by first alignment, then rebase the
    // per-piece run offsets to global ones
    struct Piece { PinnedPool::Block blk; uint64_t first, runs; };
    std::vector<Piece> all;
    for (ShardOut &so : shards) {
        for (size_t k = 0; k < so.pieces.size(); k++) all.push_back({so.pieces[k], so.piece_first[k], so.piece_runs[k]});
        so.pieces.clear();
    }
    std::sort(all.begin(), all.end(), [](const Piece &a, const Piece &b) { return a.first < b.first; });
    uint64_t run0 = 0;
    for (const Piece &pc : all) {
        res->piece_first.push_back(pc.first);
        res->piece_run0.push_back(run0);
        res->piece_runs.push_back(pc.runs);
        res->pieces.push_back(pc.blk);
        run0 += pc.runs;
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

// Restores the caller's current CUDA device when a public entry point returns.
struct ScopedDevice {
    int saved = -1;
    ScopedDevice() { if (cudaGetDevice(&saved) != cudaSuccess) { cudaGetLastError(); saved = -1; } }
    ~ScopedDevice() { if (saved >= 0) cudaSetDevice(saved); }
};

int run_all(sg_ctx *ctx, const Workload &w, const uint64_t *woff, uint64_t per_unit_extra, uint64_t n, sg_result **out)
{
    auto t_begin = std::chrono::steady_clock::now();
    g_trace_t0 = t_begin;
    ScopedDevice keep_device;
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
        BatchQueue q;
        q.cuts = sub_batch_cuts(ctx->batch_bytes, ctx->max_batch_bytes, ctx->min_batch_units, woff, per_unit_extra, 0, n);
        // a call with fewer than two sub-batches per GPU would leave GPUs idle or without anything to overlap: cut finer
        if (nd > 1 && q.cuts.size() - 1 < (size_t)nd * 2) {
            const uint64_t total = woff[n] - woff[0] + per_unit_extra * n;
            q.cuts = sub_batch_cuts(std::max<uint64_t>(1, total / ((uint64_t)nd * 2)), ctx->max_batch_bytes, 1, woff, per_unit_extra, 0, n);
        }
        if (ctx->taper) taper_tail(q.cuts, woff, per_unit_extra, 4ull << 20);
        q.mapping = w.mapping;
        q.blobs = w.query.off != nullptr && (w.mapping || w.text.off != nullptr);
        const bool want_cigar = !(w.flags & SG_FLAG_DISTANCE_ONLY);
        for (size_t b = 0; b + 1 < q.cuts.size(); b++) {
            const uint64_t a0 = q.cuts[b], a1 = q.cuts[b + 1];
            q.max_n = std::max(q.max_n, a1 - a0);
            if (!q.blobs || w.mapping) continue;   // mapping: the reads of a sub-batch are known only after a scan of its candidates
            q.max_text = std::max(q.max_text, w.text.off[a1] - w.text.off[a0]);
            q.max_query = std::max(q.max_query, w.query.off[a1] - w.query.off[a0]);
            if (want_cigar) q.max_slab = std::max<uint64_t>(q.max_slab, 2 * (w.query.off[a1] - w.query.off[a0]) + (ctx->emit ? 12 : 8) * (a1 - a0) + 4);
        }
        if (nd == 1) {
            ScopedAffinity bound(ctx->devs[0].worker_cpus);   // the caller's thread is this GPU's worker for the call
            run_shard(ctx, ctx->devs[0], w, q, res.get(), shards[0]);
        } else {
            std::vector<std::thread> th;
            for (int k = 0; k < nd; k++)
                th.emplace_back([&, k]() {
                    bind_this_thread(ctx->devs[k].worker_cpus);
                    run_shard(ctx, ctx->devs[k], w, q, res.get(), shards[k]);
                });
            for (auto &t : th) t.join();
        }
        for (ShardOut &so : shards)
            if (so.rc) return fail(so.rc, so.err);
    }
    finalize(res.get(), shards);
    g_pool.call_done();
    res->total_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
    sg_call_stats &S = res->stats;
    memset(&S, 0, sizeof S);
    S.total_ns = res->total_ns;
    S.kernel_ns = res->kernel_ns;
    S.n_devices = (uint32_t)nd;
    S.host_threads_per_device = (uint32_t)ctx->host_threads;
    for (ShardOut &so : shards) {
        const ShardStats &c = so.stats;
        S.upload_ns = std::max<int64_t>(S.upload_ns, (int64_t)(c.upload * 1e9));
        S.wait_ns = std::max<int64_t>(S.wait_ns, (int64_t)(c.wait * 1e9));
        S.host_other_ns = std::max<int64_t>(S.host_other_ns, (int64_t)((c.desc + c.out) * 1e9));
        S.pack_thread_ns += (int64_t)c.pack_thread_ns.load();
        S.h2d_ascii_bytes += c.h2d_ascii;
        S.h2d_packed_bytes += c.h2d_packed.load();
        S.h2d_other_bytes += c.h2d_other;
        S.d2h_bytes += c.d2h;
        S.sub_batches += c.sub_batches;
        S.packers_in_use = std::max(S.packers_in_use, c.packers_in_use);
    }
    if (g_debug)
        fprintf(stderr, "[sg] call %.1f ms on %d GPU(s): ingest %.1f (packer threads busy %.1f thread-ms), waits %.1f, other host %.1f; "
                        "H2D %.1f MB ASCII + %.1f MB packed + %.1f MB descriptors, D2H %.1f MB, %u sub-batches\n",
                S.total_ns / 1e6, nd, S.upload_ns / 1e6, S.pack_thread_ns / 1e6, S.wait_ns / 1e6, S.host_other_ns / 1e6, S.h2d_ascii_bytes / 1e6,
                S.h2d_packed_bytes / 1e6, S.h2d_other_bytes / 1e6, S.d2h_bytes / 1e6, S.sub_batches);
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

static std::atomic<int> g_live_contexts{0};

int sg_ctx_create_wo(sg_ctx **out, const int *device_ids, int n_devices, int W, int O)
{
    if (!out) return fail(SG_ERR_BAD_ARG, "sg_ctx_create: null out");
    ScopedDevice keep_device;
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
        // Host threads per GPU for packing: the CPUs this process may use (its affinity mask, or SG_CPUS=<list>) divided
        // among the context's GPUs; SG_HOST_THREADS=<n> sets the count per GPU directly (0: no packer threads, every blob
        // crosses PCIe as ASCII).  One of a GPU's CPUs is left to its worker thread (the feeder of the adaptive ingest).
        const std::vector<int> allowed = process_cpus();
        ctx->host_threads = std::max(1, (int)allowed.size() / n_devices - (allowed.size() / n_devices >= 4 ? 1 : 0));
        if (const char *v = std::getenv("SG_HOST_THREADS")) ctx->host_threads = std::max(0, std::atoi(v));
        ctx->host_pack = ctx->host_threads >= 10;   // ~7-10 GB/s of ASCII per thread against ~47 GB/s of PCIe per GPU
        if (const char *v = std::getenv("SG_HOST_PACK")) ctx->host_pack = std::atoi(v) != 0;
        if (const char *v = std::getenv("SG_ASCII_MIN_BYTES")) ctx->ascii_min_bytes = (uint64_t)std::max(0ll, std::atoll(v));
        if (const char *v = std::getenv("SG_ASCII_PCT")) ctx->ascii_frac = std::min(100, std::max(0, std::atoi(v))) / 100.0;
        // SG_INGEST=adaptive (default) | fixed: the pre-adaptive policies (host packing with a fixed ASCII share when the
        // GPU has >= 10 host threads, else ASCII upload + device packing); SG_HOST_PACK / SG_ASCII_PCT imply fixed
        ctx->adaptive = !std::getenv("SG_HOST_PACK") && !std::getenv("SG_ASCII_PCT");
        if (const char *v = std::getenv("SG_INGEST")) ctx->adaptive = std::string(v) == "adaptive";
        if (const char *v = std::getenv("SG_CHUNK_KB")) ctx->chunk_bytes = std::max<uint64_t>(256, ((uint64_t)std::max(1ll, std::atoll(v)) << 10) & ~255ull);
        // SG_DMA_DEPTH=<n>: ASCII chunk copies the feeder keeps in flight (0: none -- the host packs everything)
        if (const char *v = std::getenv("SG_DMA_DEPTH")) ctx->dma_depth = std::min(kMaxDmaDepth, std::max(0, std::atoi(v)));
        if (ctx->host_threads == 0) ctx->dma_depth = std::max(1, ctx->dma_depth);
    }
    // which CPUs serve which GPU (sg_host_threads.h); SG_AFFINITY=0 leaves every thread where the scheduler puts it
    std::vector<std::vector<int>> cpu_sets(n_devices);
    if (!(std::getenv("SG_AFFINITY") && std::atoi(std::getenv("SG_AFFINITY")) == 0)) {
        std::vector<std::vector<int>> local(n_devices);
        for (int k = 0; k < n_devices; k++) {
            char bus[32] = {0};
            const int id = device_ids ? device_ids[k] : k;
            if (id >= 0 && id < avail && cudaDeviceGetPCIBusId(bus, sizeof bus, id) == cudaSuccess) local[k] = pci_local_cpus(bus);
            cudaGetLastError();
        }
        cpu_sets = assign_cpus(process_cpus(), local);
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
        d.cpus = cpu_sets[k];
        d.worker_cpus = d.packer_cpus = d.cpus;
        if (d.cpus.size() >= 2 && ctx->host_threads > 0) {
            d.worker_cpus.assign(1, d.cpus.back());
            d.packer_cpus.assign(d.cpus.begin(), d.cpus.end() - 1);
        }
        SG_CUDA(cudaStreamCreateWithFlags(&d.h2d, cudaStreamNonBlocking));
        if (const char *v = std::getenv("SG_SLOTS")) d.n_slots = std::min(kMaxSlots, std::max(2, std::atoi(v)));
        for (int q = 0; q < d.n_slots; q++) R(d.slots[q].create());
        if (ctx->host_threads > 0) {
            const int dev_id = d.id;
            d.team.start(ctx->host_threads, d.packer_cpus, [dev_id](int) { cudaSetDevice(dev_id); });
        }
        d.tuner.init(ctx->host_threads);
        int wps = 0, sms = 0;
        R(sg_dev_align_geometry_wo(W, O, &wps, nullptr, &sms));
        // one alignment per resident lane fills the device (104 192 lanes on a B200 at W=64)
        ctx->min_batch_units = std::max<uint64_t>(ctx->min_batch_units, 32ull * (uint64_t)wps * (uint64_t)sms);
    }
    if (const char *v = std::getenv("SG_LONGEST_FIRST")) ctx->longest_first = std::atoi(v) != 0;
    if (const char *v = std::getenv("SG_EMIT")) ctx->emit = std::string(v) == "bytes" ? 0 : 1;
    if (const char *v = std::getenv("SG_TAPER")) ctx->taper = std::atoi(v) != 0;
    if (const char *v = std::getenv("SG_MIN_BATCH_UNITS")) ctx->min_batch_units = (uint64_t)std::max(1ll, std::atoll(v));
    if (g_debug)
        for (const Device &d : ctx->devs) {
            std::string l;
            for (int c : d.cpus) l += (l.empty() ? "" : ",") + std::to_string(c);
            fprintf(stderr, "[sg] GPU %d: %d packer threads on CPUs {%s}, the last one for the worker\n", d.id, ctx->host_threads, l.c_str());
        }
    g_live_contexts++;
    ctx->counted = true;
    *out = ctx.release();
    return SG_OK;
}

void sg_ctx_destroy(sg_ctx *ctx)
{
    if (!ctx) return;
    {
        ScopedDevice keep_device;
        for (Device &d : ctx->devs) {
            d.team.stop();
            cudaSetDevice(d.id);
            for (Slot &s : d.slots) s.destroy();
            if (d.h2d) cudaStreamDestroy(d.h2d);
            d.h2d = nullptr;
            d.genome.release();
            d.piece_bad.release();
        }
    }
    const bool counted = ctx->counted;
    delete ctx;
    // the page-locked blocks the pool kept for reuse go back to the system with the last context
    if (counted && --g_live_contexts == 0) g_pool.trim(0);
}

void sg_trim_host_cache(void) { g_pool.trim(0); }

uint64_t sg_plan_sub_batches(const uint64_t *weight_off, uint64_t n, uint64_t per_unit_extra, uint64_t batch_bytes, uint64_t max_batch_bytes,
                             uint64_t min_batch_units, int taper, uint64_t *cuts, uint64_t cuts_cap)
{
    if (!weight_off || !n) return 0;
    std::vector<uint64_t> c = sub_batch_cuts(batch_bytes, max_batch_bytes, min_batch_units, weight_off, per_unit_extra, 0, n);
    if (taper) taper_tail(c, weight_off, per_unit_extra, 4ull << 20);
    for (uint64_t k = 0; k < c.size() && cuts && k < cuts_cap; k++) cuts[k] = c[k];
    return c.size();
}

int sg_result_stats(const sg_result *r, sg_call_stats *out)
{
    if (!r || !out) return fail(SG_ERR_BAD_ARG, "sg_result_stats: null argument");
    *out = r->stats;
    return SG_OK;
}

int sg_ctx_num_devices(const sg_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

// n+1 offsets must not decrease, and a blob that holds bytes must exist: otherwise sizes wrap to ~2^64 further down
static int check_blob(const char *name, const char *blob, const uint64_t *off, uint64_t n, int threads)
{
    std::atomic<uint64_t> first_bad{~0ull};
    parallel_for(n, threads, [&](uint64_t k0, uint64_t k1) {
        for (uint64_t k = k0; k < k1; k++)
            if (off[k + 1] < off[k]) {
                uint64_t cur = first_bad.load();
                while (k < cur && !first_bad.compare_exchange_weak(cur, k)) {}
                return;
            }
    });
    if (first_bad.load() != ~0ull)
        return fail(SG_ERR_BAD_ARG, std::string(name) + " offsets decrease at entry " + std::to_string(first_bad.load() + 1));
    if (!blob && off[n] > off[0]) return fail(SG_ERR_BAD_ARG, std::string(name) + " blob is NULL but its offsets span " + std::to_string(off[n] - off[0]) + " bytes");
    return SG_OK;
}

static int check_separate(const char *name, const char *const *ptr, const uint64_t *len, uint64_t n)
{
    for (uint64_t k = 0; k < n; k++)
        if (!ptr[k] && len[k]) return fail(SG_ERR_BAD_ARG, std::string(name) + " " + std::to_string(k) + " is NULL with a non-zero length");
    return SG_OK;
}

static int align_pairs_common(sg_ctx *ctx, const Strings &text, const Strings &query, uint64_t n_pairs, uint32_t flags, sg_result **out)
{
    std::lock_guard<std::mutex> lock(ctx->mu);
    if (text.off) {
        R(check_blob("text", text.blob, text.off, n_pairs, ctx->host_threads));
        R(check_blob("query", query.blob, query.off, n_pairs, ctx->host_threads));
    } else {
        R(check_separate("text", text.ptr, text.len, n_pairs));
        R(check_separate("query", query.ptr, query.len, n_pairs));
    }
    Workload w;
    w.text = text; w.query = query; w.flags = flags & ~SG_FLAG_RUN_WORDS;   // the slab layout is this layer's business
    // sub-batches are cut by uploaded bytes (text + query), shards by the same weight
    std::unique_ptr<uint64_t[]> woff(new uint64_t[n_pairs + 1]);
    if (text.off && query.off) {   // a prefix sum of sizes is a difference of offsets: no serial pass over the pairs
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
        // upload in 256 Mbase pieces (a multiple of 16 bases, so every piece packs to whole words) through two staging
        // buffers: copies go to a copy stream, packing kernels to the slot's stream, chained by events -- the copy of
        // piece k+1 overlaps the packing of piece k, a buffer is reused only after the kernel that read it.  The
        // offending-base word is checked once, after the last piece (the kernel keeps the smallest position).
        const uint64_t piece = 256ull << 20;
        DevBuf *stage[2] = {&s.ascii_t, &s.ascii_q};
        R(stage[0]->reserve(std::min<uint64_t>(piece, genome_len) + 64));
        if (genome_len > piece) R(stage[1]->reserve(std::min<uint64_t>(piece, genome_len - piece) + 64));
        cudaStream_t copy_st = d.h2d;
        cudaEvent_t copied[2] = {s.ev_dma[0], s.ev_dma[1]}, packed[2] = {s.ev_dma[2], s.ev_dma[3]};
        R(d.piece_bad.reserve(16 * (genome_len / piece + 1)));
        SG_CUDA(cudaMemsetAsync(d.piece_bad.p, 0xFF, 16 * (genome_len / piece + 1), s.stream));
        SG_CUDA(cudaEventRecord(packed[0], s.stream));
        SG_CUDA(cudaEventRecord(packed[1], s.stream));
        uint64_t pos = 0, first_bad = ~0ull, npieces = 0;
        int k2 = 0;
        while (pos < genome_len) {
            const uint64_t len = std::min<uint64_t>(piece, genome_len - pos);
            SG_CUDA(cudaStreamWaitEvent(copy_st, packed[k2], 0));
            SG_CUDA(cudaMemcpyAsync(stage[k2]->p, genome_ascii + pos, len, cudaMemcpyHostToDevice, copy_st));
            SG_CUDA(cudaEventRecord(copied[k2], copy_st));
            SG_CUDA(cudaStreamWaitEvent(s.stream, copied[k2], 0));
            R(sg_dev_pack_2bit(stage[k2]->as<char>(), len, d.genome.as<uint32_t>() + pos / 16, d.piece_bad.as<uint64_t>() + 2 * npieces, s.stream));
            SG_CUDA(cudaEventRecord(packed[k2], s.stream));
            pos += len;
            npieces++;
            k2 ^= 1;
        }
        std::vector<uint64_t> hbad(2 * std::max<uint64_t>(npieces, 1), ~0ull);
        if (npieces) SG_CUDA(cudaMemcpyAsync(hbad.data(), d.piece_bad.p, 16 * npieces, cudaMemcpyDeviceToHost, s.stream));
        SG_CUDA(cudaStreamSynchronize(s.stream));
        for (uint64_t q = 0; q < npieces && first_bad == ~0ull; q++)
            if (hbad[2 * q] != ~0ull) first_bad = q * piece + hbad[2 * q];
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
    if (reads.off) R(check_blob("read", reads.blob, reads.off, n_reads, ctx->host_threads));
    else R(check_separate("read", reads.ptr, reads.len, n_reads));
    // weight of a candidate = its read's length: every thread validates and prefix-sums its own range, the ranges are
    // then rebased by the totals of the ranges before them
    std::vector<uint64_t> woff(n_cand + 1, 0);
    {
        std::atomic<uint64_t> bad_read{~0ull}, bad_start{~0ull};
        auto note = [](std::atomic<uint64_t> &a, uint64_t c) { uint64_t cur = a.load(); while (c < cur && !a.compare_exchange_weak(cur, c)) {} };
        std::mutex mu;
        std::vector<std::pair<uint64_t, uint64_t>> ranges;   // (first candidate of a range, its total weight)
        parallel_for(n_cand, ctx->host_threads, [&](uint64_t c0, uint64_t c1) {
            uint64_t acc = 0;
            for (uint64_t c = c0; c < c1; c++) {
                if (cand_read[c] >= n_reads) { note(bad_read, c); continue; }
                if (cand_start[c] > genome_len) note(bad_start, c);
                acc += reads.size(cand_read[c]);
                woff[c + 1] = acc;
            }
            std::lock_guard<std::mutex> g(mu);
            ranges.emplace_back(c0, acc);
        });
        if (bad_read.load() != ~0ull) return fail(SG_ERR_BAD_ARG, "candidate " + std::to_string(bad_read.load()) + ": read index out of range");
        if (bad_start.load() != ~0ull) return fail(SG_ERR_BAD_ARG, "candidate " + std::to_string(bad_start.load()) + ": start beyond the reference");
        std::sort(ranges.begin(), ranges.end());
        uint64_t base = 0;
        for (size_t r = 0; r < ranges.size(); r++) {
            const uint64_t c0 = ranges[r].first, c1 = r + 1 < ranges.size() ? ranges[r + 1].first : n_cand;
            if (base) parallel_for(c1 - c0, ctx->host_threads, [&](uint64_t k0, uint64_t k1) { for (uint64_t k = k0; k < k1; k++) woff[c0 + k + 1] += base; });
            base += ranges[r].second;
        }
    }
    Workload w;
    w.mapping = true;
    w.query = reads; w.cand_start = cand_start; w.cand_read = cand_read; w.flags = flags & ~SG_FLAG_RUN_WORDS;
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

// "%d%c" per packed run (reference src/genasm_gpu.cu:881-888): sg_host_render.cpp, 64 runs per step with AVX-512 VBMI2
extern "C" uint64_t sg_host_runs_text_len(const uint8_t *runs, uint64_t cnt);
extern "C" char *sg_host_runs_render(const uint8_t *runs, uint64_t cnt, char *out);
extern "C" char *sg_host_runs_render_stream(const uint8_t *runs, uint64_t cnt, char *out, char *scratch);
extern "C" void sg_host_stream_fence(void);
static inline uint64_t runs_text_len(const uint8_t *p, uint64_t cnt) { return cnt ? sg_host_runs_text_len(p, cnt) : 0; }
static inline char *runs_render(const uint8_t *p, uint64_t cnt, char *o) { return cnt ? sg_host_runs_render(p, cnt, o) : o; }

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
    // a blob of a few megabytes stays in the caches and is rendered in place; a large one is written around them
    const bool stream = total >= (64ull << 20);
    parallel([&](uint64_t a0, uint64_t a1) {
        std::vector<char> scratch;
        for (uint64_t a = a0; a < a1; a++) {
            uint64_t cnt;
            const uint8_t *p = runs_of(r, a, &cnt);
            if (r->wide_runs) wide_render(p, cnt, blob + text_off[a]);
            else if (stream && cnt) {
                if (scratch.size() < 3 * cnt + 192) scratch.resize(3 * cnt + 4096);
                sg_host_runs_render_stream(p, cnt, blob + text_off[a], scratch.data());
            } else runs_render(p, cnt, blob + text_off[a]);
        }
        if (stream) sg_host_stream_fence();
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
