// sg_host_api.cu -- layer 1 of include/scrooge_b200.h: host buffers in, host results out.
//
// Replaces the host side of the reference GPU library (src/genasm_gpu.cu:692-1065): cudaMallocManaged
// blobs, per-string descriptor loops and a linked-list walk per alignment become
//   * one contiguous ASCII upload per batch, packed to 2 bit/base on the device,
//   * descriptors derived on the device from the offset arrays,
//   * a run slab with per-alignment capacity 2*|query|+8 (reference: 2*|query| entries,
//     src/genasm_gpu.cu:995-1001), compacted on the device and downloaded in one transfer,
//   * a host-side scatter over the context's GPUs: alignments are independent
//     (src/genasm_cpu.cpp:451-455), so each GPU gets a contiguous share balanced by query bases and there
//     is no inter-GPU exchange of any kind.  In mapping mode every GPU holds its own packed reference.
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <string>
#include <vector>
#include <thread>
#include <chrono>
#include <algorithm>
#include <memory>
#include <cuda_runtime.h>

#include "../../include/scrooge_b200.h"
#include "sg_internal.h"

namespace sg {

// ---- tiny RAII device / pinned buffers that only grow ----------------------------------------------
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
            e = cudaMalloc(&p, bytes);
            want = bytes;
            if (e != cudaSuccess) { cudaGetLastError(); p = nullptr; return fail(SG_ERR_OOM, "cudaMalloc failed for " + std::to_string(bytes) + " bytes"); }
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
        if (cudaMallocHost(&p, want) != cudaSuccess) { cudaGetLastError(); p = nullptr; return fail(SG_ERR_OOM, "cudaMallocHost failed"); }
        cap = want;
        return SG_OK;
    }
    void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return (T *)p; }
};

// descriptors from offset arrays: start[a] = off[a] - off[0] + base, len[a] = off[a+1]-off[a], slab_off
__global__ void __launch_bounds__(256) pair_descriptors_kernel(const uint64_t *__restrict__ toff, const uint64_t *__restrict__ qoff,
                                                                uint64_t n, uint64_t *__restrict__ tstart, uint64_t *__restrict__ tlen,
                                                                uint64_t *__restrict__ qstart, uint64_t *__restrict__ qlen,
                                                                uint64_t *__restrict__ slab_off)
{
    const uint64_t a = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (a > n) return;
    const uint64_t q0 = qoff[0];
    slab_off[a] = 2ull * (qoff[a] - q0) + 8ull * a;
    if (a == n) return;
    tstart[a] = toff[a] - toff[0];
    tlen[a] = toff[a + 1] - toff[a];
    qstart[a] = qoff[a] - q0;
    qlen[a] = qoff[a + 1] - qoff[a];
}

// mapping mode: candidate c -> text = genome suffix at cand_start[c], query = read cand_read[c]
__global__ void __launch_bounds__(256) cand_descriptors_kernel(const uint64_t *__restrict__ cand_start, const uint32_t *__restrict__ cand_read,
                                                                const uint64_t *__restrict__ roff, uint32_t read_base, uint64_t genome_len, uint64_t n,
                                                                uint64_t *__restrict__ tstart, uint64_t *__restrict__ tlen,
                                                                uint64_t *__restrict__ qstart, uint64_t *__restrict__ qlen,
                                                                uint32_t *__restrict__ qlen32)
{
    const uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    const uint64_t s = cand_start[c];
    const uint32_t r = cand_read[c] - read_base;
    tstart[c] = s;
    tlen[c] = genome_len - s;  // the text runs to the end of the genome (src/genasm_cpu.cpp:512-514)
    qstart[c] = roff[r] - roff[0];
    const uint64_t ql = roff[r + 1] - roff[r];
    qlen[c] = ql;
    qlen32[c] = (uint32_t)(2ull * ql + 8ull);  // slab capacity; scanned into slab offsets
}

struct Device {
    int id = 0;
    cudaStream_t stream = nullptr;
    cudaEvent_t ev0 = nullptr, ev1 = nullptr;
    DevBuf ascii_t, ascii_q, packed_t, packed_q, toff, qoff, tstart, tlen, qstart, qlen, slab_off, slab, counter, edit,
        refc, nruns, status, run_off, scan_tmp, runs, bad, cstart, cread, cap32;
    DevBuf genome;  // packed reference, resident across calls
    uint64_t genome_len = 0;
    bool has_genome = false;
    PinBuf h_small;
    void release_all()
    {
        for (DevBuf *b : {&ascii_t, &ascii_q, &packed_t, &packed_q, &toff, &qoff, &tstart, &tlen, &qstart, &qlen, &slab_off,
                          &slab, &counter, &edit, &refc, &nruns, &status, &run_off, &scan_tmp, &runs, &bad, &cstart, &cread,
                          &cap32, &genome})
            b->release();
        h_small.release();
    }
};

}  // namespace sg

using namespace sg;

struct sg_ctx {
    int W = 64;
    std::vector<Device> devs;
    uint64_t max_batch_query_bases = 1ull << 31;  // bounds the per-batch slab (2 B per query base)
};

struct sg_result {
    uint64_t n = 0;
    bool has_cigar = false;
    std::vector<int64_t> edit;
    std::vector<uint64_t> refc;
    std::vector<uint64_t> run_off;  // n+1
    // runs of the alignments, one contiguous piece per processed batch; piece_of[a] gives the piece
    std::vector<std::vector<uint8_t>> pieces;
    std::vector<uint64_t> piece_first;   // first alignment of each piece
    std::vector<uint64_t> piece_run0;    // global run offset of each piece's first run
    std::vector<uint8_t> flat;           // lazily flattened view for sg_result_runs
    int64_t kernel_ns = 0, total_ns = 0;
};

namespace {

struct ShardOut {
    int rc = SG_OK;
    std::string err;
    double kernel_ms = 0;
    std::vector<std::vector<uint8_t>> pieces;
    std::vector<uint64_t> piece_first;
};

// Runs one batch [a0, a1) of the unstructured interface on device d.
int run_pairs_batch(sg_ctx *ctx, Device &d, const char *tb, const uint64_t *toff, const char *qb, const uint64_t *qoff,
                    uint64_t a0, uint64_t a1, uint32_t flags, sg_result *res, ShardOut &so)
{
    const uint64_t n = a1 - a0;
    const uint64_t tbytes = toff[a1] - toff[a0], qbytes = qoff[a1] - qoff[a0];
    const bool want_cigar = !(flags & SG_FLAG_DISTANCE_ONLY);
    cudaStream_t st = d.stream;
    int rc;
#define R(x) do { rc = (x); if (rc) return rc; } while (0)
    R(d.ascii_t.reserve(tbytes + 64)); R(d.ascii_q.reserve(qbytes + 64));
    R(d.packed_t.reserve(sg_packed_words(tbytes) * 4)); R(d.packed_q.reserve(sg_packed_words(qbytes) * 4));
    R(d.toff.reserve((n + 1) * 8)); R(d.qoff.reserve((n + 1) * 8));
    R(d.tstart.reserve(n * 8)); R(d.tlen.reserve(n * 8)); R(d.qstart.reserve(n * 8)); R(d.qlen.reserve(n * 8));
    R(d.slab_off.reserve((n + 1) * 8));
    const uint64_t slab_bytes = 2ull * qbytes + 8ull * n;
    if (want_cigar) R(d.slab.reserve(slab_bytes + 16));
    R(d.counter.reserve(8)); R(d.edit.reserve(n * 8)); R(d.refc.reserve(n * 8)); R(d.nruns.reserve(n * 4));
    R(d.status.reserve(n)); R(d.run_off.reserve((n + 1) * 8)); R(d.scan_tmp.reserve(sg_scan_tmp_bytes(n)));
    R(d.bad.reserve(16));
    R(d.h_small.reserve(64));

    SG_CUDA(cudaMemcpyAsync(d.ascii_t.p, tb + toff[a0], tbytes, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d.ascii_q.p, qb + qoff[a0], qbytes, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d.toff.p, toff + a0, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d.qoff.p, qoff + a0, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemsetAsync(d.bad.p, 0xFF, 16, st));
    R(sg_dev_pack_2bit(d.ascii_t.as<char>(), tbytes, d.packed_t.as<uint32_t>(), d.bad.as<uint64_t>(), st));
    R(sg_dev_pack_2bit(d.ascii_q.as<char>(), qbytes, d.packed_q.as<uint32_t>(), d.bad.as<uint64_t>() + 1, st));
    pair_descriptors_kernel<<<(unsigned)((n + 1 + 255) / 256), 256, 0, st>>>(
        d.toff.as<uint64_t>(), d.qoff.as<uint64_t>(), n, d.tstart.as<uint64_t>(), d.tlen.as<uint64_t>(),
        d.qstart.as<uint64_t>(), d.qlen.as<uint64_t>(), d.slab_off.as<uint64_t>());
    SG_CUDA(cudaGetLastError());
    SG_CUDA(cudaEventRecord(d.ev0, st));
    R(sg_dev_align(ctx->W, d.packed_t.as<uint32_t>(), d.tstart.as<uint64_t>(), d.tlen.as<uint64_t>(),
                   d.packed_q.as<uint32_t>(), d.qstart.as<uint64_t>(), d.qlen.as<uint64_t>(), n, flags,
                   d.slab.as<uint8_t>(), d.slab_off.as<uint64_t>(), d.counter.as<uint64_t>(), d.edit.as<int64_t>(),
                   d.refc.as<uint64_t>(), d.nruns.as<uint32_t>(), d.status.as<uint8_t>(), nullptr, st));
    SG_CUDA(cudaEventRecord(d.ev1, st));
    uint64_t *h = d.h_small.as<uint64_t>();
    SG_CUDA(cudaMemcpyAsync(h, d.bad.p, 16, cudaMemcpyDeviceToHost, st));
    if (want_cigar) {
        R(sg_dev_scan_runs(d.nruns.as<uint32_t>(), n, d.run_off.as<uint64_t>(), d.scan_tmp.p, st));
        SG_CUDA(cudaMemcpyAsync(h + 2, d.run_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    }
    SG_CUDA(cudaMemcpyAsync(res->edit.data() + a0, d.edit.p, n * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(res->refc.data() + a0, d.refc.p, n * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (h[0] != ~0ull || h[1] != ~0ull) {
        // name the pair like the reference's assert would have stopped on it (src/genasm_gpu.cu:636)
        const bool in_text = h[0] != ~0ull;
        const uint64_t pos = (in_text ? h[0] + toff[a0] : h[1] + qoff[a0]);
        const uint64_t *off = in_text ? toff : qoff;
        uint64_t p = (uint64_t)(std::upper_bound(off + a0, off + a1 + 1, pos) - off) - 1;
        return fail(SG_ERR_BAD_BASE, std::string("non-ACGT character in ") + (in_text ? "text" : "query") + " of pair " +
                                         std::to_string(p) + " at position " + std::to_string(pos - off[p]));
    }
    float ms = 0;
    SG_CUDA(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
    so.kernel_ms += ms;
    if (want_cigar) {
        const uint64_t total_runs = h[2];
        R(d.runs.reserve(total_runs + 16));
        R(sg_dev_gather_runs(d.slab.as<uint8_t>(), d.slab_off.as<uint64_t>(), d.nruns.as<uint32_t>(), d.run_off.as<uint64_t>(),
                             n, d.runs.as<uint8_t>(), st));
        std::vector<uint8_t> piece(total_runs);
        std::vector<uint8_t> status(n);
        SG_CUDA(cudaMemcpyAsync(piece.data(), d.runs.p, total_runs, cudaMemcpyDeviceToHost, st));
        // per-batch run offsets land in the result's global array; rebased by the caller after all shards finish
        SG_CUDA(cudaMemcpyAsync(res->run_off.data() + a0, d.run_off.p, n * 8, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(status.data(), d.status.p, n, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
        for (uint64_t k = 0; k < n; k++)
            if (status[k]) return fail(SG_ERR_CIGAR_OVERFLOW, "alignment " + std::to_string(a0 + k) + " exceeded its run capacity");
        so.pieces.push_back(std::move(piece));
        so.piece_first.push_back(a0);
    }
#undef R
    return SG_OK;
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

void finalize_runs(sg_result *res, std::vector<ShardOut> &shards)
{
    // pieces are in alignment order once shards are concatenated; rebase per-piece offsets to global ones
    uint64_t run0 = 0;
    for (ShardOut &so : shards) {
        for (size_t k = 0; k < so.pieces.size(); k++) {
            res->piece_first.push_back(so.piece_first[k]);
            res->piece_run0.push_back(run0);
            run0 += so.pieces[k].size();
            res->pieces.push_back(std::move(so.pieces[k]));
        }
    }
    for (size_t k = 0; k < res->pieces.size(); k++) {
        const uint64_t a0 = res->piece_first[k];
        const uint64_t a1 = k + 1 < res->pieces.size() ? res->piece_first[k + 1] : res->n;
        for (uint64_t a = a0; a < a1; a++) res->run_off[a] += res->piece_run0[k];
    }
    res->run_off[res->n] = run0;
}

const uint8_t *runs_of(const sg_result *r, uint64_t idx, uint64_t *count)
{
    *count = r->run_off[idx + 1] - r->run_off[idx];
    size_t k = (size_t)(std::upper_bound(r->piece_first.begin(), r->piece_first.end(), idx) - r->piece_first.begin()) - 1;
    return r->pieces[k].data() + (r->run_off[idx] - r->piece_run0[k]);
}

}  // namespace

extern "C" {

int sg_ctx_create(sg_ctx **out, const int *device_ids, int n_devices, int W)
{
    if (!out) return fail(SG_ERR_BAD_ARG, "sg_ctx_create: null out");
    if (W != 64 && W != 32) return fail(SG_ERR_BAD_ARG, "W must be 64 (O=33) or 32 (O=17)");
    const int avail = sg_device_count();
    if (avail == 0) return fail(SG_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (n_devices <= 0) n_devices = avail;
    std::unique_ptr<sg_ctx> ctx(new sg_ctx);
    ctx->W = W;
    ctx->devs.resize(n_devices);
    for (int k = 0; k < n_devices; k++) {
        Device &d = ctx->devs[k];
        d.id = device_ids ? device_ids[k] : k;
        if (d.id < 0 || d.id >= avail) return fail(SG_ERR_BAD_ARG, "device id out of range");
        SG_CUDA(cudaSetDevice(d.id));
        SG_CUDA(cudaStreamCreateWithFlags(&d.stream, cudaStreamNonBlocking));
        SG_CUDA(cudaEventCreate(&d.ev0));
        SG_CUDA(cudaEventCreate(&d.ev1));
        int wps = 0;
        int rc = sg_dev_align_geometry(W, &wps, nullptr, nullptr);
        if (rc) return rc;
    }
    *out = ctx.release();
    return SG_OK;
}

void sg_ctx_destroy(sg_ctx *ctx)
{
    if (!ctx) return;
    for (Device &d : ctx->devs) {
        cudaSetDevice(d.id);
        d.release_all();
        if (d.ev0) cudaEventDestroy(d.ev0);
        if (d.ev1) cudaEventDestroy(d.ev1);
        if (d.stream) cudaStreamDestroy(d.stream);
    }
    delete ctx;
}

int sg_ctx_num_devices(const sg_ctx *ctx) { return ctx ? (int)ctx->devs.size() : 0; }

int sg_align_pairs(sg_ctx *ctx, const char *text_blob, const uint64_t *text_off, const char *query_blob,
                   const uint64_t *query_off, uint64_t n_pairs, uint32_t flags, sg_result **out)
{
    if (!ctx || !out || !text_off || !query_off) return fail(SG_ERR_BAD_ARG, "sg_align_pairs: null argument");
    auto t_begin = std::chrono::steady_clock::now();
    std::unique_ptr<sg_result> res(new sg_result);
    res->n = n_pairs;
    res->has_cigar = !(flags & SG_FLAG_DISTANCE_ONLY);
    res->edit.assign(n_pairs, 0);
    res->refc.assign(n_pairs, 0);
    res->run_off.assign(n_pairs + 1, 0);
    const int nd = (int)ctx->devs.size();
    std::vector<ShardOut> shards(nd);
    if (n_pairs) {
        const std::vector<uint64_t> cut = split_by_weight(query_off, n_pairs, nd);
        auto work = [&](int k) {
            Device &d = ctx->devs[k];
            ShardOut &so = shards[k];
            if (cudaSetDevice(d.id) != cudaSuccess) { so.rc = SG_ERR_CUDA; so.err = "cudaSetDevice failed"; return; }
            uint64_t a = cut[k];
            while (a < cut[k + 1]) {  // batches bounded by query bases (slab size)
                uint64_t b = a + 1;
                while (b < cut[k + 1] && query_off[b + 1] - query_off[a] <= ctx->max_batch_query_bases) b++;
                so.rc = run_pairs_batch(ctx, d, text_blob, text_off, query_blob, query_off, a, b, flags, res.get(), so);
                if (so.rc) { so.err = g_last_error; return; }
                a = b;
            }
        };
        if (nd == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int k = 0; k < nd; k++) th.emplace_back(work, k);
            for (auto &t : th) t.join();
        }
        for (ShardOut &so : shards) if (so.rc) return fail(so.rc, so.err);
        if (res->has_cigar) finalize_runs(res.get(), shards);
    }
    double kms = 0;
    for (ShardOut &so : shards) kms = std::max(kms, so.kernel_ms);
    res->kernel_ns = (int64_t)(kms * 1e6);
    res->total_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
    *out = res.release();
    return SG_OK;
}

int sg_set_reference(sg_ctx *ctx, const char *genome_ascii, uint64_t genome_len)
{
    if (!ctx || (!genome_ascii && genome_len)) return fail(SG_ERR_BAD_ARG, "sg_set_reference: null argument");
    const int nd = (int)ctx->devs.size();
    std::vector<int> rcs(nd, SG_OK);
    std::vector<std::string> errs(nd);
    auto work = [&](int k) -> int {
        Device &d = ctx->devs[k];
        SG_CUDA(cudaSetDevice(d.id));
        d.has_genome = false;
        int rc = d.genome.reserve(sg_packed_words(genome_len) * 4 + 64);
        if (rc) return rc;
        rc = d.bad.reserve(16); if (rc) return rc;
        rc = d.h_small.reserve(64); if (rc) return rc;
        SG_CUDA(cudaMemsetAsync(d.bad.p, 0xFF, 16, d.stream));
        // upload in 256 Mbase pieces (a multiple of 16 bases, so every piece packs to whole words)
        const uint64_t piece = 256ull << 20;
        rc = d.ascii_t.reserve(std::min<uint64_t>(piece, genome_len) + 64); if (rc) return rc;
        for (uint64_t pos = 0; pos < genome_len || pos == 0; pos += piece) {
            const uint64_t len = std::min<uint64_t>(piece, genome_len - pos);
            SG_CUDA(cudaMemcpyAsync(d.ascii_t.p, genome_ascii + pos, len, cudaMemcpyHostToDevice, d.stream));
            rc = sg_dev_pack_2bit(d.ascii_t.as<char>(), len, d.genome.as<uint32_t>() + pos / 16, d.bad.as<uint64_t>(), d.stream);
            if (rc) return rc;
            uint64_t *h = d.h_small.as<uint64_t>();
            SG_CUDA(cudaMemcpyAsync(h, d.bad.p, 8, cudaMemcpyDeviceToHost, d.stream));
            SG_CUDA(cudaStreamSynchronize(d.stream));
            if (h[0] != ~0ull) return fail(SG_ERR_BAD_BASE, "non-ACGT character in reference at position " + std::to_string(pos + h[0]));
            if (genome_len == 0) break;
        }
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

}  // extern "C"

namespace {

int run_cand_batch(sg_ctx *ctx, Device &d, const char *rb, const uint64_t *roff, const uint64_t *cand_start,
                   const uint32_t *cand_read, uint64_t c0, uint64_t c1, uint32_t flags, sg_result *res, ShardOut &so)
{
    const uint64_t n = c1 - c0;
    const bool want_cigar = !(flags & SG_FLAG_DISTANCE_ONLY);
    cudaStream_t st = d.stream;
    // reads referenced by this batch: the contiguous index range [r0, r1]
    uint32_t r0 = cand_read[c0], r1 = cand_read[c0];
    for (uint64_t c = c0; c < c1; c++) { r0 = std::min(r0, cand_read[c]); r1 = std::max(r1, cand_read[c]); }
    const uint64_t nr = (uint64_t)r1 - r0 + 1;
    const uint64_t rbytes = roff[r1 + 1] - roff[r0];
    uint64_t slab_bytes = 0;
    if (want_cigar) for (uint64_t c = c0; c < c1; c++) slab_bytes += 2ull * (roff[cand_read[c] + 1] - roff[cand_read[c]]) + 8ull;
    int rc;
#define R(x) do { rc = (x); if (rc) return rc; } while (0)
    R(d.ascii_q.reserve(rbytes + 64)); R(d.packed_q.reserve(sg_packed_words(rbytes) * 4));
    R(d.qoff.reserve((nr + 1) * 8)); R(d.cstart.reserve(n * 8)); R(d.cread.reserve(n * 4)); R(d.cap32.reserve(n * 4));
    R(d.tstart.reserve(n * 8)); R(d.tlen.reserve(n * 8)); R(d.qstart.reserve(n * 8)); R(d.qlen.reserve(n * 8));
    R(d.slab_off.reserve((n + 1) * 8));
    if (want_cigar) R(d.slab.reserve(slab_bytes + 16));
    R(d.counter.reserve(8)); R(d.edit.reserve(n * 8)); R(d.refc.reserve(n * 8)); R(d.nruns.reserve(n * 4));
    R(d.status.reserve(n)); R(d.run_off.reserve((n + 1) * 8)); R(d.scan_tmp.reserve(sg_scan_tmp_bytes(n)));
    R(d.bad.reserve(16)); R(d.h_small.reserve(64));

    SG_CUDA(cudaMemcpyAsync(d.ascii_q.p, rb + roff[r0], rbytes, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d.qoff.p, roff + r0, (nr + 1) * 8, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d.cstart.p, cand_start + c0, n * 8, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemcpyAsync(d.cread.p, cand_read + c0, n * 4, cudaMemcpyHostToDevice, st));
    SG_CUDA(cudaMemsetAsync(d.bad.p, 0xFF, 16, st));
    R(sg_dev_pack_2bit(d.ascii_q.as<char>(), rbytes, d.packed_q.as<uint32_t>(), d.bad.as<uint64_t>() + 1, st));
    cand_descriptors_kernel<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(
        d.cstart.as<uint64_t>(), d.cread.as<uint32_t>(), d.qoff.as<uint64_t>(), r0, d.genome_len, n, d.tstart.as<uint64_t>(),
        d.tlen.as<uint64_t>(), d.qstart.as<uint64_t>(), d.qlen.as<uint64_t>(), d.cap32.as<uint32_t>());
    SG_CUDA(cudaGetLastError());
    // slab offsets = exclusive scan of the capacities
    R(sg_dev_scan_runs(d.cap32.as<uint32_t>(), n, d.slab_off.as<uint64_t>(), d.scan_tmp.p, st));
    SG_CUDA(cudaEventRecord(d.ev0, st));
    R(sg_dev_align(ctx->W, d.genome.as<uint32_t>(), d.tstart.as<uint64_t>(), d.tlen.as<uint64_t>(), d.packed_q.as<uint32_t>(),
                   d.qstart.as<uint64_t>(), d.qlen.as<uint64_t>(), n, flags, d.slab.as<uint8_t>(), d.slab_off.as<uint64_t>(),
                   d.counter.as<uint64_t>(), d.edit.as<int64_t>(), d.refc.as<uint64_t>(), d.nruns.as<uint32_t>(),
                   d.status.as<uint8_t>(), nullptr, st));
    SG_CUDA(cudaEventRecord(d.ev1, st));
    uint64_t *h = d.h_small.as<uint64_t>();
    SG_CUDA(cudaMemcpyAsync(h, d.bad.p, 16, cudaMemcpyDeviceToHost, st));
    if (want_cigar) {
        R(sg_dev_scan_runs(d.nruns.as<uint32_t>(), n, d.run_off.as<uint64_t>(), d.scan_tmp.p, st));
        SG_CUDA(cudaMemcpyAsync(h + 2, d.run_off.as<uint64_t>() + n, 8, cudaMemcpyDeviceToHost, st));
    }
    SG_CUDA(cudaMemcpyAsync(res->edit.data() + c0, d.edit.p, n * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaMemcpyAsync(res->refc.data() + c0, d.refc.p, n * 8, cudaMemcpyDeviceToHost, st));
    SG_CUDA(cudaStreamSynchronize(st));
    if (h[1] != ~0ull) {
        const uint64_t pos = h[1] + roff[r0];
        uint64_t r = (uint64_t)(std::upper_bound(roff + r0, roff + r1 + 2, pos) - roff) - 1;
        return fail(SG_ERR_BAD_BASE, "non-ACGT character in read " + std::to_string(r) + " at position " + std::to_string(pos - roff[r]));
    }
    float ms = 0;
    SG_CUDA(cudaEventElapsedTime(&ms, d.ev0, d.ev1));
    so.kernel_ms += ms;
    if (want_cigar) {
        const uint64_t total_runs = h[2];
        R(d.runs.reserve(total_runs + 16));
        R(sg_dev_gather_runs(d.slab.as<uint8_t>(), d.slab_off.as<uint64_t>(), d.nruns.as<uint32_t>(), d.run_off.as<uint64_t>(),
                             n, d.runs.as<uint8_t>(), st));
        std::vector<uint8_t> piece(total_runs), status(n);
        SG_CUDA(cudaMemcpyAsync(piece.data(), d.runs.p, total_runs, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(res->run_off.data() + c0, d.run_off.p, n * 8, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaMemcpyAsync(status.data(), d.status.p, n, cudaMemcpyDeviceToHost, st));
        SG_CUDA(cudaStreamSynchronize(st));
        for (uint64_t k = 0; k < n; k++)
            if (status[k]) return fail(SG_ERR_CIGAR_OVERFLOW, "alignment " + std::to_string(c0 + k) + " exceeded its run capacity");
        so.pieces.push_back(std::move(piece));
        so.piece_first.push_back(c0);
    }
#undef R
    return SG_OK;
}

}  // namespace

extern "C" {

int sg_align_candidates(sg_ctx *ctx, const char *read_blob, const uint64_t *read_off, uint64_t n_reads,
                        const uint64_t *cand_start, const uint32_t *cand_read, uint64_t n_cand, uint32_t flags,
                        sg_result **out)
{
    if (!ctx || !out || !read_off || (n_cand && (!cand_start || !cand_read)))
        return fail(SG_ERR_BAD_ARG, "sg_align_candidates: null argument");
    for (Device &d : ctx->devs)
        if (!d.has_genome) return fail(SG_ERR_NO_REFERENCE, "sg_align_candidates: call sg_set_reference first");
    const uint64_t genome_len = ctx->devs[0].genome_len;
    for (uint64_t c = 0; c < n_cand; c++) {
        if (cand_read[c] >= n_reads) return fail(SG_ERR_BAD_ARG, "candidate " + std::to_string(c) + ": read index out of range");
        if (cand_start[c] > genome_len) return fail(SG_ERR_BAD_ARG, "candidate " + std::to_string(c) + ": start beyond the reference");
    }
    auto t_begin = std::chrono::steady_clock::now();
    std::unique_ptr<sg_result> res(new sg_result);
    res->n = n_cand;
    res->has_cigar = !(flags & SG_FLAG_DISTANCE_ONLY);
    res->edit.assign(n_cand, 0);
    res->refc.assign(n_cand, 0);
    res->run_off.assign(n_cand + 1, 0);
    const int nd = (int)ctx->devs.size();
    std::vector<ShardOut> shards(nd);
    if (n_cand) {
        // weight of a candidate = its read's length
        std::vector<uint64_t> woff(n_cand + 1, 0);
        for (uint64_t c = 0; c < n_cand; c++) woff[c + 1] = woff[c] + (read_off[cand_read[c] + 1] - read_off[cand_read[c]]);
        const std::vector<uint64_t> cut = split_by_weight(woff.data(), n_cand, nd);
        auto work = [&](int k) {
            Device &d = ctx->devs[k];
            ShardOut &so = shards[k];
            if (cudaSetDevice(d.id) != cudaSuccess) { so.rc = SG_ERR_CUDA; so.err = "cudaSetDevice failed"; return; }
            uint64_t a = cut[k];
            while (a < cut[k + 1]) {
                uint64_t b = a + 1;
                while (b < cut[k + 1] && woff[b + 1] - woff[a] <= ctx->max_batch_query_bases) b++;
                so.rc = run_cand_batch(ctx, d, read_blob, read_off, cand_start, cand_read, a, b, flags, res.get(), so);
                if (so.rc) { so.err = g_last_error; return; }
                a = b;
            }
        };
        if (nd == 1) work(0);
        else {
            std::vector<std::thread> th;
            for (int k = 0; k < nd; k++) th.emplace_back(work, k);
            for (auto &t : th) t.join();
        }
        for (ShardOut &so : shards) if (so.rc) return fail(so.rc, so.err);
        if (res->has_cigar) finalize_runs(res.get(), shards);
    }
    double kms = 0;
    for (ShardOut &so : shards) kms = std::max(kms, so.kernel_ms);
    res->kernel_ns = (int64_t)(kms * 1e6);
    res->total_ns = std::chrono::duration_cast<std::chrono::nanoseconds>(std::chrono::steady_clock::now() - t_begin).count();
    *out = res.release();
    return SG_OK;
}

uint64_t sg_result_count(const sg_result *r) { return r ? r->n : 0; }
const int64_t *sg_result_edit_distances(const sg_result *r) { return r ? r->edit.data() : nullptr; }
const uint64_t *sg_result_ref_consumed(const sg_result *r) { return r ? r->refc.data() : nullptr; }
const uint64_t *sg_result_run_offsets(const sg_result *r) { return r && r->has_cigar ? r->run_off.data() : nullptr; }

const uint8_t *sg_result_runs(const sg_result *r)
{
    if (!r || !r->has_cigar) return nullptr;
    if (r->pieces.size() == 1) return r->pieces[0].data();
    sg_result *m = const_cast<sg_result *>(r);
    if (m->flat.empty() && r->run_off[r->n]) {
        m->flat.reserve(r->run_off[r->n]);
        for (const auto &p : r->pieces) m->flat.insert(m->flat.end(), p.begin(), p.end());
    }
    return m->flat.data();
}

int64_t sg_result_kernel_ns(const sg_result *r) { return r ? r->kernel_ns : 0; }
int64_t sg_result_total_ns(const sg_result *r) { return r ? r->total_ns : 0; }

uint64_t sg_result_cigar_len(const sg_result *r, uint64_t idx)
{
    if (!r || !r->has_cigar || idx >= r->n) return 0;
    uint64_t cnt;
    const uint8_t *p = runs_of(r, idx, &cnt);
    uint64_t len = 0;
    for (uint64_t k = 0; k < cnt; k++) len += SG_RUN_COUNT(p[k]) >= 10 ? 3 : 2;
    return len;
}

int64_t sg_result_render_cigar(const sg_result *r, uint64_t idx, char *buf, uint64_t cap)
{
    if (!r || !r->has_cigar || idx >= r->n || !buf) return -1;
    static const char ops[4] = {'=', 'X', 'I', 'D'};
    uint64_t cnt;
    const uint8_t *p = runs_of(r, idx, &cnt);
    uint64_t len = 0;
    for (uint64_t k = 0; k < cnt; k++) {
        const unsigned c = SG_RUN_COUNT(p[k]);
        if (len + (c >= 10 ? 3u : 2u) + 1u > cap) return -1;
        if (c >= 10) buf[len++] = (char)('0' + c / 10);
        buf[len++] = (char)('0' + c % 10);
        buf[len++] = ops[SG_RUN_OP(p[k])];
    }
    if (len + 1 > cap) return -1;
    buf[len] = '\0';
    return (int64_t)len;
}

int64_t sg_result_entries(const sg_result *r, uint64_t idx, sg_cigar_entry *out, uint64_t cap)
{
    if (!r || !r->has_cigar || idx >= r->n || !out) return -1;
    static const char ops[4] = {'=', 'X', 'I', 'D'};
    uint64_t cnt;
    const uint8_t *p = runs_of(r, idx, &cnt);
    if (cnt > cap) return -1;
    for (uint64_t k = 0; k < cnt; k++) {
        out[k].edit_count = (uint8_t)SG_RUN_COUNT(p[k]);
        out[k].edit_type = ops[SG_RUN_OP(p[k])];
    }
    return (int64_t)cnt;
}

void sg_result_free(sg_result *r) { delete r; }

}  // extern "C"
