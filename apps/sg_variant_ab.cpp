// sg_variant_ab -- A/B of the alignment kernel's run-emission variants (bytes / SG_FLAG_RUN_WORDS / no run storage) on device-resident
// synthetic workloads, in seconds and without Python: every variant must produce the same distances, consumed prefixes, run
// counts and run bytes as the default, and its kernel time (CUDA events, best and mean of `reps` launches after a warm-up)
// is printed beside the default's.  One JSON line per workload on stdout.
//
// Workloads (generator: libscrooge_b200_bench.so, identical to bench.py's):
//   long        n x 10 kbp pairs at 10 % (the headline shape)
//   stress      the same reads, 7 of 8 against the text of ANOTHER pair (what a spurious candidate location of read
//               mapping looks like to the kernel: windows that are mostly edits, ~20 runs per window)
//   short_w64   n x 150 bp pairs at 5 %, W=64        short_w32   the same at W=32/O=17
// Also: sg_set_reference + sg_align_candidates through the host API with SG_EMIT=bytes and =words on a small case.
//
//   build/sg_variant_ab [--pairs 303104] [--short-pairs 4000000] [--reps 3]
#include <cuda_runtime.h>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "scrooge_b200.h"
#include "scrooge_b200_bench.h"

#define CK(x)                                                                                          \
    do {                                                                                               \
        cudaError_t e_ = (x);                                                                          \
        if (e_ != cudaSuccess) { fprintf(stderr, "%s:%d %s: %s\n", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); exit(2); } \
    } while (0)
#define SG(x)                                                                                          \
    do {                                                                                               \
        int r_ = (x);                                                                                  \
        if (r_ != 0) { fprintf(stderr, "%s:%d %s -> %d: %s / %s\n", __FILE__, __LINE__, #x, r_, sg_last_error(), sg_bench_last_error()); exit(3); } \
    } while (0)

static double now_s() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

template <typename T> static T *dalloc(size_t n)
{
    void *p = nullptr;
    CK(cudaMalloc(&p, std::max<size_t>(n, 1) * sizeof(T)));
    return (T *)p;
}

struct Outputs {
    std::vector<int64_t> edit;
    std::vector<uint64_t> refc;
    std::vector<uint32_t> nruns;
    std::vector<uint8_t> status, runs;
    std::vector<uint64_t> run_off;
    float best_ms = 0, mean_ms = 0;
};

struct Batch {
    int W = 64;
    uint64_t n = 0, L = 0, cap = 0;
    uint32_t *ptext = nullptr, *pquery = nullptr;
    uint64_t *tstart = nullptr, *tlen = nullptr, *qstart = nullptr, *qlen = nullptr, *slab_off = nullptr;
    uint8_t *slab = nullptr;
};

// one variant: warm-up + reps timed launches, then compaction of the last launch and the outputs on the host
static Outputs run_variant(const Batch &b, uint32_t flags, int reps, uint64_t sample, cudaStream_t st)
{
    Outputs o;
    const uint64_t n = b.n;
    uint64_t *counter = dalloc<uint64_t>(1), *refc = dalloc<uint64_t>(n), *run_off = dalloc<uint64_t>(n + 1);
    int64_t *edit = dalloc<int64_t>(n);
    uint32_t *nruns = dalloc<uint32_t>(n);
    uint8_t *status = dalloc<uint8_t>(n);
    void *scan_tmp = dalloc<uint8_t>(sg_scan_tmp_bytes(n));
    CK(cudaMemsetAsync(b.slab, 0xEE, b.n * b.cap, st));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    float best = 1e30f, sum = 0;
    for (int r = -1; r < reps; r++) {
        CK(cudaEventRecord(e0, st));
        SG(sg_dev_align(b.W, b.ptext, b.tstart, b.tlen, b.pquery, b.qstart, b.qlen, n, flags, b.slab, b.slab_off, counter, edit, refc, nruns, status,
                        nullptr, nullptr, st));
        CK(cudaEventRecord(e1, st));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (r >= 0) { best = std::min(best, ms); sum += ms; }
    }
    o.best_ms = best; o.mean_ms = sum / reps;
    SG(sg_dev_scan_runs(nruns, n, run_off, scan_tmp, st));
    o.edit.resize(n); o.refc.resize(n); o.nruns.resize(n); o.status.resize(n); o.run_off.resize(sample + 1);
    CK(cudaMemcpyAsync(o.edit.data(), edit, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o.refc.data(), refc, n * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o.nruns.data(), nruns, n * 4, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o.status.data(), status, n, cudaMemcpyDeviceToHost, st));
    CK(cudaMemcpyAsync(o.run_off.data(), run_off, (sample + 1) * 8, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // dense runs of the first `sample` alignments
    const uint64_t total = o.run_off[sample];
    uint8_t *runs = dalloc<uint8_t>(total + 16);
    SG(sg_dev_gather_runs(b.slab, b.slab_off, nruns, run_off, sample, runs, st));
    o.runs.resize(total);
    CK(cudaMemcpyAsync(o.runs.data(), runs, total, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    for (void *p : {(void *)counter, (void *)refc, (void *)run_off, (void *)edit, (void *)nruns, (void *)status, scan_tmp, (void *)runs}) CK(cudaFree(p));
    CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    return o;
}

static bool same(const Outputs &a, const Outputs &b)
{
    return a.edit == b.edit && a.refc == b.refc && a.nruns == b.nruns && a.status == b.status && a.run_off == b.run_off && a.runs == b.runs;
}

static void ab(const char *name, const Batch &b, int reps, cudaStream_t st)
{
    const uint64_t sample = std::min<uint64_t>(b.n, 32768);
    Outputs bytes = run_variant(b, 0u, reps, sample, st);
    Outputs words = run_variant(b, SG_FLAG_RUN_WORDS, reps, sample, st);
    // no run storage at all (traceback and run counting still happen): what any emission scheme can save at most
    float dist_only_ms = 1e30f;
    {
        uint64_t *counter = dalloc<uint64_t>(1), *refc = dalloc<uint64_t>(b.n);
        int64_t *edit = dalloc<int64_t>(b.n);
        uint32_t *nruns = dalloc<uint32_t>(b.n);
        uint8_t *status = dalloc<uint8_t>(b.n);
        cudaEvent_t e0, e1;
        CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
        for (int r = -1; r < reps; r++) {
            CK(cudaEventRecord(e0, st));
            SG(sg_dev_align(b.W, b.ptext, b.tstart, b.tlen, b.pquery, b.qstart, b.qlen, b.n, SG_FLAG_DISTANCE_ONLY, nullptr, nullptr, counter, edit, refc, nruns,
                            status, nullptr, nullptr, st));
            CK(cudaEventRecord(e1, st));
            CK(cudaEventSynchronize(e1));
            float ms = 0;
            CK(cudaEventElapsedTime(&ms, e0, e1));
            if (r >= 0) dist_only_ms = std::min(dist_only_ms, ms);
        }
        std::vector<int64_t> ed2(b.n);
        CK(cudaMemcpy(ed2.data(), edit, b.n * 8, cudaMemcpyDeviceToHost));
        if (ed2 != bytes.edit) dist_only_ms = -1.0f;   // distance-only must give the same distances
        for (void *p : {(void *)counter, (void *)refc, (void *)edit, (void *)nruns, (void *)status}) CK(cudaFree(p));
        CK(cudaEventDestroy(e0)); CK(cudaEventDestroy(e1));
    }
    uint64_t total_runs = 0, bad_status = 0;
    double ed = 0;
    for (uint64_t k = 0; k < b.n; k++) { total_runs += bytes.nruns[k]; bad_status += bytes.status[k] != 0; ed += (double)bytes.edit[k]; }
    printf("{\"workload\": \"%s\", \"W\": %d, \"alignments\": %llu, \"read_len\": %llu, \"runs_per_alignment\": %.1f, \"mean_edit\": %.1f, \"status_nonzero\": %llu, "
           "\"identical\": {\"words\": %s}, \"kernel_ms_best_mean\": {\"bytes\": [%.4f, %.4f], \"words\": [%.4f, %.4f]}, \"distance_only_ms\": %.4f, "
           "\"alignments_per_s\": {\"bytes\": %.4g, \"words\": %.4g}, "
           "\"speedup_over_bytes\": {\"words\": %.4f}, \"compared\": \"edit, ref_consumed, nruns, status of every alignment; run bytes of the first %llu\"}\n",
           name, b.W, (unsigned long long)b.n, (unsigned long long)b.L, (double)total_runs / (double)b.n, ed / (double)b.n, (unsigned long long)bad_status,
           same(bytes, words) ? "true" : "false", bytes.best_ms, bytes.mean_ms, words.best_ms, words.mean_ms, dist_only_ms,
           b.n / (bytes.best_ms * 1e-3), b.n / (words.best_ms * 1e-3), bytes.best_ms / words.best_ms, (unsigned long long)sample);
    fflush(stdout);
}

// pairs of the synthetic generator, packed and described on the device; stress: 7 of 8 reads against another pair's text
static Batch make_pairs(int W, uint64_t n, uint32_t L, double err, uint32_t ws, uint32_t wi, uint32_t wd, bool stress, cudaStream_t st)
{
    Batch b;
    b.W = W; b.n = n; b.L = L;
    b.cap = (2ull * L + 8ull + 3ull) & ~3ull;   // slots on 4-byte boundaries (SG_FLAG_RUN_WORDS)
    const uint64_t stride = sg_synth_text_stride(L, 64);
    char *text = dalloc<char>(n * stride), *reads = dalloc<char>(n * L);
    b.tlen = dalloc<uint64_t>(n);
    SG(sg_dev_synth_pairs(0x5C2006Eull + 3, 0, n, L, err, ws, wi, wd, 64, text, stride, b.tlen, reads, st));
    b.ptext = dalloc<uint32_t>(sg_packed_words(n * stride));
    b.pquery = dalloc<uint32_t>(sg_packed_words(n * (uint64_t)L));
    uint64_t *bad = dalloc<uint64_t>(2);
    CK(cudaMemsetAsync(bad, 0xFF, 16, st));
    SG(sg_dev_pack_2bit(text, n * stride, b.ptext, bad, st));
    SG(sg_dev_pack_2bit(reads, n * (uint64_t)L, b.pquery, bad + 1, st));
    std::vector<uint64_t> h(4 * n + 1);
    uint64_t *ts = h.data(), *qs = ts + n, *ql = qs + n, *so = ql + n;
    for (uint64_t k = 0; k < n; k++) {
        ts[k] = k * stride;
        const uint64_t q = (stress && (k & 7u)) ? (k * 7919ull + 13ull) % n : k;
        qs[k] = q * (uint64_t)L;
        ql[k] = L;
        so[k] = k * b.cap;
    }
    so[n] = n * b.cap;
    b.tstart = dalloc<uint64_t>(n); b.qstart = dalloc<uint64_t>(n); b.qlen = dalloc<uint64_t>(n); b.slab_off = dalloc<uint64_t>(n + 1);
    CK(cudaMemcpyAsync(b.tstart, ts, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b.qstart, qs, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b.qlen, ql, n * 8, cudaMemcpyHostToDevice, st));
    CK(cudaMemcpyAsync(b.slab_off, so, (n + 1) * 8, cudaMemcpyHostToDevice, st));
    b.slab = dalloc<uint8_t>(n * b.cap + 16);
    uint64_t hbad[2];
    CK(cudaMemcpyAsync(hbad, bad, 16, cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    if (hbad[0] != ~0ull || hbad[1] != ~0ull) { fprintf(stderr, "generator produced a bad base\n"); exit(4); }
    CK(cudaFree(text)); CK(cudaFree(reads)); CK(cudaFree(bad));
    return b;
}

static void free_batch(Batch &b)
{
    for (void *p : {(void *)b.ptext, (void *)b.pquery, (void *)b.tstart, (void *)b.tlen, (void *)b.qstart, (void *)b.qlen, (void *)b.slab_off, (void *)b.slab})
        if (p) CK(cudaFree(p));
    b = Batch();
}

// host API: the same candidate list through sg_align_candidates with SG_EMIT=bytes and SG_EMIT=words
static void host_api_check()
{
    const uint64_t G = 4000000, n_reads = 4096, L = 1000, ncand = 4;
    std::vector<char> genome(G), reads(n_reads * L);
    std::vector<uint64_t> pos(n_reads), roff(n_reads + 1), cs(n_reads * ncand);
    std::vector<uint32_t> cr(n_reads * ncand);
    SG(sg_synth_genome(77, 0, G, genome.data(), nullptr, nullptr));
    SG(sg_synth_reads(78, 0, n_reads, (uint32_t)L, 0.10, 6, 50, 54, genome.data(), G, reads.data(), pos.data(), 0, nullptr));
    for (uint64_t r = 0; r <= n_reads; r++) roff[r] = r * L;
    for (uint64_t r = 0; r < n_reads; r++)
        for (uint64_t c = 0; c < ncand; c++) {
            cs[r * ncand + c] = c == 0 ? pos[r] : (pos[r] * 2654435761ull + c * 40503ull) % (G - 3 * L);   // the true locus + unrelated ones
            cr[r * ncand + c] = (uint32_t)r;
        }
    std::vector<int64_t> ed[2];
    std::vector<uint8_t> runs[2];
    std::vector<uint64_t> ro[2];
    double ms[2];
    const char *modes[2] = {"bytes", "words"};
    for (int m = 0; m < 2; m++) {
        setenv("SG_EMIT", modes[m], 1);
        sg_ctx *ctx = nullptr;
        int dev0 = 0;
        SG(sg_ctx_create(&ctx, &dev0, 1, 64));
        SG(sg_set_reference(ctx, genome.data(), G));
        sg_result *res = nullptr;
        for (int rep = 0; rep < 2; rep++) {
            if (res) sg_result_free(res);
            const double t0 = now_s();
            SG(sg_align_candidates(ctx, reads.data(), roff.data(), n_reads, cs.data(), cr.data(), n_reads * ncand, 0, &res));
            ms[m] = (now_s() - t0) * 1e3;
        }
        const uint64_t n = n_reads * ncand;
        ed[m].assign(sg_result_edit_distances(res), sg_result_edit_distances(res) + n);
        ro[m].assign(sg_result_run_offsets(res), sg_result_run_offsets(res) + n + 1);
        runs[m].assign(sg_result_runs(res), sg_result_runs(res) + ro[m][n]);
        sg_result_free(res);
        sg_ctx_destroy(ctx);
    }
    unsetenv("SG_EMIT");
    const bool ok = ed[0] == ed[1] && ro[0] == ro[1] && runs[0] == runs[1];
    printf("{\"workload\": \"host_api_candidates\", \"alignments\": %llu, \"identical\": %s, \"call_ms\": {\"bytes\": %.2f, \"words\": %.2f}, \"runs_total\": %llu}\n",
           (unsigned long long)(n_reads * ncand), ok ? "true" : "false", ms[0], ms[1], (unsigned long long)ro[0][n_reads * ncand]);
    fflush(stdout);
}

int main(int argc, char **argv)
{
    uint64_t pairs = 303104, short_pairs = 4000000;
    int reps = 3;
    for (int a = 1; a + 1 < argc; a += 2) {
        if (!strcmp(argv[a], "--pairs")) pairs = strtoull(argv[a + 1], nullptr, 10);
        else if (!strcmp(argv[a], "--short-pairs")) short_pairs = strtoull(argv[a + 1], nullptr, 10);
        else if (!strcmp(argv[a], "--reps")) reps = atoi(argv[a + 1]);
    }
    const double t0 = now_s();
    CK(cudaSetDevice(0));
    cudaStream_t st;
    CK(cudaStreamCreate(&st));
    fprintf(stderr, "[ab] device ready after %.2f s\n", now_s() - t0);
    {
        Batch b = make_pairs(64, pairs, 10000, 0.10, 6, 50, 54, false, st);
        ab("long_10kbp", b, reps, st);
        free_batch(b);
        b = make_pairs(64, pairs, 10000, 0.10, 6, 50, 54, true, st);
        ab("stress_10kbp_1true_7unrelated", b, reps, st);
        free_batch(b);
    }
    fprintf(stderr, "[ab] long workloads done after %.2f s\n", now_s() - t0);
    {
        Batch b = make_pairs(64, short_pairs, 150, 0.05, 90, 5, 5, false, st);
        ab("short_150bp_w64", b, reps, st);
        b.W = 32;
        ab("short_150bp_w32", b, reps, st);
        free_batch(b);
    }
    fprintf(stderr, "[ab] short workloads done after %.2f s\n", now_s() - t0);
    host_api_check();
    fprintf(stderr, "[ab] all done after %.2f s\n", now_s() - t0);
    return 0;
}
