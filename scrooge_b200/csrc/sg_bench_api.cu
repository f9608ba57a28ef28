// sg_bench_api.cu -- libscrooge_b200_bench.so: the C entry points of include/scrooge_b200_bench.h.
// Measurement, synthetic-data and checking helpers used by bench.py, the tests and the apps; none of it is on the
// alignment path, and the product library (libscrooge_b200.so) neither contains nor needs it -- nor does this library
// need the product one: it links against the CUDA runtime only and keeps its own error string (sg_bench_last_error).
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <algorithm>
#include <cuda_runtime.h>

#include "../../include/scrooge_b200_bench.h"
#include "sg_bench_aux.cuh"

using namespace sg;

namespace {

// status codes as in scrooge_b200.h
constexpr int SG_OK = 0, SG_ERR_CUDA = 1, SG_ERR_BAD_ARG = 3;
thread_local std::string g_bench_error;

int fail(int code, const std::string &msg)
{
    g_bench_error = msg;
    return code;
}

#define SG_CUDA(call)                                                                                  \
    do {                                                                                               \
        cudaError_t sg_e_ = (call);                                                                    \
        if (sg_e_ != cudaSuccess) {                                                                    \
            cudaGetLastError();                                                                        \
            return fail(SG_ERR_CUDA, std::string(#call ": ") + cudaGetErrorString(sg_e_));             \
        }                                                                                              \
    } while (0)

int num_sms(int *out)
{
    int dev = 0;
    SG_CUDA(cudaGetDevice(&dev));
    SG_CUDA(cudaDeviceGetAttribute(out, cudaDevAttrMultiProcessorCount, dev));
    return SG_OK;
}

}  // namespace

extern "C" {

const char *sg_bench_last_error(void) { return g_bench_error.c_str(); }

int sg_dev_check_runs(const uint8_t *d_runs, const uint64_t *d_run_off, uint64_t n, const uint64_t *d_query_len,
                      const int64_t *d_edit, const uint64_t *d_ref_consumed, uint32_t max_count, uint64_t *d_n_bad, void *stream)
{
    if (n == 0) return SG_OK;
    if (!d_runs || !d_run_off || !d_query_len || !d_edit || !d_ref_consumed || !d_n_bad)
        return fail(SG_ERR_BAD_ARG, "sg_dev_check_runs: null pointer");
    int sms = 0;
    if (int rc = num_sms(&sms)) return rc;
    check_runs_kernel<<<sms * 8, 256, 0, (cudaStream_t)stream>>>(d_runs, d_run_off, n, d_query_len, d_edit, d_ref_consumed, max_count,
                                                                      (unsigned long long *)d_n_bad);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_dev_int32_peak(int kind, double ms, double *gops)
{
    if (kind < 0 || kind > 6 || !gops) return fail(SG_ERR_BAD_ARG, "sg_dev_int32_peak: bad argument");
    int sms = 0;
    if (int rc = num_sms(&sms)) return rc;
    uint32_t *sink = nullptr;
    SG_CUDA(cudaMalloc(&sink, 64));
    cudaEvent_t e0, e1;
    SG_CUDA(cudaEventCreate(&e0));
    SG_CUDA(cudaEventCreate(&e1));
    const int blocks = sms * 8, threads = 256;
    auto launch = [&](int iters) {
        switch (kind) {
            case 0: int32_peak_kernel<0><<<blocks, threads>>>(sink, iters, 1u); break;
            case 1: int32_peak_kernel<1><<<blocks, threads>>>(sink, iters, 1u); break;
            case 2: int32_peak_kernel<2><<<blocks, threads>>>(sink, iters, 1u); break;
            case 3: int32_peak_kernel<3><<<blocks, threads>>>(sink, iters, 1u); break;
            case 4: int32_peak_kernel<4><<<blocks, threads>>>(sink, iters, 1u); break;
            case 5: int32_peak_kernel<5><<<blocks, threads>>>(sink, iters, 1u); break;
            default: int32_peak_kernel<6><<<blocks, threads>>>(sink, iters, 1u); break;
        }
    };
    int iters = 2048;
    float t = 0.f;
    for (int round = 0; round < 6; round++) {  // grow until the launch lasts about `ms`
        launch(iters);  // warm-up at this size
        cudaEventRecord(e0);
        launch(iters);
        cudaEventRecord(e1);
        SG_CUDA(cudaEventSynchronize(e1));
        SG_CUDA(cudaEventElapsedTime(&t, e0, e1));
        if (t >= ms * 0.5 || iters >= (1 << 24)) break;
        double scale = ms / (t > 1e-3 ? t : 1e-3);
        iters = (int)std::min<double>((double)iters * std::min(scale, 16.0), (double)(1 << 24));
    }
    const double ops = (double)blocks * threads * (double)iters * (double)kPeakOpsPerIter[kind];
    *gops = ops / ((double)t * 1e-3) / 1e9;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    return SG_OK;
}

uint64_t sg_synth_text_stride(uint32_t read_len, uint32_t slack)
{
    uint64_t s = sg_synth_stride(read_len, slack);
    return (s + 15ull) & ~15ull;
}

static SgSynthParams make_synth(uint64_t seed, uint32_t read_len, double err, uint32_t w_sub, uint32_t w_ins,
                                uint32_t w_del, uint32_t slack)
{
    SgSynthParams p;
    p.seed = seed;
    p.read_len = read_len;
    double e = err < 0 ? 0 : (err > 1 ? 1 : err);
    double thr = e * 4294967296.0;
    p.err_threshold = thr >= 4294967295.0 ? 0xFFFFFFFFu : (uint32_t)thr;
    p.w_sub = w_sub; p.w_ins = w_ins; p.w_del = w_del;
    p.slack = slack;
    return p;
}

int sg_synth_pairs_host(uint64_t seed, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, double err,
                        uint32_t w_sub, uint32_t w_ins, uint32_t w_del, uint32_t slack, char *text,
                        uint64_t text_stride, uint64_t *text_len, char *reads)
{
    if (text_stride < sg_synth_stride(read_len, slack)) return fail(SG_ERR_BAD_ARG, "text_stride too small");
    const SgSynthParams p = make_synth(seed, read_len, err, w_sub, w_ins, w_del, slack);
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < (long long)n_pairs; k++) {
        char *t = text + (uint64_t)k * text_stride;
        const uint64_t tl = sg_synth_pair(p, first_pair + (uint64_t)k, t, reads + (uint64_t)k * read_len);
        text_len[k] = tl;
        memset(t + tl, 'A', text_stride - tl);  // keep the whole slot packable
    }
    return SG_OK;
}

int sg_dev_synth_pairs(uint64_t seed, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, double err,
                       uint32_t w_sub, uint32_t w_ins, uint32_t w_del, uint32_t slack, char *d_text,
                       uint64_t text_stride, uint64_t *d_text_len, char *d_reads, void *stream)
{
    if (text_stride < sg_synth_stride(read_len, slack)) return fail(SG_ERR_BAD_ARG, "text_stride too small");
    if (n_pairs == 0) return SG_OK;
    const SgSynthParams p = make_synth(seed, read_len, err, w_sub, w_ins, w_del, slack);
    const unsigned blocks = (unsigned)((n_pairs + 127ull) / 128ull);
    synth_pairs_kernel<<<blocks, 128, 0, (cudaStream_t)stream>>>(p, first_pair, n_pairs, d_text, text_stride,
                                                                d_text_len, d_reads);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_synth_genome(uint64_t seed, uint64_t first, uint64_t n, char *out, void *d_out, void *stream)
{
    if (out) {
#pragma omp parallel for schedule(static)
        for (long long i = 0; i < (long long)n; i++) out[i] = sg_synth_genome_base(seed, first + (uint64_t)i);
    }
    if (d_out && n) {
        int sms = 0;
        if (int rc = num_sms(&sms)) return rc;
        synth_genome_kernel<<<sms * 8, 256, 0, (cudaStream_t)stream>>>(seed, first, n, (char *)d_out);
        SG_CUDA(cudaGetLastError());
    }
    return SG_OK;
}

int sg_synth_reads(uint64_t seed, uint64_t first_read, uint64_t n_reads, uint32_t read_len, double err, uint32_t w_sub, uint32_t w_ins,
                   uint32_t w_del, const char *genome, uint64_t genome_len, char *reads, uint64_t *pos, int on_device, void *stream)
{
    if (genome_len <= 2ull * read_len + 64ull) return fail(SG_ERR_BAD_ARG, "genome too short for this read length");
    if (n_reads == 0) return SG_OK;
    const SgSynthParams p = make_synth(seed, read_len, err, w_sub, w_ins, w_del, 0);
    if (on_device) {
        synth_reads_kernel<<<(unsigned)((n_reads + 127ull) / 128ull), 128, 0, (cudaStream_t)stream>>>(p, first_read, n_reads, genome, genome_len,
                                                                                                reads, pos);
        SG_CUDA(cudaGetLastError());
        return SG_OK;
    }
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < (long long)n_reads; k++)
        pos[k] = sg_synth_read_from_genome(p, first_read + (uint64_t)k, genome, genome_len, reads + (uint64_t)k * read_len);
    return SG_OK;
}

}  // extern "C"
