// sg_e2e_probe.cpp -- where the END-TO-END time goes, and what the host could deliver at best.
//
// The end-to-end path (host ASCII in, distances + CIGAR runs out) is bound by the host, not by the GPUs:
// this tool measures, on the box it runs on and for 1..N GPUs at once,
//   1. the topology the library sees (CPUs, NUMA-local CPU lists and PCIe link of every GPU);
//   2. the host-side CEILINGS with all GPUs busy at the same time:
//        h2d    pinned ASCII -> device, copies only          (what the copy engines can pull out of host DRAM)
//        pack   ASCII -> 2 bit/base by the packer threads    (what the cores can pack)
//        both   the two at once on disjoint halves           (they share the host's DRAM bandwidth)
//      ceiling of the adaptive ingest = h2d + 0.75 * pack bytes/s (a packed chunk still crosses PCIe at a quarter size);
//   3. the end-to-end rate through sg_align_pairs under the ingest policies
//        adaptive (default) | dma (no packer threads: SG_HOST_THREADS=0) | hostpack (no ASCII copies: SG_DMA_DEPTH=0)
//      with and without CPU binding, each with the per-call breakdown of sg_call_stats, plus the rendered variant
//      (sg_result_render_all: CIGAR text for every alignment, what the reference's CPU arm includes).
// Two process layouts: --layout threads (ONE context over N GPUs: the library's own scatter) and --layout procs (N
// processes with one GPU each, as bench.py under torchrun), forked before CUDA is touched and synchronised through a
// process-shared barrier.
//
//   build/sg_e2e_probe --gpus 8 --pairs 262144 --len 10000 --steps 3 --layout procs [--modes adaptive,dma,hostpack]
// One JSON object per line on stdout.
#include <sys/mman.h>
#include <sys/wait.h>
#include <pthread.h>
#include <sched.h>
#include <unistd.h>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime_api.h>

#include "scrooge_b200.h"
#include "scrooge_b200_bench.h"

using clk = std::chrono::steady_clock;
static double now_s() { return std::chrono::duration<double>(clk::now().time_since_epoch()).count(); }

struct Args {
    int gpus = 1, steps = 3, len = 10000;
    uint64_t pairs = 262144;   // per GPU
    std::string layout = "threads", modes = "adaptive,adaptive_all,adaptive_half,dma,hostpack,adaptive_nobind";
    bool ceilings = true, render = true;
};

static std::string slurp(const std::string &path)
{
    std::ifstream f(path);
    std::stringstream ss;
    ss << f.rdbuf();
    std::string s = ss.str();
    while (!s.empty() && (s.back() == '\n' || s.back() == ' ')) s.pop_back();
    return s;
}

// ---- one GPU's shard of the workload in pinned host memory -------------------------------------------------------------
struct Shard {
    char *text = nullptr, *query = nullptr;
    std::vector<uint64_t> toff, qoff;
    uint64_t n = 0;
    uint64_t bytes() const { return toff[n] + qoff[n]; }
};

static void make_shard(Shard &s, uint64_t first_pair, uint64_t n, int len, int threads)
{
    const uint64_t stride = sg_synth_text_stride((uint32_t)len, 64);
    std::vector<char> text(n * stride), reads(n * (uint64_t)len);
    std::vector<uint64_t> tlen(n);
    (void)threads;
    if (sg_synth_pairs_host(0x5C2006E + 3, first_pair, n, (uint32_t)len, 0.10, 6, 50, 54, 64, text.data(), stride, tlen.data(), reads.data())) {
        fprintf(stderr, "synth failed: %s\n", sg_bench_last_error());
        exit(1);
    }
    s.n = n;
    s.toff.assign(n + 1, 0);
    s.qoff.assign(n + 1, 0);
    for (uint64_t k = 0; k < n; k++) { s.toff[k + 1] = s.toff[k] + tlen[k]; s.qoff[k + 1] = s.qoff[k] + (uint64_t)len; }
    s.text = (char *)sg_host_alloc(s.toff[n] + 64);
    s.query = (char *)sg_host_alloc(s.qoff[n] + 64);
    if (!s.text || !s.query) { fprintf(stderr, "pinned allocation failed: %s\n", sg_last_error()); exit(1); }
#pragma omp parallel for schedule(static)
    for (long long k = 0; k < (long long)n; k++) memcpy(s.text + s.toff[k], text.data() + (uint64_t)k * stride, tlen[k]);
    memcpy(s.query, reads.data(), s.qoff[n]);
}

// ---- cross-process / cross-thread barrier ------------------------------------------------------------------------------
struct Shared {
    pthread_barrier_t bar;
    double t_end[64];
    double val[64][8];
};
static Shared *g_sh = nullptr;
static void barrier() { pthread_barrier_wait(&g_sh->bar); }

// max over ranks of (my_end - common start); every rank gets the same number
static double span(int rank, int world, double t0_common_ignored, double my_elapsed)
{
    (void)t0_common_ignored;
    g_sh->t_end[rank] = my_elapsed;
    barrier();
    double m = 0;
    for (int r = 0; r < world; r++) m = std::max(m, g_sh->t_end[r]);
    barrier();
    return m;
}

#define CU(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

// ---- ceilings: every rank drives ONE GPU; called by all ranks at once ---------------------------------------------------
static void ceilings(int rank, int world, int dev, const Shard &s, int threads, int steps)
{
    CU(cudaSetDevice(dev));
    char *d_buf = nullptr;
    const uint64_t bytes = s.bytes();
    CU(cudaMalloc((void **)&d_buf, bytes + 64));
    cudaStream_t st;
    CU(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
    uint32_t *packed = (uint32_t *)sg_host_alloc(sg_packed_words(bytes) * 4 + 256);
    auto h2d = [&](double frac) {
        const uint64_t a = (uint64_t)(s.toff[s.n] * frac), b = (uint64_t)(s.qoff[s.n] * frac);
        CU(cudaMemcpyAsync(d_buf, s.text, a, cudaMemcpyHostToDevice, st));
        CU(cudaMemcpyAsync(d_buf + a, s.query, b, cudaMemcpyHostToDevice, st));
        return a + b;
    };
    auto pack = [&](double from) {
        const uint64_t a0 = (uint64_t)(s.toff[s.n] * from) & ~63ull, b0 = (uint64_t)(s.qoff[s.n] * from) & ~63ull;
        sg_host_pack_2bit(s.text + a0, s.toff[s.n] - a0, packed, threads);
        sg_host_pack_2bit(s.query + b0, s.qoff[s.n] - b0, packed + sg_packed_words(s.toff[s.n]), threads);
        return (s.toff[s.n] - a0) + (s.qoff[s.n] - b0);
    };
    h2d(1.0); CU(cudaStreamSynchronize(st)); pack(0.0);   // warm-up
    double r_h2d = 0, r_pack = 0, r_both_h2d = 0, r_both_pack = 0;
    {
        barrier();
        const double t0 = now_s();
        uint64_t moved = 0;
        for (int k = 0; k < steps; k++) moved += h2d(1.0);
        CU(cudaStreamSynchronize(st));
        const double el = span(rank, world, t0, now_s() - t0);
        r_h2d = moved / el;
    }
    if (threads > 0) {
        barrier();
        const double t0 = now_s();
        uint64_t done = 0;
        for (int k = 0; k < steps; k++) done += pack(0.0);
        const double el = span(rank, world, t0, now_s() - t0);
        r_pack = done / el;
    }
    if (threads > 0) {   // both at once: the copy engine takes the first half, the packers the second
        barrier();
        const double t0 = now_s();
        uint64_t moved = 0, done = 0;
        double t_h = 0, t_p = 0;
        for (int k = 0; k < steps; k++) {
            moved += h2d(0.5);
            done += pack(0.5);
            t_p = now_s() - t0;
        }
        CU(cudaStreamSynchronize(st));
        t_h = now_s() - t0;
        const double el_h = span(rank, world, t0, t_h), el_p = span(rank, world, t0, t_p);
        r_both_h2d = moved / el_h;
        r_both_pack = done / el_p;
    }
    g_sh->val[rank][0] = r_h2d; g_sh->val[rank][1] = r_pack; g_sh->val[rank][2] = r_both_h2d; g_sh->val[rank][3] = r_both_pack;
    barrier();
    if (rank == 0) {
        double s0 = 0, s1 = 0, s2 = 0, s3 = 0;
        for (int r = 0; r < world; r++) { s0 += g_sh->val[r][0]; s1 += g_sh->val[r][1]; s2 += g_sh->val[r][2]; s3 += g_sh->val[r][3]; }
        const double bytes_per_pair = (double)bytes / (double)s.n;
        printf("{\"probe\": \"ceilings\", \"gpus\": %d, \"packer_threads_per_gpu\": %d, \"h2d_ascii_gbs\": %.2f, \"host_pack_gbs\": %.2f, "
               "\"concurrent_h2d_gbs\": %.2f, \"concurrent_pack_gbs\": %.2f, \"ceiling_gbs\": %.2f, \"ceiling_alignments_per_s\": %.0f, "
               "\"concurrent_ceiling_alignments_per_s\": %.0f, \"bytes_per_pair\": %.0f, \"per_gpu_h2d_gbs\": [",
               world, threads, s0 / 1e9, s1 / 1e9, s2 / 1e9, s3 / 1e9, (s0 + 0.75 * s1) / 1e9, (s0 + 0.75 * s1) / bytes_per_pair,
               (s2 + 0.75 * s3) / bytes_per_pair, bytes_per_pair);
        for (int r = 0; r < world; r++) printf("%s%.2f", r ? ", " : "", g_sh->val[r][0] / 1e9);
        printf("]}\n");
        fflush(stdout);
    }
    barrier();
    sg_host_free(packed);
    cudaStreamDestroy(st);
    cudaFree(d_buf);
}

static void print_stats(const sg_call_stats &S)
{
    printf("\"stats\": {\"total_ms\": %.1f, \"kernel_ms\": %.1f, \"upload_ms\": %.1f, \"pack_thread_ms\": %.1f, \"wait_ms\": %.1f, "
           "\"host_other_ms\": %.1f, \"h2d_ascii_mb\": %.1f, \"h2d_packed_mb\": %.1f, \"h2d_other_mb\": %.1f, \"d2h_mb\": %.1f, "
           "\"sub_batches\": %u, \"host_threads_per_device\": %u, \"packers_in_use\": %u}",
           S.total_ns / 1e6, S.kernel_ns / 1e6, S.upload_ns / 1e6, S.pack_thread_ns / 1e6, S.wait_ns / 1e6, S.host_other_ns / 1e6,
           S.h2d_ascii_bytes / 1e6, S.h2d_packed_bytes / 1e6, S.h2d_other_bytes / 1e6, S.d2h_bytes / 1e6, S.sub_batches,
           S.host_threads_per_device, S.packers_in_use);
}

static int threads_per_gpu_default = 1;

static void set_mode(const std::string &mode, int threads_per_gpu)
{
    unsetenv("SG_HOST_THREADS"); unsetenv("SG_DMA_DEPTH"); unsetenv("SG_AFFINITY");
    if (threads_per_gpu >= 0) setenv("SG_HOST_THREADS", std::to_string(threads_per_gpu).c_str(), 1);
    unsetenv("SG_TUNE"); unsetenv("SG_PACKERS");
    if (mode == "dma") setenv("SG_HOST_THREADS", "0", 1);
    else if (mode == "hostpack") setenv("SG_DMA_DEPTH", "0", 1);
    else if (mode == "adaptive_nobind") setenv("SG_AFFINITY", "0", 1);
    else if (mode == "adaptive_all") setenv("SG_TUNE", "0", 1);          // every packer thread, no tuning
    else if (mode == "adaptive_half") setenv("SG_PACKERS", std::to_string(std::max(1, threads_per_gpu_default / 2)).c_str(), 1);
}

// one rank = one context over `devs`; all ranks run the modes in lockstep
static void run_modes(const Args &A, int rank, int world, const std::vector<int> &devs, std::vector<Shard> &shards, int threads_per_gpu)
{
    // the rank's input: its shards concatenated (one blob pair per call, as a caller of the library would hold them)
    Shard all;
    if (shards.size() == 1) all = shards[0];
    else {
        uint64_t tb = 0, qb = 0, n = 0;
        for (auto &s : shards) { tb += s.toff[s.n]; qb += s.qoff[s.n]; n += s.n; }
        all.text = (char *)sg_host_alloc(tb + 64); all.query = (char *)sg_host_alloc(qb + 64);
        all.toff.assign(1, 0); all.qoff.assign(1, 0);
        for (auto &s : shards) {
            memcpy(all.text + all.toff.back(), s.text, s.toff[s.n]);
            memcpy(all.query + all.qoff.back(), s.query, s.qoff[s.n]);
            const uint64_t t0 = all.toff.back(), q0 = all.qoff.back();
            for (uint64_t k = 1; k <= s.n; k++) { all.toff.push_back(t0 + s.toff[k]); all.qoff.push_back(q0 + s.qoff[k]); }
        }
        all.n = n;
    }
    std::stringstream ms(A.modes);
    std::string mode;
    std::vector<int64_t> ref_edit;
    while (std::getline(ms, mode, ',')) {
        set_mode(mode, threads_per_gpu);
        sg_ctx *ctx = nullptr;
        if (sg_ctx_create(&ctx, devs.data(), (int)devs.size(), 64)) { fprintf(stderr, "ctx: %s\n", sg_last_error()); exit(1); }
        sg_result *res = nullptr;
        for (int warm = 0; warm < 2; warm++) {   // two untimed calls: buffers sized, the ingest tuner has seen its three settings
            if (sg_align_pairs(ctx, all.text, all.toff.data(), all.query, all.qoff.data(), all.n, 0, &res)) { fprintf(stderr, "align: %s\n", sg_last_error()); exit(1); }
            sg_result_free(res);
        }
        barrier();
        const double t0 = now_s();
        sg_call_stats S{};
        for (int k = 0; k < A.steps; k++) {
            if (sg_align_pairs(ctx, all.text, all.toff.data(), all.query, all.qoff.data(), all.n, 0, &res)) { fprintf(stderr, "align: %s\n", sg_last_error()); exit(1); }
            sg_result_stats(res, &S);
            if (k + 1 < A.steps) sg_result_free(res);
        }
        const double el = span(rank, world, t0, now_s() - t0);
        // results must not depend on the ingest policy
        const int64_t *ed = sg_result_edit_distances(res);
        if (ref_edit.empty()) ref_edit.assign(ed, ed + all.n);
        else if (memcmp(ref_edit.data(), ed, all.n * 8)) { fprintf(stderr, "MISMATCH between ingest modes\n"); exit(2); }
        sg_result_free(res);
        double el_r = 0;
        uint64_t text_bytes = 0;
        if (A.render) {   // the same with CIGAR text for every alignment (what the CPU arm's sprintf loop produces)
            std::vector<uint64_t> off(all.n + 1);
            std::vector<char> blob;
            barrier();
            const double t1 = now_s();
            for (int k = 0; k < A.steps; k++) {
                if (sg_align_pairs(ctx, all.text, all.toff.data(), all.query, all.qoff.data(), all.n, 0, &res)) exit(1);
                const uint64_t total = sg_result_render_all(res, nullptr, 0, off.data(), 0);
                if (blob.size() < total) blob.resize(total);
                sg_result_render_all(res, blob.data(), blob.size(), off.data(), 0);
                text_bytes = total;
                sg_result_free(res);
            }
            el_r = span(rank, world, t1, now_s() - t1);
        }
        g_sh->val[rank][0] = (double)S.h2d_ascii_bytes; g_sh->val[rank][1] = (double)S.h2d_packed_bytes;
        g_sh->val[rank][2] = (double)S.upload_ns; g_sh->val[rank][3] = (double)S.pack_thread_ns; g_sh->val[rank][4] = (double)S.wait_ns;
        barrier();
        if (rank == 0) {
            const double total_pairs = (double)all.n * world * A.steps;
            printf("{\"probe\": \"e2e\", \"mode\": \"%s\", \"layout\": \"%s\", \"gpus\": %d, \"pairs_per_call\": %llu, \"alignments_per_s\": %.0f, "
                   "\"ascii_gbs\": %.2f, \"ms_per_call\": %.1f, ",
                   mode.c_str(), A.layout.c_str(), (int)devs.size() * world, (unsigned long long)all.n, total_pairs / el,
                   (double)all.bytes() * world * A.steps / el / 1e9, el / A.steps * 1e3);
            if (A.render) printf("\"rendered_alignments_per_s\": %.0f, \"cigar_text_mb_per_call\": %.1f, ", total_pairs / el_r, text_bytes / 1e6);
            print_stats(S);
            printf(", \"per_rank\": [");
            for (int r = 0; r < world; r++)
                printf("%s{\"h2d_ascii_mb\": %.0f, \"h2d_packed_mb\": %.0f, \"upload_ms\": %.1f, \"pack_thread_ms\": %.1f, \"wait_ms\": %.1f}", r ? ", " : "",
                       g_sh->val[r][0] / 1e6, g_sh->val[r][1] / 1e6, g_sh->val[r][2] / 1e6, g_sh->val[r][3] / 1e6, g_sh->val[r][4] / 1e6);
            printf("]}\n");
            fflush(stdout);
        }
        barrier();
        sg_ctx_destroy(ctx);
    }
}

static std::string cpulist_slice(int rank, int world)
{
    cpu_set_t set;
    CPU_ZERO(&set);
    sched_getaffinity(0, sizeof set, &set);
    std::vector<int> cpus;
    for (int c = 0; c < CPU_SETSIZE; c++) if (CPU_ISSET(c, &set)) cpus.push_back(c);
    const size_t a = cpus.size() * rank / world, b = cpus.size() * (rank + 1) / world;
    std::string s;
    for (size_t k = a; k < b; k++) s += (s.empty() ? "" : ",") + std::to_string(cpus[k]);
    return s;
}

int main(int argc, char **argv)
{
    Args A;
    for (int k = 1; k < argc; k++) {
        std::string a = argv[k];
        auto val = [&]() { return k + 1 < argc ? std::string(argv[++k]) : std::string(); };
        if (a == "--gpus") A.gpus = std::stoi(val());
        else if (a == "--pairs") A.pairs = std::stoull(val());
        else if (a == "--len") A.len = std::stoi(val());
        else if (a == "--steps") A.steps = std::stoi(val());
        else if (a == "--layout") A.layout = val();
        else if (a == "--modes") A.modes = val();
        else if (a == "--no-ceilings") A.ceilings = false;
        else if (a == "--no-render") A.render = false;
        else { fprintf(stderr, "unknown option %s\n", a.c_str()); return 1; }
    }
    const bool procs = A.layout == "procs";
    const int world = procs ? A.gpus : 1;
    const int ncpu = (int)std::thread::hardware_concurrency();

    g_sh = (Shared *)mmap(nullptr, sizeof(Shared), PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    pthread_barrierattr_t ba;
    pthread_barrierattr_init(&ba);
    pthread_barrierattr_setpshared(&ba, PTHREAD_PROCESS_SHARED);

    int rank = 0;
    if (procs) {   // fork BEFORE the first CUDA call: every child gets its own CUDA context
        pthread_barrier_init(&g_sh->bar, &ba, (unsigned)world);
        for (int r = 1; r < world; r++) {
            const pid_t pid = fork();
            if (pid == 0) { rank = r; break; }
        }
        // every rank owns a disjoint slice of the CPUs (the library divides the CPUs of a PROCESS among its GPUs; several
        // processes have to be told apart by their caller: SG_CPUS)
        setenv("SG_CPUS", cpulist_slice(rank, world).c_str(), 1);
    }
    if (rank == 0) {
        printf("{\"probe\": \"host\", \"cpus\": %d, \"layout\": \"%s\", \"gpus\": %d, \"numa_nodes\": \"%s\", \"cpu_model\": \"", ncpu, A.layout.c_str(), A.gpus,
               slurp("/sys/devices/system/node/online").c_str());
        std::ifstream ci("/proc/cpuinfo");
        std::string line;
        while (std::getline(ci, line)) if (line.rfind("model name", 0) == 0) { printf("%s", line.substr(line.find(':') + 2).c_str()); break; }
        printf("\", \"mem_total\": \"%s\"}\n", slurp("/proc/meminfo").substr(0, slurp("/proc/meminfo").find('\n')).c_str());
        int nd = 0;
        cudaGetDeviceCount(&nd);
        for (int d = 0; d < nd && d < A.gpus; d++) {
            char bus[32] = {0};
            cudaDeviceGetPCIBusId(bus, sizeof bus, d);
            std::string id(bus);
            for (char &c : id) c = (char)tolower(c);
            const std::string base = "/sys/bus/pci/devices/" + id + "/";
            printf("{\"probe\": \"gpu\", \"index\": %d, \"pci\": \"%s\", \"numa_node\": \"%s\", \"local_cpulist\": \"%s\", \"link_speed\": \"%s\", \"link_width\": \"%s\"}\n",
                   d, id.c_str(), slurp(base + "numa_node").c_str(), slurp(base + "local_cpulist").c_str(), slurp(base + "current_link_speed").c_str(),
                   slurp(base + "current_link_width").c_str());
        }
        fflush(stdout);
    }

    const int threads_per_gpu = std::max(1, ncpu / A.gpus - (ncpu / A.gpus >= 4 ? 1 : 0));
    threads_per_gpu_default = threads_per_gpu;
    if (procs) {
        std::vector<Shard> sh(1);
        make_shard(sh[0], (uint64_t)rank * A.pairs, A.pairs, A.len, ncpu);
        std::vector<int> devs{rank};
        if (A.ceilings) ceilings(rank, world, rank, sh[0], threads_per_gpu, A.steps);
        run_modes(A, rank, world, devs, sh, -1);
        if (rank == 0) { int st; while (wait(&st) > 0) {} }
        return 0;
    }
    // threads layout: one process; the ceilings use one thread per GPU, the e2e modes ONE context over all GPUs
    std::vector<Shard> sh(A.gpus);
    for (int d = 0; d < A.gpus; d++) make_shard(sh[d], (uint64_t)d * A.pairs, A.pairs, A.len, ncpu);
    if (A.ceilings) {
        pthread_barrier_init(&g_sh->bar, &ba, (unsigned)A.gpus);
        std::vector<std::thread> th;
        for (int d = 0; d < A.gpus; d++) th.emplace_back([&, d]() { ceilings(d, A.gpus, d, sh[d], threads_per_gpu, A.steps); });
        for (auto &t : th) t.join();
        pthread_barrier_destroy(&g_sh->bar);
    }
    pthread_barrier_init(&g_sh->bar, &ba, 1u);
    std::vector<int> devs;
    for (int d = 0; d < A.gpus; d++) devs.push_back(d);
    run_modes(A, 0, 1, devs, sh, -1);
    return 0;
}
