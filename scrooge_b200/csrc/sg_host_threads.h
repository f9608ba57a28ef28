// sg_host_threads.h -- the host threads of a context: which CPUs belong to which GPU, and a persistent pool of packer
// threads per GPU.
//
// Why: end to end the path is bound by the HOST (ASCII ingest: 20 KB per 10 kbp pair against 5 ms of GPU time per
// thousand pairs), and on an 8-GPU box eight pipelines share one host.  Round 1 opened an OpenMP team per blob and per
// sub-batch from whatever thread made the call (teams spin after their region, nothing was pinned, and pinned staging
// memory landed wherever its first touch happened).  Now every GPU of a context owns
//   * a CPU set: the CPUs the process may run on (sched_getaffinity, or SG_CPUS=<list>), restricted to the GPU's
//     NUMA-local CPUs when the machine reports them (/sys/bus/pci/devices/<bdf>/local_cpulist) and divided among the
//     GPUs that share them -- disjoint sets, so eight pipelines do not migrate over each other's caches;
//   * a pool of packer threads bound to that set, created once and parked on a condition variable between jobs;
//   * staging buffers allocated (first-touched) by a thread of that set.
#pragma once
#include <sched.h>
#include <pthread.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <thread>
#include <mutex>
#include <condition_variable>
#include <functional>
#include <algorithm>
#include <cstdint>

namespace sg {

// ---- run slab layout ---------------------------------------------------------------------------------------------------
// Alignment k of a sub-batch owns slab bytes [off[k], off[k+1]) for its CIGAR runs, at least 2*|query|+8 of them (the
// reference reserves 2*|query| entries, src/genasm_gpu.cu:995-1001).  When the kernel stores runs as whole words
// (SG_FLAG_RUN_WORDS) every offset is a multiple of 4.
//   slab_capacity      one alignment's slot when the offsets are accumulated one by one (separate strings, candidates)
//   slab_offset_blob   off[k] when the queries are one blob: q_prefix = bases of queries 0..k-1.  A difference of offsets,
//                      so all n+1 of them are computed in parallel; with 12 instead of 8 bytes of slack per alignment each
//                      offset can be rounded up on its own and no slot drops under 2*|query|+8.
inline uint64_t slab_capacity(uint64_t query_len, bool words)
{
    const uint64_t c = 2ull * query_len + 8ull;
    return words ? (c + 3ull) & ~3ull : c;
}
inline uint64_t slab_offset_blob(uint64_t q_prefix, uint64_t k, bool words)
{
    return words ? (2ull * q_prefix + 12ull * k + 3ull) & ~3ull : 2ull * q_prefix + 8ull * k;
}

// "0-3,8,10-11" -> sorted CPU numbers; empty on a malformed list
inline std::vector<int> parse_cpulist(const char *s)
{
    std::vector<int> out;
    if (!s) return out;
    while (*s) {
        while (*s == ' ' || *s == ',' || *s == '\n') s++;
        if (!*s) break;
        char *e = nullptr;
        long a = std::strtol(s, &e, 10);
        if (e == s || a < 0) return {};
        long b = a;
        if (*e == '-') {
            s = e + 1;
            b = std::strtol(s, &e, 10);
            if (e == s || b < a) return {};
        }
        for (long c = a; c <= b && c < CPU_SETSIZE; c++) out.push_back((int)c);
        s = e;
    }
    std::sort(out.begin(), out.end());
    out.erase(std::unique(out.begin(), out.end()), out.end());
    return out;
}

inline std::vector<int> allowed_cpus()
{
    if (const char *v = std::getenv("SG_CPUS")) {
        std::vector<int> l = parse_cpulist(v);
        if (!l.empty()) return l;
    }
    std::vector<int> out;
    cpu_set_t set;
    CPU_ZERO(&set);
    if (sched_getaffinity(0, sizeof set, &set) == 0)
        for (int c = 0; c < CPU_SETSIZE; c++)
            if (CPU_ISSET(c, &set)) out.push_back(c);
    if (out.empty())
        for (unsigned c = 0; c < std::max(1u, std::thread::hardware_concurrency()); c++) out.push_back((int)c);
    return out;
}

// CPUs local to the PCI device "0000:1b:00.0" (as cudaDeviceGetPCIBusId prints it); empty when unknown
inline std::vector<int> pci_local_cpus(const char *bus_id)
{
    std::string id(bus_id ? bus_id : "");
    for (char &c : id) c = (char)std::tolower((unsigned char)c);
    const std::string path = "/sys/bus/pci/devices/" + id + "/local_cpulist";
    FILE *f = std::fopen(path.c_str(), "r");
    if (!f) return {};
    char buf[4096];
    const size_t n = std::fread(buf, 1, sizeof buf - 1, f);
    std::fclose(f);
    buf[n] = 0;
    return parse_cpulist(buf);
}

// Divides `allowed` among n_dev devices.  local[k] = the CPUs the machine reports as close to device k (may be empty
// or all of them).  Devices with the same effective local set share it in equal contiguous slices; a device whose
// local set does not intersect `allowed` falls back to a slice of everything.
inline std::vector<std::vector<int>> assign_cpus(const std::vector<int> &allowed, const std::vector<std::vector<int>> &local)
{
    const int nd = (int)local.size();
    std::vector<std::vector<int>> eff(nd), out(nd);
    for (int k = 0; k < nd; k++) {
        for (int c : local[k])
            if (std::binary_search(allowed.begin(), allowed.end(), c)) eff[k].push_back(c);
        if (eff[k].empty()) eff[k] = allowed;
    }
    for (int k = 0; k < nd; k++) {
        int same = 0, rank = 0;   // devices with the same set, and this one's position among them
        for (int j = 0; j < nd; j++)
            if (eff[j] == eff[k]) { if (j < k) rank++; same++; }
        const size_t n = eff[k].size();
        const size_t a = n * (size_t)rank / (size_t)same, b = n * (size_t)(rank + 1) / (size_t)same;
        if (b > a) out[k].assign(eff[k].begin() + (long)a, eff[k].begin() + (long)b);
        else out[k] = eff[k];   // more devices than CPUs: share
    }
    return out;
}

inline void bind_this_thread(const std::vector<int> &cpus)
{
    if (cpus.empty()) return;
    cpu_set_t set;
    CPU_ZERO(&set);
    for (int c : cpus)
        if (c >= 0 && c < CPU_SETSIZE) CPU_SET(c, &set);
    pthread_setaffinity_np(pthread_self(), sizeof set, &set);   // best effort: a refusal leaves the thread where it was
}

// Restores the calling thread's affinity when it goes out of scope (public entry points borrow the caller's thread).
struct ScopedAffinity {
    cpu_set_t saved;
    bool have = false;
    explicit ScopedAffinity(const std::vector<int> &cpus)
    {
        if (cpus.empty()) return;
        have = pthread_getaffinity_np(pthread_self(), sizeof saved, &saved) == 0;
        bind_this_thread(cpus);
    }
    ~ScopedAffinity() { if (have) pthread_setaffinity_np(pthread_self(), sizeof saved, &saved); }
};

// A fixed team of threads that runs one job at a time: run(fn) calls fn(tid) on every thread and returns when all are
// done.  Threads sleep between jobs (no spinning: the cores belong to whoever has work).
class ThreadTeam {
public:
    ThreadTeam() = default;
    ThreadTeam(const ThreadTeam &) = delete;
    ThreadTeam &operator=(const ThreadTeam &) = delete;
    ~ThreadTeam() { stop(); }

    int size() const { return (int)threads_.size(); }

    // init(tid) runs once on each new thread (device binding etc.)
    void start(int n, const std::vector<int> &cpus, std::function<void(int)> init)
    {
        stop();
        quit_ = false;
        generation_ = 0;
        pending_ = 0;
        for (int t = 0; t < n; t++)
            threads_.emplace_back([this, t, cpus, init]() {
                bind_this_thread(cpus);
                if (init) init(t);
                uint64_t seen = 0;
                while (true) {
                    std::function<void(int)> *job;
                    {
                        std::unique_lock<std::mutex> lk(mu_);
                        cv_.wait(lk, [&] { return quit_ || generation_ != seen; });
                        if (quit_) return;
                        seen = generation_;
                        job = job_;
                    }
                    (*job)(t);
                    {
                        std::lock_guard<std::mutex> lk(mu_);
                        if (--pending_ == 0) done_.notify_all();
                    }
                }
            });
    }

    void stop()
    {
        {
            std::lock_guard<std::mutex> lk(mu_);
            quit_ = true;
        }
        cv_.notify_all();
        for (auto &t : threads_) t.join();
        threads_.clear();
    }

    // Starts fn on every thread; the caller may do its own share of the work and then calls wait().
    void launch(std::function<void(int)> &fn)
    {
        std::lock_guard<std::mutex> lk(mu_);
        job_ = &fn;
        pending_ = (int)threads_.size();
        generation_++;
        cv_.notify_all();
    }
    void wait()
    {
        std::unique_lock<std::mutex> lk(mu_);
        done_.wait(lk, [&] { return pending_ == 0; });
    }

private:
    std::vector<std::thread> threads_;
    std::mutex mu_;
    std::condition_variable cv_, done_;
    std::function<void(int)> *job_ = nullptr;
    uint64_t generation_ = 0;
    int pending_ = 0;
    bool quit_ = false;
};

// How many of a GPU's packer threads the adaptive ingest uses.  Copy engines and packers read the same host DRAM: on a
// box where PCIe is the narrow link (one GPU, many cores) every packer helps, on a box where DRAM is (eight GPUs pulling
// ASCII over eight links) a packed byte costs 1.5 bytes of DRAM traffic against 1.0 for a copied one and the packers
// only take bandwidth from the copy engines.  Instead of a rule in terms of core counts the context measures: the first
// large sub-batches run with all, half and none of the packers, the ingest rate (ASCII bytes per second of upload
// phase) of each is recorded, and the best setting is kept (SG_PACKERS=<n> fixes it, SG_TUNE=0 keeps all).
struct IngestTuner {
    static constexpr uint64_t kMinBytes = 96ull << 20;   // smaller sub-batches say little about a rate
    int fixed = -1;        // SG_PACKERS
    bool enabled = true;   // SG_TUNE
    int candidates[3] = {0, 0, 0};
    double rate[3] = {0, 0, 0};
    int tried = 0, best = 0, current = 0;
    void init(int threads)
    {
        candidates[0] = threads; candidates[1] = threads / 2; candidates[2] = 0;
        if (const char *v = std::getenv("SG_PACKERS")) fixed = std::max(0, std::min(threads, std::atoi(v)));
        if (const char *v = std::getenv("SG_TUNE")) enabled = std::atoi(v) != 0;
        if (threads < 2) enabled = false;
    }
    int packers(int threads, uint64_t bytes)
    {
        if (fixed >= 0) return fixed;
        if (!enabled) return threads;
        current = (tried < 3 && bytes >= kMinBytes) ? tried : best;
        return candidates[current];
    }
    void report(uint64_t bytes, double seconds)
    {
        if (fixed >= 0 || !enabled || tried >= 3 || bytes < kMinBytes || current != tried || seconds <= 0) return;
        rate[tried] = (double)bytes / seconds;
        tried++;
        if (tried == 3) {
            best = 0;
            for (int k = 1; k < 3; k++) if (rate[k] > rate[best] * 1.03) best = k;   // a setting has to win clearly
        }
    }
};

}  // namespace sg
