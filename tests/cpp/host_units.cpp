// host_units.cpp -- unit tests of the CUDA-free host helpers (csrc/sg_host_threads.h): CPU lists, the division of CPUs
// among GPUs, the packer team, the ingest tuner.  Built and run by tests/test_host_units.py with plain g++ (no GPU).
#include <atomic>
#include <cassert>
#include <cstdio>
#include <numeric>
#include <set>

#include "../../scrooge_b200/csrc/sg_host_threads.h"

#define CHECK(x) do { if (!(x)) { std::printf("FAILED %s:%d: %s\n", __FILE__, __LINE__, #x); return 1; } } while (0)

using namespace sg;

int main()
{
    // ---- CPU lists
    CHECK((parse_cpulist("0-3,8,10-11\n") == std::vector<int>{0, 1, 2, 3, 8, 10, 11}));
    CHECK((parse_cpulist("5") == std::vector<int>{5}));
    CHECK((parse_cpulist("3,1,1-2") == std::vector<int>{1, 2, 3}));
    CHECK(parse_cpulist("").empty() && parse_cpulist("a-b").empty() && parse_cpulist("4-2").empty() && parse_cpulist(nullptr).empty());

    // ---- CPUs per GPU: unknown locality -> equal disjoint slices of what the process may use
    std::vector<int> all(32);
    std::iota(all.begin(), all.end(), 0);
    {
        auto sets = assign_cpus(all, std::vector<std::vector<int>>(8));
        std::set<int> seen;
        for (auto &s : sets) { CHECK(s.size() == 4); for (int c : s) CHECK(seen.insert(c).second); }
        CHECK(seen.size() == 32);
    }
    {   // two NUMA nodes with four GPUs each: every GPU stays on its node, nobody shares a CPU
        std::vector<int> n0(all.begin(), all.begin() + 16), n1(all.begin() + 16, all.end());
        std::vector<std::vector<int>> local{n0, n0, n0, n0, n1, n1, n1, n1};
        auto sets = assign_cpus(all, local);
        std::set<int> seen;
        for (int k = 0; k < 8; k++) {
            CHECK(sets[k].size() == 4);
            for (int c : sets[k]) { CHECK((k < 4) == (c < 16)); CHECK(seen.insert(c).second); }
        }
    }
    {   // the process's mask wins; a GPU whose node lies outside it gets a slice of the mask; more GPUs than CPUs share
        auto sets = assign_cpus({0, 1, 2, 3}, {{0, 1, 2, 3, 4, 5}, {16, 17}});
        for (auto &s : sets) { CHECK(!s.empty()); for (int c : s) CHECK(c >= 0 && c <= 3); }
        auto many = assign_cpus({0, 1}, std::vector<std::vector<int>>(4));
        for (auto &s : many) CHECK(!s.empty());
    }

    // ---- the team: every thread runs every job exactly once, jobs do not overlap, threads are reused
    {
        ThreadTeam team;
        std::atomic<int> inits{0};
        team.start(6, {}, [&](int) { inits++; });
        CHECK(team.size() == 6);
        for (int round = 0; round < 200; round++) {
            std::atomic<int> hits{0};
            std::atomic<unsigned> mask{0};
            std::function<void(int)> job = [&](int tid) { hits++; mask |= 1u << tid; };
            team.launch(job);
            team.wait();
            CHECK(hits.load() == 6 && mask.load() == 0x3Fu);
        }
        CHECK(inits.load() == 6);
        // chunks handed out from a shared counter: everything is processed once whatever the interleaving
        std::atomic<long> next{0}, sum{0};
        std::function<void(int)> job = [&](int) { for (long c; (c = next.fetch_add(1)) < 100000;) sum += c; };
        team.launch(job);
        team.wait();
        CHECK(sum.load() == 100000L * 99999L / 2);
        team.stop();
        CHECK(team.size() == 0);
        team.start(2, {0}, nullptr);   // restart with another size and a binding
        std::atomic<int> h2{0};
        std::function<void(int)> j2 = [&](int) { h2++; };
        team.launch(j2);
        team.wait();
        CHECK(h2.load() == 2);
    }

    // ---- the ingest tuner: tries all / half / none on the first large sub-batches, keeps a setting only if it wins clearly
    {
        const uint64_t big = 512ull << 20, small = 8ull << 20;
        unsetenv("SG_PACKERS"); unsetenv("SG_TUNE");
        IngestTuner t;
        t.init(15);
        CHECK(t.packers(15, small) == 15);          // a small sub-batch does not start the experiment
        t.report(small, 0.001);
        CHECK(t.packers(15, big) == 15); t.report(big, big / 120e9);
        CHECK(t.packers(15, big) == 7);  t.report(big, big / 100e9);
        CHECK(t.packers(15, big) == 0);  t.report(big, big / 50e9);
        CHECK(t.packers(15, big) == 15 && t.packers(15, small) == 15);   // one GPU, many cores: all packers
        IngestTuner u;                               // a host whose DRAM is the limit: copies alone are best
        u.init(3);
        CHECK(u.packers(3, big) == 3); u.report(big, big / 19e9);
        CHECK(u.packers(3, big) == 1); u.report(big, big / 20e9);
        CHECK(u.packers(3, big) == 0); u.report(big, big / 23e9);
        CHECK(u.packers(3, big) == 0);
        IngestTuner v;                               // within 3 %: stay with all
        v.init(8);
        v.packers(8, big); v.report(big, 1.00);
        v.packers(8, big); v.report(big, 0.99);
        v.packers(8, big); v.report(big, 0.98);
        CHECK(v.packers(8, big) == 8);
        setenv("SG_PACKERS", "2", 1);
        IngestTuner f;
        f.init(8);
        CHECK(f.packers(8, big) == 2 && f.packers(8, small) == 2);
        unsetenv("SG_PACKERS");
        setenv("SG_TUNE", "0", 1);
        IngestTuner o;
        o.init(8);
        CHECK(o.packers(8, big) == 8);
        unsetenv("SG_TUNE");
    }
    // ---- run slab layout: slots of at least 2*|query|+8 bytes, on 4-byte boundaries when runs are stored as words
    {
        uint64_t lens[] = {0, 1, 2, 3, 150, 151, 1000, 9999, 10000, 7, 0, 0, 5, 100001};
        for (int words = 0; words < 2; words++) {
            uint64_t prefix = 0, serial = 0;
            for (uint64_t k = 0; k < sizeof lens / sizeof *lens; k++) {
                const uint64_t o0 = slab_offset_blob(prefix, k, words), o1 = slab_offset_blob(prefix + lens[k], k + 1, words);
                CHECK(o1 - o0 >= 2 * lens[k] + 8 && o1 - o0 <= 2 * lens[k] + 16);
                CHECK(!words || (o0 % 4 == 0 && o1 % 4 == 0));
                const uint64_t c = slab_capacity(lens[k], words);
                CHECK(c >= 2 * lens[k] + 8 && c <= 2 * lens[k] + 11 && (!words || c % 4 == 0) && (words || c == 2 * lens[k] + 8));
                CHECK(!words || serial % 4 == 0);
                serial += c;
                prefix += lens[k];
            }
            CHECK(slab_offset_blob(0, 0, words) == 0);
        }
    }
    std::printf("host units ok\n");
    return 0;
}
