// sg_tests.cpp -- the reference's `tests` binary (reference src/tests.cu:726-815) against this library: same options,
// same output lines.  --unit_tests runs the reference's known-answer tests on the GPU path; without it the performance
// test reads a FASTA reference, FASTQ reads and MAF/PAF seeds, aligns every (read, forward-strand candidate) and
// validates every CIGAR.  --dump_inputs parses the files and prints what was read (no GPU needed).
#include <algorithm>
#include <cstdio>
#include <iostream>
#include <string>
#include <vector>

#include "genasm_gpu.hpp"
#include "scrooge_b200.h"
#include "scrooge_b200_bench.h"
#include "scrooge_io.hpp"

using namespace std;

static bool enable_log = false;

static void print_gpu_info()
{
    cout << sg_device_count() << " visible GPU(s)" << endl << endl;
}

// reference src/tests.cu:583-647 pins its own byte layout (first base in bits 7:6); ours is 16 bases per little-endian
// 32-bit word with base k in bits 2k+1:2k (include/scrooge_b200.h).  Same test strings, our layout.
static void ascii_to_two_bit_correctness_test()
{
    const vector<pair<string, vector<uint32_t>>> cases = {
        {"", {}},
        {"A", {0x0u}},
        {"T", {0x3u}},
        {"ACGT", {0xE4u}},
        {"ACGTA", {0xE4u}},
        {"ACGTC", {0x1E4u}},
        {"acgtacgtacgtacgtacgtacgtacgtacgt", {0xE4E4E4E4u, 0xE4E4E4E4u}},
        {"ACGTACGTACGTACGTACGTACGTACGTACGTG", {0xE4E4E4E4u, 0xE4E4E4E4u, 0x2u}},
    };
    bool ok = true;
    for (const auto &c : cases) {
        vector<uint32_t> out((c.first.size() + 15) / 16 + 1, 0xFFFFFFFFu);
        const uint64_t bad = sg_host_pack_2bit(c.first.data(), c.first.size(), out.data(), 1);
        out.resize((c.first.size() + 15) / 16);
        if (bad != ~0ull || out != c.second) {
            cout << "FAILED ascii_to_two_bit_correctness_test for \"" << c.first << "\"" << endl;
            ok = false;
        }
    }
    uint32_t w[2];
    if (sg_host_pack_2bit("ACGNACGT", 8, w, 1) != 3) { cout << "FAILED ascii_to_two_bit_correctness_test: 'N' not reported" << endl; ok = false; }
    if (ok) cout << "PASSED ascii_to_two_bit_correctness_test" << endl;
}

// reference src/tests.cu:171-222
static void gpu_algorithm_correctness_test()
{
    Genome_t reference;
    reference.content = "AAAACCCCGGGGTTTT";
    CandidateLocation_t ref_begin{};
    ref_begin.start_in_reference = 0;
    ref_begin.strand = true;
    const vector<CandidateLocation_t> loc(1, ref_begin);
    vector<Read_t> reads = {
        {"test_read_4d12m4i", "CCCCGGGGTTTTAAAA", loc},      {"test_read_16m", "AAAACCCCGGGGTTTT", loc},
        {"test_read_3d7m", "ACCCCGG", loc},                   {"test_read_4m4d4m4i4m", "AAAAGGGGAAAATTTT", loc},
        {"test_read_12s4m", "AAAAAAAAAAAAAAAA", loc},         {"test_read_1m1s1i3m1s2m3i", "ATTAACGCCTTT", loc},
        {"test_read_oversized", "TTTTAAAACCCCGGGGTTTTAAAA", loc}, {"test_read_empty", "", loc},
        {"test_read_len64", "TTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTTAAAACCCCGGGGTTTTAAAA", loc},
    };
    const vector<int> correct = {8, 0, 3, 8, 12, 6, 8, 0, 48};  // reference src/tests.cu:190
    vector<Alignment_t> alignments = genasm_gpu::align_all(reference, reads);
    if (alignments.size() != correct.size()) {
        cout << "FAILED gpu_algorithm_correctness_test: align_all() produced wrong number of alignments" << endl;
        return;
    }
    bool ok = true;
    for (size_t i = 0; i < alignments.size(); i++) {
        if (alignments[i].edit_distance != correct[i]) {
            cout << "FAILED gpu_algorithm_correctness_test: align_all() produced distance " << alignments[i].edit_distance << " instead of "
                 << correct[i] << " for read \"" << reads[i].description << "\"" << endl;
            ok = false;
        }
        const string why = validate_cigar(alignments[i], ref_begin, reads[i], reference);
        if (!why.empty()) { cout << why << endl; ok = false; }
    }
    if (ok) cout << "PASSED gpu_algorithm_correctness_test" << endl;
}

// reference src/tests.cu:273-333: every entry point must return the same strings.  The CPU entry points are the
// reference's own (not part of this library); their answers for these pairs are the golden strings below, produced by
// the unmodified reference (tests/golden/golden_w64.json, group differential_tests_cu).
static void library_interface_correctness_test()
{
    struct Case { const char *query, *text, *cigar; long long ed; };
    const vector<Case> cases = {
        {"ACGT", "ACGT", "4=", 0},
        {"CAAATCTATTAAGTCAAACGGTCCGTAAGCTAGAACCTCCTGCCGTGTAAGTTACGACGTGGTCGAGTTACTTTCGTTCTTATTAACACAATGTCCATCA",
         "CAAACCTATCAAGTCAAACGGTCCGTAGCTACACCTCCTGCCGTGTAAAGTTACGACGTGGTTGAGTTACTTTCGTTCTTATTAACAACAATGTTCCATCA",
         "4=1X4=1X16=1I4=1=1I1X14=1D14=1=1X23=1D5=1=1D7=", 9},
    };
    const bool w64 = !getenv("SG_WINDOW") || string(getenv("SG_WINDOW")) == "64";
    for (const Case &c : cases) {
        Genome_t reference;
        reference.content = c.text;
        CandidateLocation_t ref_begin{};
        ref_begin.start_in_reference = 0;
        ref_begin.strand = true;
        vector<Read_t> reads(1, Read_t{"test", c.query, vector<CandidateLocation_t>(1, ref_begin)});
        vector<string> queries(1, c.query), texts(1, c.text);
        vector<Alignment_t> pairwise = genasm_gpu::align_all(texts, queries);
        vector<Alignment_t> mapping = genasm_gpu::align_all(reference, reads);
        bool same = pairwise[0].cigar == mapping[0].cigar && pairwise[0].edit_distance == mapping[0].edit_distance;
        if (w64) same = same && pairwise[0].cigar == c.cigar && pairwise[0].edit_distance == c.ed;
        if (!same) {
            cout << "FAILED library_interface_correctness_test: align_all() produced different CIGAR strings" << endl;
            return;
        }
    }
    cout << "PASSED library_interface_correctness_test" << endl;
}

static void load_inputs(const string &reference_file, const string &reads_file, const string &seeds_file, Genome_t &genome,
                        vector<Read_t> &reads)
{
    if (enable_log) cerr << "Reading reference sequence..." << endl;
    genome = read_genome(reference_file);
    if (enable_log) cerr << "Reading reads files..." << endl;
    read_fastq_and_seed_locations(genome, reads_file, seeds_file, reads);
    if (enable_log) cerr << "Filtering reads..." << endl;
    for (Read_t &read : reads)  // forward strand only, as the reference's driver (src/tests.cu:347-355)
        read.locations.erase(remove_if(read.locations.begin(), read.locations.end(), [](const CandidateLocation_t &l) { return !l.strand; }),
                             read.locations.end());
    if (enable_log) cerr << "Sorting reads..." << endl;
    stable_sort(reads.begin(), reads.end(), [](const Read_t &a, const Read_t &b) { return a.content.size() > b.content.size(); });
}

// reference src/tests.cu:335-409
static int gpu_algorithm_performance_test(const string &reference_file, const string &reads_file, const string &seeds_file)
{
    Genome_t genome;
    vector<Read_t> reads;
    load_inputs(reference_file, reads_file, seeds_file, genome, reads);
    if (enable_log) cerr << "Running alignment algorithm..." << endl;
    vector<Alignment_t> alignments;
    long long core_algorithm_ns = 0;
    const long long end_to_end_ns = measure_ns([&]() { alignments = genasm_gpu::align_all(genome, reads, &core_algorithm_ns); });
    if (enable_log) cerr << "Sanity checking alignments..." << endl;
    size_t pair_idx = 0, failures = 0;
    for (const Read_t &read : reads)
        for (const CandidateLocation_t &location : read.locations) {
            const string why = validate_cigar(alignments[pair_idx], location, read, genome);
            if (!why.empty()) {
                cout << why << endl << "FAILED sanity check in algorithm_performance_test for alignment " << pair_idx << endl;
                failures++;
            }
            pair_idx++;
        }
    cout << "align_all() took " << (end_to_end_ns / 1000000) << "ms (data transfers, conversion, gpu kernel and post-processing)" << endl;
    cout << "GPU kernel took " << (core_algorithm_ns / 1000000) << "ms" << endl;
    cout << "GPU kernel ran at " << (core_algorithm_ns ? (long long)alignments.size() * 1000000000ll / core_algorithm_ns : 0) << " aligns/second" << endl;
    cout << alignments.size() << " alignments, " << failures << " failed the sanity check" << endl;
    return failures ? 1 : 0;
}

// Synthetic end-to-end run of the drop-in C++ interface (BASELINE.json configs[2] shape): n pairs of 10 kbp reads at 10 %,
// std::vector<std::string> in, std::vector<Alignment_t> (rendered CIGAR strings) out, every CIGAR validated.
static int synthetic_performance_test(uint64_t n_pairs, uint32_t read_len)
{
    const uint64_t stride = sg_synth_text_stride(read_len, 64);
    vector<char> text(n_pairs * stride), rd(n_pairs * (uint64_t)read_len);
    vector<uint64_t> tlen(n_pairs);
    sg_synth_pairs_host(0x5C2006E + 3, 0, n_pairs, read_len, 0.10, 6, 50, 54, 64, text.data(), stride, tlen.data(), rd.data());
    vector<string> texts(n_pairs), queries(n_pairs);
    for (uint64_t p = 0; p < n_pairs; p++) {
        texts[p].assign(text.data() + p * stride, tlen[p]);
        queries[p].assign(rd.data() + p * read_len, read_len);
    }
    vector<char>().swap(text);
    vector<char>().swap(rd);
    vector<Alignment_t> alignments;
    long long core_ns = 0, best_ns = 0;
    for (int rep = 0; rep < 3; rep++) {  // the first call also creates the context and sizes its buffers
        const long long ns = measure_ns([&]() { alignments = genasm_gpu::align_all(texts, queries, &core_ns); });
        if (rep == 0 || ns < best_ns) best_ns = ns;
        cout << "align_all() took " << (ns / 1000000) << "ms (data transfers, conversion, gpu kernel and post-processing), GPU kernel "
             << (core_ns / 1000000) << "ms" << endl;
    }
    size_t failures = 0;
    Genome_t g;
    for (uint64_t p = 0; p < n_pairs; p += 97) {
        g.content = texts[p];
        CandidateLocation_t loc{};
        Read_t r{"", queries[p], {}};
        if (!validate_cigar(alignments[p], loc, r, g).empty()) failures++;
    }
    cout << "end to end " << (long long)((double)n_pairs * 1e9 / (double)best_ns) << " aligns/second, GPU kernel "
         << (long long)((double)n_pairs * 1e9 / (double)core_ns) << " aligns/second, " << failures << " failed the sanity check" << endl;
    return failures ? 1 : 0;
}

static void dump_inputs(const string &reference_file, const string &reads_file, const string &seeds_file)
{
    Genome_t genome;
    vector<Read_t> reads;
    load_inputs(reference_file, reads_file, seeds_file, genome, reads);
    cout << "genome " << genome.content.size() << " bases, " << genome.chromosome_starts.size() << " chromosome(s)" << endl;
    for (const auto &kv : genome.chromosome_starts) cout << "chromosome \"" << kv.first << "\" starts at " << kv.second << endl;
    for (const Read_t &r : reads) {
        cout << "read \"" << r.description << "\" " << r.content.size() << " bases:";
        for (const CandidateLocation_t &l : r.locations) cout << " " << l.chromosome << "@" << l.start_in_chromosome << "->" << l.start_in_reference;
        cout << endl;
    }
}

int main(int argc, char **argv)
{
    string reference_file = "datasets/human_genome/pacbio-chr1-simulated-m10k-k5_0001.ref";
    string reads_file = "datasets/human_genome/pacbio-chr1-simulated-m10k-k5_0001.fastq";
    string seeds_file = "datasets/human_genome/pacbio-chr1-simulated-m10k-k5_0001.maf";
    const bool gpu_info_only = OPT_EXISTS == get_cmd_option(argc, argv, "--gpu_info_only");
    const bool verbose = OPT_EXISTS == get_cmd_option(argc, argv, "--verbose");
    const bool unit_tests = OPT_EXISTS == get_cmd_option(argc, argv, "--unit_tests");
    const bool dump = OPT_EXISTS == get_cmd_option(argc, argv, "--dump_inputs");
    string synthetic;
    const int has_synth = get_cmd_option(argc, argv, "--synthetic", synthetic);
    bool help = false;
    // the window configuration is a run-time choice here (the reference needs a rebuild with -DCLI_W/-DCLI_K/-DCLI_O,
    // src/genasm_gpu.cu:1-63): --window / --overlap set what the drop-in reads from SG_WINDOW / SG_OVERLAP
    string window, overlap;
    const int has_w = get_cmd_option(argc, argv, "--window", window), has_o = get_cmd_option(argc, argv, "--overlap", overlap);
    help |= has_w == OPT_INVALID || has_o == OPT_INVALID;
    if (has_w == OPT_EXISTS) setenv("SG_WINDOW", window.c_str(), 1);
    if (has_o == OPT_EXISTS) setenv("SG_OVERLAP", overlap.c_str(), 1);
    help |= OPT_INVALID == get_cmd_option(argc, argv, "--reference", reference_file);
    help |= OPT_INVALID == get_cmd_option(argc, argv, "--reads", reads_file);
    help |= OPT_INVALID == get_cmd_option(argc, argv, "--seeds", seeds_file);
    for (const char *flag : {"--gpu_info_only", "--verbose", "--unit_tests", "--dump_inputs"}) help |= OPT_INVALID == get_cmd_option(argc, argv, flag);
    help |= OPT_MISSING != get_cmd_option(argc, argv, "--help");
    help |= has_synth == OPT_INVALID;
    help |= !check_options(argc, argv, {"--reference", "--reads", "--seeds", "--help", "--gpu_info_only", "--verbose", "--unit_tests", "--dump_inputs", "--synthetic", "--window", "--overlap"});
    if (help) {
        cout << "sg_tests [options]\n"
                "Options:\n"
                "--reference=[path to reference FASTA] -- reference data for the performance test\n"
                "--reads=[path to reads FASTQ]         -- reads data for the performance test\n"
                "--seeds=[path to MAF or PAF]          -- seeds data for the performance test\n"
                "--gpu_info_only                       -- only print GPU info\n"
                "--verbose                             -- print progress to stderr\n"
                "--unit_tests                          -- run unit tests (default: performance test)\n"
                "--dump_inputs                         -- parse the input files and print them (no GPU needed)\n"
                "--synthetic=[pairs]                   -- end-to-end run of align_all() on synthetic 10 kbp / 10 % pairs\n"
                "--window=[W] --overlap=[O]            -- window configuration (default 64 / min(W/2+1, W-1); W <= 256, W-O <= 128)\n"
                "--help                                -- displays this information\n";
        return 0;
    }
    if (gpu_info_only) { print_gpu_info(); return 0; }
    genasm_gpu::enabled_algorithm_log = verbose;
    enable_log = verbose;
    try {
        if (dump) { dump_inputs(reference_file, reads_file, seeds_file); return 0; }
        print_gpu_info();
        if (has_synth == OPT_EXISTS) return synthetic_performance_test(strtoull(synthetic.c_str(), nullptr, 10), 10000);
        if (unit_tests) {
            ascii_to_two_bit_correctness_test();
            gpu_algorithm_correctness_test();
            library_interface_correctness_test();
            return 0;
        }
        return gpu_algorithm_performance_test(reference_file, reads_file, seeds_file);
    } catch (const exception &e) {
        cout << "FAILED: " << e.what() << endl;
        return 1;
    }
}
