// sg_io.cpp -- dataset readers and option parsing declared in include/scrooge_io.hpp.  Host-only C++.
#include <algorithm>
#include <cctype>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <map>
#include <stdexcept>
#include <unordered_map>

#include "../../include/scrooge_io.hpp"

namespace {

// a cursor over the file image with the few scanning primitives the formats need
struct Scanner {
    const std::string &s;
    size_t p = 0;
    explicit Scanner(const std::string &str) : s(str) {}
    bool done() const { return p >= s.size(); }
    // the current line without its terminator; moves to the next line
    std::string line()
    {
        size_t e = s.find('\n', p);
        size_t stop = e == std::string::npos ? s.size() : e;
        size_t q = stop;
        while (q > p && (s[q - 1] == '\r' || s[q - 1] == '\n')) q--;
        std::string out = s.substr(p, q - p);
        p = e == std::string::npos ? s.size() : e + 1;
        return out;
    }
    // skip forward to just after the next `c`; false when there is none
    bool seek_after(char c)
    {
        size_t e = s.find(c, p);
        if (e == std::string::npos) { p = s.size(); return false; }
        p = e + 1;
        return true;
    }
};

std::vector<std::string> fields(const std::string &line, bool tabs_only)
{
    std::vector<std::string> out;
    size_t i = 0;
    while (i <= line.size()) {
        if (tabs_only) {
            size_t e = line.find('\t', i);
            if (e == std::string::npos) e = line.size();
            out.push_back(line.substr(i, e - i));
            i = e + 1;
        } else {
            while (i < line.size() && std::isspace((unsigned char)line[i])) i++;
            if (i >= line.size()) break;
            size_t e = i;
            while (e < line.size() && !std::isspace((unsigned char)line[e])) e++;
            out.push_back(line.substr(i, e - i));
            i = e;
        }
    }
    return out;
}

long long to_ll(const std::string &s) { return s.empty() ? 0 : std::strtoll(s.c_str(), nullptr, 10); }

}  // namespace

void remove_whitespaces(std::string &str)
{
    str.erase(std::remove_if(str.begin(), str.end(), [](unsigned char c) { return std::isspace(c); }), str.end());
}

std::string read_file(std::string file_path)
{
    std::ifstream f(file_path, std::ios::binary | std::ios::ate);
    if (!f.is_open()) throw std::runtime_error("could not read file \"" + file_path + "\"");
    const std::streamsize size = f.tellg();
    std::string out((size_t)size, '\0');
    f.seekg(0);
    f.read(&out[0], size);
    return out;
}

std::vector<Sequence_t> read_fasta(std::string file_path)
{
    const std::string img = read_file(file_path);
    Scanner sc(img);
    std::vector<Sequence_t> out;
    if (!sc.seek_after('>')) return out;
    while (!sc.done()) {
        Sequence_t seq;
        // header: up to the end of the line
        size_t e = img.find_first_of("\r\n", sc.p);
        if (e == std::string::npos) e = img.size();
        seq.description = img.substr(sc.p, e - sc.p);
        sc.p = e;
        // body: everything up to the next '>' minus line breaks and blanks
        size_t next = img.find('>', sc.p);
        if (next == std::string::npos) next = img.size();
        seq.content.reserve(next - sc.p);
        for (size_t i = sc.p; i < next; i++) {
            const char c = img[i];
            if (c != '\n' && c != '\r' && c != ' ') seq.content.push_back(c);
        }
        out.push_back(std::move(seq));
        sc.p = next < img.size() ? next + 1 : img.size();
    }
    return out;
}

Genome_t read_genome(std::string fasta_file_path)
{
    std::vector<Sequence_t> chromosomes = read_fasta(fasta_file_path);
    Genome_t genome;
    size_t total = 0;
    for (const Sequence_t &c : chromosomes) total += c.content.size();
    genome.content.reserve(total);
    for (const Sequence_t &c : chromosomes) {
        genome.chromosome_starts[c.description] = (long long)genome.content.size();
        genome.content += c.content;
    }
    return genome;
}

std::vector<Read_t> read_fastq(std::string file_path)
{
    const std::string img = read_file(file_path);
    Scanner sc(img);
    std::vector<Read_t> reads;
    while (sc.seek_after('@')) {
        Read_t r;
        std::string header = sc.line();
        for (char c : header)
            if (c != '\r' && c != ' ') r.description.push_back(c);
        r.content = sc.line();
        reads.push_back(std::move(r));
        // '+' line and qualities are skipped by the search for the next '@'; a quality line that starts with '@'
        // would be misread exactly as in the reference (src/util.cpp:119-155)
    }
    return reads;
}

std::vector<CandidateLocation_t> read_maf(std::string file_path)
{
    const std::string img = read_file(file_path);
    Scanner sc(img);
    std::vector<CandidateLocation_t> out;
    while (!sc.done()) {
        std::string l = sc.line();
        if (l.empty() || l[0] != 'a') continue;
        CandidateLocation_t loc{};
        while (!sc.done()) {  // the block ends at the first empty line
            l = sc.line();
            if (l.empty()) break;
            if (l[0] != 's') continue;
            const std::vector<std::string> f = fields(l.substr(1), false);  // src start size strand srcSize text
            if (f.size() < 5) continue;
            if (f[0] == "ref") {
                loc.start_in_chromosome = to_ll(f[1]);
                loc.chromosome = "ref";
            } else {
                loc.read_description = f[0];
                loc.start_of_aligned_region = to_ll(f[1]);
                loc.size_of_aligned_region = to_ll(f[2]);
                loc.strand = f[3] == "+";
            }
        }
        out.push_back(std::move(loc));
    }
    return out;
}

std::vector<CandidateLocation_t> read_paf(std::string file_path)
{
    const std::string img = read_file(file_path);
    Scanner sc(img);
    std::vector<CandidateLocation_t> out;
    while (!sc.done()) {
        const std::string l = sc.line();
        if (l.empty()) continue;
        const std::vector<std::string> f = fields(l, true);
        if (f.size() < 9) continue;
        CandidateLocation_t loc{};
        loc.read_description = f[0];
        const long long qstart = to_ll(f[2]), qend = to_ll(f[3]);
        loc.strand = f[4] == "+";
        loc.chromosome = f[5];
        loc.start_in_chromosome = to_ll(f[7]);
        loc.start_of_aligned_region = qstart;
        loc.size_of_aligned_region = qend - qstart;
        out.push_back(std::move(loc));
    }
    return out;
}

bool ends_with(std::string const &s, std::string const &ending)
{
    return ending.size() <= s.size() && s.compare(s.size() - ending.size(), ending.size(), ending) == 0;
}

void left_extend_locations(std::vector<CandidateLocation_t> &locations)
{
    for (CandidateLocation_t &l : locations) {
        l.start_in_chromosome = std::max(0ll, l.start_in_chromosome - l.start_of_aligned_region);
        l.size_of_aligned_region += l.start_of_aligned_region;
        l.start_of_aligned_region = 0;
    }
}

void get_global_seeds(Genome_t &genome, std::vector<CandidateLocation_t> &locations)
{
    const bool multi = genome.chromosome_starts.size() > 1;
    for (CandidateLocation_t &l : locations)
        l.start_in_reference = (multi ? genome.chromosome_starts[l.chromosome] : 0) + l.start_in_chromosome;
}

std::vector<Read_t> read_fastq_and_seed_locations(Genome_t &genome, std::string fastq_file_path, std::string seed_file_path,
                                                  std::vector<Read_t> &reads)
{
    std::vector<CandidateLocation_t> locations;
    if (ends_with(seed_file_path, ".paf")) locations = read_paf(seed_file_path);
    else if (ends_with(seed_file_path, ".maf")) locations = read_maf(seed_file_path);
    else throw std::invalid_argument("unknown seed file ending\n");
    left_extend_locations(locations);
    get_global_seeds(genome, locations);

    reads = read_fastq(fastq_file_path);
    std::unordered_map<std::string, size_t> by_name;
    by_name.reserve(reads.size() * 2);
    for (size_t i = 0; i < reads.size(); i++) by_name[reads[i].description] = i;
    for (CandidateLocation_t &l : locations) {
        auto it = by_name.find(l.read_description);
        if (it == by_name.end()) {
            std::cerr << "candidate location specified unknown read \"" << l.read_description << "\"" << std::endl;
            std::exit(1);
        }
        reads[it->second].locations.push_back(l);
    }
    return reads;
}

bool cigar_char_equals(char c, char d)
{
    const char a = (char)std::toupper((unsigned char)c), b = (char)std::toupper((unsigned char)d);
    if (a != 'A' && a != 'C' && a != 'G' && a != 'T') {
        std::cerr << "compared invalid character" << std::endl;
        return false;
    }
    return a == b;
}

// ---- --key[=value] options (reference src/util.cpp:368-427) ---------------------------------------------------
static int find_option(int argc, char **argv, const std::string &key, std::string *value)
{
    // the first argument whose key part (text before '=') equals `key` decides
    for (int i = 1; i < argc; i++) {
        const std::string arg = argv[i];
        const size_t eq = arg.find('=');
        if (arg.compare(0, eq == std::string::npos ? arg.size() : eq, key) != 0 || (eq == std::string::npos ? arg.size() : eq) != key.size())
            continue;
        if (!value) return eq == std::string::npos ? OPT_EXISTS : OPT_INVALID;   // a flag must not carry a value
        if (eq == std::string::npos || eq + 1 >= arg.size()) return OPT_INVALID;  // a valued option needs one
        *value = arg.substr(eq + 1);
        return OPT_EXISTS;
    }
    return OPT_MISSING;
}

int get_cmd_option(int argc, char **argv, std::string key) { return find_option(argc, argv, key, nullptr); }
int get_cmd_option(int argc, char **argv, std::string key, std::string &value) { return find_option(argc, argv, key, &value); }

bool check_options(int argc, char **argv, std::vector<std::string> valid_options)
{
    for (int i = 1; i < argc; i++) {
        std::string arg = argv[i];
        const size_t eq = arg.find('=');
        if (eq != std::string::npos) arg.resize(eq);
        if (std::find(valid_options.begin(), valid_options.end(), arg) == valid_options.end()) return false;
    }
    return true;
}

std::vector<std::string> parse_csv_strings(std::string csv)
{
    // "a,,b" -> {"a", "", "b"}; "" -> {""}
    std::vector<std::string> out;
    size_t i = 0;
    while (true) {
        const size_t e = csv.find(',', i);
        out.push_back(csv.substr(i, e == std::string::npos ? std::string::npos : e - i));
        if (e == std::string::npos) break;
        i = e + 1;
    }
    return out;
}

std::vector<int> parse_csv_numbers(std::string csv)
{
    std::vector<int> out;
    for (const std::string &s : parse_csv_strings(csv)) out.push_back(std::stoi(s));
    return out;
}

// ---- CIGAR validator -------------------------------------------------------------------------------------------
std::string validate_cigar(const Alignment_t &alignment, const CandidateLocation_t &location, const Read_t &read,
                           const Genome_t &reference)
{
    const std::string &cg = alignment.cigar, &ref = reference.content, &rd = read.content;
    unsigned long long i = (unsigned long long)location.start_in_reference, j = 0;
    long long edits = 0;
    size_t p = 0;
    while (p < cg.size()) {
        if (!std::isdigit((unsigned char)cg[p])) return "CIGAR had bad format";
        unsigned long long count = 0;
        while (p < cg.size() && std::isdigit((unsigned char)cg[p])) count = count * 10 + (unsigned long long)(cg[p++] - '0');
        if (p >= cg.size()) return "CIGAR had bad format";
        const char type = cg[p++];
        if (count == 0) return "CIGAR cannot contain edits with count 0";
        if (type == 'I') { j += count; edits += (long long)count; }
        else if (type == 'D') { i += count; edits += (long long)count; }
        else if (type == 'X' || type == '=' || type == 'M') {
            for (unsigned long long e = 0; e < count; e++, i++, j++) {
                if (j >= rd.size()) return "CIGAR went out of bounds of read";
                if (i >= ref.size()) return "CIGAR went out of bounds of reference";
                const bool same = std::toupper((unsigned char)ref[i]) == std::toupper((unsigned char)rd[j]);
                if (type == 'X' && same) return "CIGAR contains 'X' but reference[i] and read[j] match";
                if (type == '=' && !same) return "CIGAR contains '=' but reference[i] and read[j] mismatch";
                if (type == 'M' && !same) edits++;
            }
            if (type == 'X') edits += (long long)count;
        } else {
            return std::string("CIGAR contains unknown edit type '") + type + "'";
        }
    }
    if (j < rd.size()) return "CIGAR didn't cover entire read";
    if (j > rd.size()) return "CIGAR went out of bounds of read";
    if (i > ref.size()) return "CIGAR went out of bounds of reference";
    if (edits != alignment.edit_distance)
        return "CIGAR has " + std::to_string(edits) + " edits, while the reported edit distance is " + std::to_string(alignment.edit_distance);
    return "";
}
