/*
 * ref_io_dump.cpp -- TEST INFRASTRUCTURE ONLY.  Compiled by oracle/Makefile together with the UNMODIFIED reference
 * src/util.cpp (read from where it lies); prints what the reference's own readers make of a FASTA / FASTQ / MAF|PAF
 * triple in the same format as `build/sg_tests --dump_inputs`, so that tests/test_io.py can diff the two.
 * Mirrors the input stage of the reference's driver (src/tests.cu:339-355,377).
 */
#include <algorithm>
#include <iostream>
#include <string>
#include <vector>
#include "util.hpp"

using namespace std;

int main(int argc, char **argv)
{
    if (argc != 4) { cerr << "usage: ref_io_dump reference.fasta reads.fastq seeds.(maf|paf)" << endl; return 2; }
    Genome_t genome = read_genome(argv[1]);
    vector<Read_t> reads;
    read_fastq_and_seed_locations(genome, argv[2], argv[3], reads);
    for (Read_t &read : reads)
        read.locations.erase(remove_if(read.locations.begin(), read.locations.end(), [](CandidateLocation_t const &l) { return l.strand == false; }),
                             read.locations.end());
    stable_sort(reads.begin(), reads.end(), [](const Read_t &a, const Read_t &b) { return a.content.size() > b.content.size(); });
    cout << "genome " << genome.content.size() << " bases, " << genome.chromosome_starts.size() << " chromosome(s)" << endl;
    for (const auto &kv : genome.chromosome_starts) cout << "chromosome \"" << kv.first << "\" starts at " << kv.second << endl;
    for (const Read_t &r : reads) {
        cout << "read \"" << r.description << "\" " << r.content.size() << " bases:";
        for (const CandidateLocation_t &l : r.locations) cout << " " << l.chromosome << "@" << l.start_in_chromosome << "->" << l.start_in_reference;
        cout << endl;
    }
    return 0;
}
