// library_example.cpp -- the GPU half of the reference's library example (reference
// src/library_example.cu:25-88) compiled against this repository's drop-in header: the same calls, the
// same types, the same output format.
#include <iostream>
#include <string>
#include <vector>

#include "genasm_gpu.hpp"

using namespace std;

static void print(const vector<Alignment_t> &alignments)
{
    for (const Alignment_t &aln : alignments)
        cout << "edit_distance:" << aln.edit_distance << " cigar:" << aln.cigar << endl;
}

static void string_pairs_example()
{
    vector<string> texts = {"ACGTACGT"};
    vector<string> queries = {"ACGTACG"};
    print(genasm_gpu::align_all(texts, queries));
}

static void mapping_example()
{
    Genome_t reference;
    reference.content = "ACGTACGT";

    CandidateLocation_t ref_begin;
    ref_begin.start_in_reference = 0;
    ref_begin.strand = true;

    Read_t read;
    read.description = "example_read_id";
    read.content = "ACGTACG";
    read.locations = vector<CandidateLocation_t>(1, ref_begin);
    vector<Read_t> reads(1, read);

    print(genasm_gpu::align_all(reference, reads));
}

int main()
{
    genasm_gpu::enabled_algorithm_log = false;
    try {
        string_pairs_example();
        mapping_example();
    } catch (const exception &e) {
        cerr << e.what() << endl;
        return 1;
    }
    return 0;
}
