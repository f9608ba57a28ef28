// scrooge_types.hpp -- the data types on the library boundary.
//
// Field names, order and types reproduce the reference's boundary structs (reference src/util.hpp:11-46)
// so that code written against genasm_gpu::align_all compiles and links unchanged.  Only the layout is
// shared; the file-format readers the reference declares next to them (FASTA/FASTQ/MAF/PAF, SURVEY.md
// section 8f-3) are outside this library.
#pragma once
// The reference's util.hpp defines the same names (it sets SEED_FILE_MAF, src/util.hpp:8): when it came first, its
// definitions stand and this header adds nothing.
#if !defined(SEED_FILE_MAF) || defined(SCROOGE_B200_TYPES)
#ifndef SCROOGE_B200_TYPES
#define SCROOGE_B200_TYPES 1
#endif

#include <chrono>
#include <cstdint>
#include <map>
#include <string>
#include <vector>

struct Sequence {
    std::string description;
    std::string content;
};
typedef Sequence Sequence_t;

// A reference genome: all chromosomes concatenated in `content`; chromosome_starts maps a chromosome name to
// the offset of its first base.  Only `content` is read by the aligner.
struct Genome {
    std::map<std::string, long long> chromosome_starts;
    std::string content;
};
typedef Genome Genome_t;

// Where a read may align.  The aligner uses start_in_reference only (reference src/genasm_cpu.cpp:512-513);
// callers filter by strand themselves (reference src/tests.cu:347-355).
struct CandidateLocation {
    std::string read_description;
    std::string chromosome;
    long long start_in_chromosome;
    long long start_in_reference;
    long long start_of_aligned_region;
    long long size_of_aligned_region;
    bool strand;
};
typedef CandidateLocation CandidateLocation_t;

struct Read {
    std::string description;
    std::string content;
    std::vector<CandidateLocation_t> locations;
};
typedef Read Read_t;

struct Alignment {
    std::string cigar;
    long long edit_distance;
};
typedef Alignment Alignment_t;

struct CigarEntry {
    uint8_t edit_count;
    char edit_type;
};
typedef CigarEntry CigarEntry_t;

// Wall-clock nanoseconds of a callable (the reference's timing helper, src/util.hpp:48-55).
template <typename Fn> long long measure_ns(Fn fn)
{
    const auto t0 = std::chrono::high_resolution_clock::now();
    fn();
    const auto t1 = std::chrono::high_resolution_clock::now();
    return std::chrono::duration_cast<std::chrono::nanoseconds>(t1 - t0).count();
}

#endif  // the reference's util.hpp was not included first
