// scrooge_io.hpp -- dataset readers next to the aligner.
//
// Same function names, argument order and results as the reference declares in src/util.hpp:57-90 and implements
// in src/util.cpp:30-459 (FASTA genome, FASTQ reads, MAF / PAF candidate locations, left extension of candidates
// to the read start, chromosome-relative -> genome-relative coordinates, --key=value options), so that the reference's
// own drivers (src/tests.cu, src/cpu_baseline.cpp) compile against this include directory.  The implementation
// (scrooge_b200/csrc/sg_io.cpp) is written from the formats, as single-pass scanners over the file image.
#pragma once

#include <string>
#include <vector>

#include "scrooge_types.hpp"

#define SEED_FILE_MAF 0
#define SEED_FILE_PAF 1

void remove_whitespaces(std::string &str);
std::string read_file(std::string file_path);

// every '>' record of a FASTA file: description = rest of the header line, content = sequence lines joined
std::vector<Sequence_t> read_fasta(std::string file_path);
// all chromosomes concatenated; chromosome_starts[description] = offset of its first base
Genome_t read_genome(std::string fasta_file_path);
// '@' records: description (blanks dropped) and the single sequence line; locations left empty
std::vector<Read_t> read_fastq(std::string file_path);
// MAF blocks ("a" line followed by "s ref ..." and "s <read> ..." lines): start_in_chromosome from the ref line,
// read name / strand / aligned region from the read line
std::vector<CandidateLocation_t> read_maf(std::string file_path);
// PAF lines: qname qlen qstart qend strand tname tlen tstart tend matches alnlen ...
std::vector<CandidateLocation_t> read_paf(std::string file_path);

bool ends_with(std::string const &s, std::string const &ending);
// move every candidate left by the unaligned read prefix so that it points at where the read's first base would map
void left_extend_locations(std::vector<CandidateLocation_t> &locations);
// start_in_reference = chromosome start (multi-chromosome genomes only) + start_in_chromosome
void get_global_seeds(Genome_t &genome, std::vector<CandidateLocation_t> &locations);
// reads from FASTQ, candidates from .maf/.paf, left-extended, made global and attached to their reads by name
std::vector<Read_t> read_fastq_and_seed_locations(Genome_t &genome, std::string fastq_file_path, std::string seed_file_path,
                                                  std::vector<Read_t> &reads);

// case-insensitive base equality used by the CIGAR validator
bool cigar_char_equals(char c, char d);

#define OPT_MISSING 0
#define OPT_EXISTS 1
#define OPT_INVALID 2
int get_cmd_option(int argc, char **argv, std::string key);
int get_cmd_option(int argc, char **argv, std::string key, std::string &value);
bool check_options(int argc, char **argv, std::vector<std::string> valid_options);
std::vector<int> parse_csv_numbers(std::string csv);
std::vector<std::string> parse_csv_strings(std::string csv);

// CIGAR validator with the checks of the reference's validateCigarString (src/tests.cu:27-169): format, the whole
// read covered, inside the reference, '='/'X' agreeing with the bases, #edits == edit_distance.  Returns an empty
// string when valid, else the reason.
std::string validate_cigar(const Alignment_t &alignment, const CandidateLocation_t &location, const Read_t &read,
                           const Genome_t &reference);
