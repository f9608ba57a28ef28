/*
 * scrooge_b200_bench.h -- C ABI of libscrooge_b200_bench.so: measurement, synthetic-data and checking helpers.
 *
 * NOT part of the drop-in boundary (that is scrooge_b200.h / libscrooge_b200.so, which holds the alignment path only).
 * bench.py, the tests and apps/ use these to generate BASELINE.json's synthetic workloads, to measure the roofline
 * denominators on the box, and to check a whole batch's CIGAR runs on the device.
 */
#ifndef SCROOGE_B200_BENCH_H
#define SCROOGE_B200_BENCH_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Status codes are scrooge_b200.h's (0 = ok).  Message of the last failing call of THIS library on this thread. */
const char *sg_bench_last_error(void);

/* Measurement / test helper: consistency of a batch's compacted runs with its other results, for EVERY alignment --
 * the sequence-independent properties of the reference's validateCigarString (src/tests.cu:106-169): every run count
 * in [1, max_count] (W-O: 31 at W=64/O=33, 15 at W=32/O=17), counts of =,X,I sum to d_query_len, of =,X,D to d_ref_consumed, of
 * X,I,D to d_edit.  *d_n_bad (device uint64, caller-initialised) is incremented once per offending alignment. */
int sg_dev_check_runs(const uint8_t *d_runs, const uint64_t *d_run_off, uint64_t n, const uint64_t *d_query_len,
                      const int64_t *d_edit, const uint64_t *d_ref_consumed, uint32_t max_count, uint64_t *d_n_bad,
                      void *stream);

/* Sustained 32-bit integer ALU throughput probe (LOP3 + SHF mix, the instruction mix of the DC
 * recurrence): runs for roughly `ms` milliseconds and returns giga-ops/s through *gops.  kind: 0 LOP3 only,
 * 1 SHF only, 2 LOP3+SHF 2:1 (the roofline denominator), 3 LOP3+IMAD 1:1, 4-6 one DC entry (4 LOP3 + a 64-bit
 * shift) with the shift as IMAD+SHF / IMAD.SHL+IMAD.WIDE / IMAD.HI+IMAD+SHL (counted as 6 ops per entry). */
int sg_dev_int32_peak(int kind, double ms, double *gops);

/* Synthetic pair generator (deterministic in (seed, pair index); identical on host and device).
 * The text is i.i.d. uniform ACGT; the read walks the text applying, with probability err per text base,
 * a substitution / insertion / deletion chosen with weights w_sub:w_ins:w_del, until it holds read_len
 * bases; `slack` random bases are then appended to the text.  Pair p's text occupies
 * [p*text_stride, p*text_stride + text_len[p]) and its read [p*read_len, (p+1)*read_len).
 * text_stride must be >= sg_synth_text_stride(read_len, slack). */
uint64_t sg_synth_text_stride(uint32_t read_len, uint32_t slack);
int sg_synth_pairs_host(uint64_t seed, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, double err,
                        uint32_t w_sub, uint32_t w_ins, uint32_t w_del, uint32_t slack, char *text,
                        uint64_t text_stride, uint64_t *text_len, char *reads);
int sg_dev_synth_pairs(uint64_t seed, uint64_t first_pair, uint64_t n_pairs, uint32_t read_len, double err,
                       uint32_t w_sub, uint32_t w_ins, uint32_t w_del, uint32_t slack, char *d_text,
                       uint64_t text_stride, uint64_t *d_text_len, char *d_reads, void *stream);

/* Read-mapping workload (BASELINE.json configs[3]).  sg_synth_genome writes bases [first, first+n) of the i.i.d.
 * uniform genome `seed` to host memory `out` and/or device memory `d_out` (either may be NULL).  sg_synth_reads draws
 * read r at a uniform position of `genome` (host or device memory according to on_device, like reads / pos) and
 * mutates it as sg_synth_pairs does; pos[r] receives the true start. */
int sg_synth_genome(uint64_t seed, uint64_t first, uint64_t n, char *out, void *d_out, void *stream);
int sg_synth_reads(uint64_t seed, uint64_t first_read, uint64_t n_reads, uint32_t read_len, double err, uint32_t w_sub,
                   uint32_t w_ins, uint32_t w_del, const char *genome, uint64_t genome_len, char *reads, uint64_t *pos,
                   int on_device, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* SCROOGE_B200_BENCH_H */
