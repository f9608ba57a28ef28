/*
 * scrooge_b200.h -- C ABI of the B200-native (sm_100a) Scrooge/GenASM aligner.
 *
 * This is the drop-in boundary: plain pointers and sizes, no C++ or torch types.  It replaces the
 * reference's GPU library interface
 *     genasm_gpu::align_all(std::vector<std::string>& texts, std::vector<std::string>& queries, long long*)
 *                                                                  (reference src/genasm_gpu.hpp:8)
 *     genasm_gpu::align_all(Genome_t& reference, std::vector<Read_t>& reads, long long*)
 *                                                                  (reference src/genasm_gpu.hpp:7)
 * The C++ overloads with the reference's exact signatures live in include/genasm_gpu.hpp and are thin
 * wrappers over the sg_* functions below.
 *
 * Two layers:
 *   1. host API  (sg_ctx_*, sg_align_pairs, sg_set_reference, sg_align_candidates, sg_result_*):
 *      HOST buffers in, HOST results out; owns devices, streams, pinned staging and the per-GPU replicated
 *      packed reference.  Work is scattered over the context's GPUs with no inter-GPU exchange.
 *   2. device API (sg_dev_*): DEVICE pointers and an explicit cudaStream_t (passed as void*), current
 *      device, asynchronous.  This is what layer 1 is built from; benchmarks and tests call it directly
 *      to time the kernels with inputs resident in HBM.
 *
 * There is no CPU fallback anywhere: every entry point that computes needs a CUDA device and fails with
 * SG_ERR_CUDA (and a message in sg_last_error) when there is none.
 *
 * Semantics (bit-exact with reference src/genasm_cpu.cpp): semi-global edit distance of the whole query
 * against a prefix of the text, computed over a chain of W x W windows with overlap O (K=W); traceback
 * priority I > D > X > '='; CIGAR runs are run-length encoded within a window and never merged across windows.
 * The window configuration is a run-time choice where the reference needs a rebuild (-DCLI_W/-DCLI_O,
 * src/genasm_cpu.cpp:22-35): W=64/O=33 (in-file default, src/genasm_cpu.cpp:7-9) and W=32/O=17 (README.md:208)
 * run on kernels tuned for them, any other 2 <= W <= 256, 0 <= O < W, W-O <= 128 (the axes of
 * scripts/profile.py:66-100,595-640) on a general kernel; results are bit-exact with the reference built at
 * the same (W, O).
 */
#ifndef SCROOGE_B200_H
#define SCROOGE_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------------------------------------ */
/* status codes (the reference exit()s / assert()s instead: src/cuda_util.hpp:3-10,                   */
/* src/genasm_gpu.cu:636,931-934,984)                                                               */
#define SG_OK                 0
#define SG_ERR_CUDA           1  /* CUDA runtime error or no device */
#define SG_ERR_BAD_BASE       2  /* a character outside ACGTacgt (sg_last_error names the position) */
#define SG_ERR_BAD_ARG        3
#define SG_ERR_OOM            4
#define SG_ERR_CIGAR_OVERFLOW 5  /* an alignment produced more runs than its slab capacity */
#define SG_ERR_NO_REFERENCE   6  /* sg_align_candidates before sg_set_reference */

/* flags */
#define SG_FLAG_DISTANCE_ONLY 1u /* skip CIGAR storage and writeback (traceback still runs) */
/* device API only: the caller promises that d_slab and every d_slab_off[a] are multiples of 4.  The tuned kernels (64/33,
 * 32/17) then collect the runs of an alignment in a register and store them as whole 32-bit words -- a quarter of the
 * store instructions and L2 sector writes of the default one-byte-per-run stores; the up to 3 bytes between an
 * alignment's last run and the next 4-byte boundary of its slot are written as padding.  Same results, bit for bit.
 * Measured on a B200 (apps/sg_variant_ab, profiles/r02_variant_ab.jsonl): 10 kbp pairs at 10 % 1.02 x, candidate lists
 * with 7 of 8 loci spurious (~20 runs per window instead of 6.4) 1.42 x, 150 bp reads 1.04 x at 32/17 and 1.00 x at 64/33.
 * The host API lays its slab out accordingly and always launches this way (SG_EMIT=bytes restores byte stores).
 * Ignored by other window configurations and with SG_FLAG_DISTANCE_ONLY. */
#define SG_FLAG_RUN_WORDS 2u

/* A CIGAR run as the kernels store it: one byte, (op << 6) | count, count in 1..W-O.  Window configurations with
 * W-O > 63 split a longer run: bytes with count 0 each stand for 63 more of the same op and the byte that follows them
 * carries the rest (sg_result_render_* / sg_result_entries / sg_result_cigar_len join them; run offsets count bytes).
 * op: 0 '=', 1 'X', 2 'I', 3 'D'.  sg_cigar_entry is the reference's CigarEntry_t (src/util.hpp:43-46). */
#define SG_RUN_OP(b)    ((unsigned)(b) >> 6)
#define SG_RUN_COUNT(b) ((unsigned)(b) & 63u)
typedef struct sg_cigar_entry { uint8_t edit_count; char edit_type; } sg_cigar_entry;

typedef struct sg_ctx sg_ctx;
typedef struct sg_result sg_result;

/* Message for the last failing call on this thread ("" when none). */
const char *sg_last_error(void);
/* ABI version, bumped on incompatible change. */
int sg_version(void);
/* Number of CUDA devices visible (0 when there is no driver/device). */
int sg_device_count(void);

/* ------------------------------------------------------------------------------------------------ */
/* 1. host API                                                                                      */

/* Create a context over n_devices GPUs (device_ids == NULL: devices 0..n_devices-1; n_devices == 0: all).
 * W is 64 (O=33) or 32 (O=17).  Replaces the reference's hard-wired GPU_ID 0 (src/genasm_gpu.cu:67). */
/* Host threads: every GPU of a context gets a disjoint share of the CPUs the process may run on (its affinity mask, or
 * SG_CPUS=<cpu list>), NUMA-local to the GPU where the machine says so, and a persistent team of packer threads bound to
 * it (SG_HOST_THREADS=<n per GPU>, 0 = none: all input crosses PCIe as ASCII; SG_AFFINITY=0 = no binding).  During a call
 * on a single-GPU context the calling thread is bound to that share too and restored on return; the caller's current
 * CUDA device is restored likewise. */
int sg_ctx_create(sg_ctx **out, const int *device_ids, int n_devices, int W);
/* The same with an explicit window configuration: 2 <= W <= 256, 0 <= O < W, W-O <= 128.  Replaces a rebuild of the reference with -DCLI_W=<W> -DCLI_K=<W> -DCLI_O=<O>
 * (src/genasm_cpu.cpp:22-35, scripts/profile.py:29,132). */
int sg_ctx_create_wo(sg_ctx **out, const int *device_ids, int n_devices, int W, int O);
/* O = min(W/2+1, W-1): the overlap the reference pairs with a window size (scripts/profile.py:78,619). */
int sg_default_overlap(int W);
int sg_ctx_window(const sg_ctx *ctx);
int sg_ctx_overlap(const sg_ctx *ctx);
void sg_ctx_destroy(sg_ctx *ctx);
int sg_ctx_num_devices(const sg_ctx *ctx);

/* Unstructured interface (reference src/genasm_gpu.cu:982-1065).  Blobs are ASCII, strings concatenated
 * without separators; *_off have n_pairs+1 entries.  Pair p aligns query p against text p.
 * Results come back in input order in *out (release with sg_result_free). */
int sg_align_pairs(sg_ctx *ctx, const char *text_blob, const uint64_t *text_off, const char *query_blob,
                   const uint64_t *query_off, uint64_t n_pairs, uint32_t flags, sg_result **out);
/* The same over separate strings (pointer + length per string, e.g. the data() / size() of the std::strings of the
 * reference's std::vector<std::string> arguments): nothing is flattened, every string is packed in place by the host
 * threads (or gathered into pinned staging for the device ingest). */
int sg_align_pairs_v(sg_ctx *ctx, const char *const *texts, const uint64_t *text_len, const char *const *queries,
                     const uint64_t *query_len, uint64_t n_pairs, uint32_t flags, sg_result **out);

/* Read-mapping interface (reference src/genasm_gpu.cu:890-980).  sg_set_reference packs the genome to
 * 2 bit/base once and replicates it into every GPU's HBM (reference twobit_reference, :692-748).
 * Candidate c aligns read cand_read[c] against the genome suffix starting at cand_start[c]
 * (src/genasm_gpu.cu:716-726); reads are packed once and shared by their candidates (:784-796).
 * Results are in candidate order. */
int sg_set_reference(sg_ctx *ctx, const char *genome_ascii, uint64_t genome_len);
int sg_align_candidates(sg_ctx *ctx, const char *read_blob, const uint64_t *read_off, uint64_t n_reads,
                        const uint64_t *cand_start, const uint32_t *cand_read, uint64_t n_cand,
                        uint32_t flags, sg_result **out);
int sg_align_candidates_v(sg_ctx *ctx, const char *const *reads, const uint64_t *read_len, uint64_t n_reads,
                          const uint64_t *cand_start, const uint32_t *cand_read, uint64_t n_cand,
                          uint32_t flags, sg_result **out);

/* Result accessors.  All pointers stay valid until sg_result_free. */
uint64_t sg_result_count(const sg_result *r);
const int64_t *sg_result_edit_distances(const sg_result *r);   /* [count] */
const uint64_t *sg_result_ref_consumed(const sg_result *r);    /* [count] consumed text prefix = #(=,X,D) */
const uint64_t *sg_result_run_offsets(const sg_result *r);     /* [count+1] into runs; NULL if distance-only */
const uint8_t *sg_result_runs(const sg_result *r);             /* packed runs, see SG_RUN_OP/COUNT */
/* device time of the alignment kernels only, max over GPUs: the reference's core_algorithm_ns
 * (src/genasm_gpu.cu:940-948) */
int64_t sg_result_kernel_ns(const sg_result *r);
/* wall time of the whole call */
int64_t sg_result_total_ns(const sg_result *r);
/* Length of alignment idx's CIGAR text ("%d%c" per run, reference src/genasm_gpu.cu:881-888). */
uint64_t sg_result_cigar_len(const sg_result *r, uint64_t idx);
/* Render alignment idx's CIGAR text into buf (cap bytes incl. NUL).  Returns the length, or -1 if cap is
 * too small. */
int64_t sg_result_render_cigar(const sg_result *r, uint64_t idx, char *buf, uint64_t cap);
/* Expand alignment idx's runs into CigarEntry_t-shaped entries; returns the number written, -1 if cap is
 * too small. */
int64_t sg_result_entries(const sg_result *r, uint64_t idx, sg_cigar_entry *out, uint64_t cap);
/* All CIGAR texts at once, rendered by `threads` host threads (0 = all): alignment a's text is
 * blob[text_off[a] .. text_off[a+1]) (no terminators); text_off has count+1 entries.  Returns the total text length;
 * when blob is NULL or blob_cap is too small only text_off is filled (call again with a big enough blob). */
uint64_t sg_result_render_all(const sg_result *r, char *blob, uint64_t blob_cap, uint64_t *text_off, int threads);
void sg_result_free(sg_result *r);

/* Where the time and the PCIe bytes of the call that produced r went (the end-to-end path is bound by the host, so this is
 * what explains an end-to-end number).  Times are wall-clock; "max over GPUs" = the slowest GPU's worker thread. */
typedef struct sg_call_stats {
    int64_t total_ns;           /* the whole call */
    int64_t kernel_ns;          /* alignment kernels only, max over GPUs (= sg_result_kernel_ns) */
    int64_t upload_ns;          /* ingest phase of all sub-batches: host packing + queuing of the copies, max over GPUs */
    int64_t pack_thread_ns;     /* time inside the packing loops, summed over all packer threads of all GPUs */
    int64_t wait_ns;            /* blocked on the device (kernels, compaction, copies back), max over GPUs */
    int64_t host_other_ns;      /* descriptors + result bookkeeping, max over GPUs */
    uint64_t h2d_ascii_bytes;   /* crossed PCIe as ASCII (packed by the ingest kernel) */
    uint64_t h2d_packed_bytes;  /* crossed PCIe at 2 bit/base (packed by the host threads) */
    uint64_t h2d_other_bytes;   /* descriptors */
    uint64_t d2h_bytes;         /* distances, consumed prefixes, run offsets, packed runs */
    uint32_t n_devices, sub_batches, host_threads_per_device;
    uint32_t packers_in_use;    /* packer threads per GPU the ingest used for its last sub-batch (chosen from measured rates:
                                 * all, half or none of host_threads_per_device; SG_PACKERS=<n> fixes it, SG_TUNE=0 = all) */
} sg_call_stats;
int sg_result_stats(const sg_result *r, sg_call_stats *out);

/* Result blocks live in page-locked memory; freed blocks are kept for reuse in a per-process cache of at most
 * min(RAM/32, 4 GB) (SG_PINNED_CACHE_GB=<n> overrides, 0 = keep nothing).  sg_trim_host_cache releases the cache now; the
 * destruction of a process's last context does the same. */
void sg_trim_host_cache(void);

/* How a call over n alignments would be cut into sub-batches (the unit of the per-GPU pipeline and of the queue the GPUs
 * of a context share): weight_off = n+1 prefix sums of the bytes an alignment uploads, per_unit_extra = descriptor bytes
 * per alignment; a sub-batch grows while it is under max_batch_bytes and either under batch_bytes or short of
 * min_batch_units alignments; with taper the last full sub-batch is cut into 1/2, 1/4, 1/8, 1/8 of its weight (the end of
 * a call is exposed).  Writes up to cuts_cap cut points (first 0, last n) and returns how many there are.  No device
 * needed: planning and tests. */
uint64_t sg_plan_sub_batches(const uint64_t *weight_off, uint64_t n, uint64_t per_unit_extra, uint64_t batch_bytes,
                             uint64_t max_batch_bytes, uint64_t min_batch_units, int taper, uint64_t *cuts, uint64_t cuts_cap);

/* Page-locked host memory for input blobs: uploads from it run at full PCIe speed and overlap with compute
 * (pageable memory works too, but the driver then stages every copy).  NULL on failure. */
void *sg_host_alloc(uint64_t bytes);
void sg_host_free(void *p);

/* Host-side ingest: the same packing as sg_dev_pack_2bit done by `threads` host threads (0 = all), AVX-512 / AVX2 when
 * the CPU has them.  packed must hold ceil(n_bases/16) words.  Returns the smallest offending position or UINT64_MAX.
 * sg_host_pack_isa: 2 AVX-512, 1 AVX2, 0 scalar. */
uint64_t sg_host_pack_2bit(const char *ascii, uint64_t n_bases, uint32_t *packed, int threads);
int sg_host_pack_isa(void);

/* ------------------------------------------------------------------------------------------------ */
/* 2. device API: device pointers, current device, asynchronous on `stream` (a cudaStream_t).        */

/* ASCII -> 2 bit/base, 16 bases per little-endian 32-bit word, base k of a word in bits 2k+1:2k,
 * A=0 C=1 G=2 T=3, case-insensitive (codes as reference src/genasm_cpu.cpp:87-90).  d_packed must hold
 * sg_packed_words(n_bases) words.  *d_bad_pos (device uint64, initialise to UINT64_MAX) receives the
 * smallest offending position if any character is outside ACGTacgt.
 * Replaces single_ascii_to_twobit_string (reference src/genasm_gpu.cu:640-685). */
uint64_t sg_packed_words(uint64_t n_bases);
int sg_dev_pack_2bit(const char *d_ascii, uint64_t n_bases, uint32_t *d_packed, uint64_t *d_bad_pos,
                     void *stream);
/* The same with flags.  SG_PACK_SIDE: the launch is meant to run BESIDE the alignment kernel of another batch (other
 * stream): CTAs of 32 KB shared memory that fit the slot sg_dev_align leaves free on every SM, instead of 96 KB ones. */
#define SG_PACK_SIDE 1u
int sg_dev_pack_2bit_ex(const char *d_ascii, uint64_t n_bases, uint32_t *d_packed, uint64_t *d_bad_pos, uint32_t flags,
                        void *stream);

/* The alignment kernel (DC + TB + per-window RLE), one launch over n alignments.
 *   d_text / d_query    packed blobs (may be the same blob)
 *   d_text_start/len    per alignment: first base and number of bases of the text in d_text
 *   d_query_start/len   same for the query
 *   d_slab, d_slab_off  run slab and n+1 byte offsets into it: alignment a may write at most
 *                       d_slab_off[a+1]-d_slab_off[a] runs (ignored with SG_FLAG_DISTANCE_ONLY)
 *   d_counter           device uint64 work-queue head; the call zeroes it on `stream`
 * outputs per alignment: d_edit (int64), d_ref_consumed (uint64), d_nruns (uint32),
 *   d_status (uint8: SG_OK or SG_ERR_CIGAR_OVERFLOW); optional (may be NULL) d_dc_entries (uint64): the
 *   alignment's algorithmic DC work, sum over its windows of (d_w+1)*(n_w+1) R[d][i] entries -- the
 *   early-termination-minimal count of reference src/genasm_cpu.cpp:214-216,278-283 (roofline numerator);
 *   optional d_windows (uint32): the number of windows the alignment took (iterations of the reference's loop
 *   src/genasm_cpu.cpp:417-436) -- the unit of the work the delta-encoded kernel actually does
 * Replaces genasm_kernel (reference src/genasm_gpu.cu:583-629).  W is 64 or 32. */
int sg_dev_align(int W, const uint32_t *d_text, const uint64_t *d_text_start, const uint64_t *d_text_len,
                 const uint32_t *d_query, const uint64_t *d_query_start, const uint64_t *d_query_len,
                 uint64_t n, uint32_t flags, uint8_t *d_slab, const uint64_t *d_slab_off,
                 uint64_t *d_counter, int64_t *d_edit, uint64_t *d_ref_consumed, uint32_t *d_nruns,
                 uint8_t *d_status, uint64_t *d_dc_entries, uint32_t *d_windows, void *stream);
/* The same for any window configuration (limits as sg_ctx_create_wo); sg_dev_align(W, ...) is
 * sg_dev_align_wo(W, sg_default_overlap(W), ...). */
int sg_dev_align_wo(int W, int O, const uint32_t *d_text, const uint64_t *d_text_start, const uint64_t *d_text_len,
                    const uint32_t *d_query, const uint64_t *d_query_start, const uint64_t *d_query_len,
                    uint64_t n, uint32_t flags, uint8_t *d_slab, const uint64_t *d_slab_off,
                    uint64_t *d_counter, int64_t *d_edit, uint64_t *d_ref_consumed, uint32_t *d_nruns,
                    uint8_t *d_status, uint64_t *d_dc_entries, uint32_t *d_windows, void *stream);

/* The same with a launch order: the kernel's work queue hands out alignment d_order[k] as its k-th item (d_order: a
 * permutation of 0..n-1, n < 2^32; NULL = input order).  Results stay indexed by alignment.  With alignments of very
 * different lengths in one launch, longest-first order keeps the last lanes from finishing long after the rest -- what
 * the reference's callers do by sorting their reads before the call (src/tests.cu:377). */
int sg_dev_align_ordered(int W, int O, const uint32_t *d_text, const uint64_t *d_text_start, const uint64_t *d_text_len,
                         const uint32_t *d_query, const uint64_t *d_query_start, const uint64_t *d_query_len,
                         uint64_t n, uint32_t flags, uint8_t *d_slab, const uint64_t *d_slab_off,
                         uint64_t *d_counter, int64_t *d_edit, uint64_t *d_ref_consumed, uint32_t *d_nruns,
                         uint8_t *d_status, uint64_t *d_dc_entries, uint32_t *d_windows, const uint32_t *d_order,
                         void *stream);

/* CIGAR compaction: exclusive scan of d_nruns into d_run_off[n+1] (d_scan_tmp: sg_scan_tmp_bytes(n)
 * bytes), then gather every alignment's runs from its slab slot into one dense array.
 * Replaces the reference's linked-list walk (src/cuda_list.hpp, src/genasm_gpu.cu:881-888). */
uint64_t sg_scan_tmp_bytes(uint64_t n);
int sg_dev_scan_runs(const uint32_t *d_nruns, uint64_t n, uint64_t *d_run_off, void *d_scan_tmp, void *stream);
int sg_dev_gather_runs(const uint8_t *d_slab, const uint64_t *d_slab_off, const uint32_t *d_nruns,
                       const uint64_t *d_run_off, uint64_t n, uint8_t *d_runs, void *stream);
/* The same with what the caller knows about the batch: runs_per_alignment_hint = an upper estimate of the runs of a typical
 * alignment (e.g. its slab capacity; 0 = unknown).  Batches of short alignments (hint <= 1024) are gathered by four lanes
 * per alignment instead of a warp. */
int sg_dev_gather_runs_sized(const uint8_t *d_slab, const uint64_t *d_slab_off, const uint32_t *d_nruns,
                             const uint64_t *d_run_off, uint64_t n, uint8_t *d_runs, uint64_t runs_per_alignment_hint,
                             void *stream);

/* Launch geometry of sg_dev_align on the current device: persistent warps per SM and shared memory per
 * warp (for reports). */
int sg_dev_align_geometry(int W, int *warps_per_sm, int *smem_per_warp, int *num_sms);
int sg_dev_align_geometry_wo(int W, int O, int *warps_per_sm, int *smem_per_warp, int *num_sms);

#ifdef __cplusplus
}
#endif
#endif /* SCROOGE_B200_H */
