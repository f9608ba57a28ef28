// sim_kernels.cpp -- the kernels' SOURCE (scrooge_b200/csrc/sg_align_delta.cuh, sg_align_generic.cuh, sg_aux.cuh) compiled for
// the host and run thread by thread on the fiber scheduler of sim_runtime.cpp.  TEST INFRASTRUCTURE: built and loaded by
// tests/test_kernel_sim.py only; nothing under scrooge_b200/ knows it exists.  Not simulated: the bulk-copy-staged ingest
// kernel (mbarrier / cp.async.bulk) and the row-wise formulation (tensor-memory PTX).
#include <vector>
#include "sim_runtime.h"
#include "../../scrooge_b200/csrc/sg_align_delta.cuh"
#include "../../scrooge_b200/csrc/sg_align_generic.cuh"
// function-scope __shared__ arrays (block_exclusive_scan's warp sums) are one array per CTA: a static of the host function
#undef __shared__
#define __shared__ static
#include "../../scrooge_b200/csrc/sg_aux.cuh"
#include "../../scrooge_b200/csrc/sg_bench_aux.cuh"

namespace sg {
__attribute__((aligned(16))) uint32_t smem_all[(227 * 1024) / 4];   // what `extern __shared__ uint32_t smem_all[]` in the kernels resolves to
}

namespace {
template <int W, int EMIT> void body(void *arg) { sg::genasm_delta_kernel<W, EMIT>(*static_cast<const sg::AlignParams *>(arg)); }
}

namespace {
struct GenericArgs { sg::AlignParams P; sg::GenericGeom G; };
template <int NW, bool GP, bool WIDE> void generic_body(void *arg)
{
    const GenericArgs *a = static_cast<const GenericArgs *>(arg);
    sg::genasm_generic_kernel<NW, GP, WIDE>(a->P, a->G);
}
template <int NW> void (*generic_fn(bool gp, bool wide))(void *)
{
    return gp ? (wide ? generic_body<NW, true, true> : generic_body<NW, true, false>) : (wide ? generic_body<NW, false, true> : generic_body<NW, false, false>);
}

struct PackArgs { const char *ascii; uint64_t n_bases; uint32_t *packed; uint64_t n_words; unsigned long long *bad; uint64_t w_first; };
void pack_body(void *arg) { const PackArgs *a = static_cast<const PackArgs *>(arg); sg::pack_2bit_kernel(a->ascii, a->n_bases, a->packed, a->n_words, a->bad, a->w_first); }

struct ScanArgs { const uint32_t *in; uint64_t n; uint64_t *tmp; uint64_t *out; uint64_t tiles; };
void scan1_body(void *arg) { const ScanArgs *a = static_cast<const ScanArgs *>(arg); sg::scan_tile_sums_kernel(a->in, a->n, a->tmp); }
void scan2_body(void *arg) { const ScanArgs *a = static_cast<const ScanArgs *>(arg); sg::scan_tile_offsets_kernel(a->tmp, a->tiles); }
void scan3_body(void *arg) { const ScanArgs *a = static_cast<const ScanArgs *>(arg); sg::scan_finish_kernel(a->in, a->n, a->tmp, a->out); }

struct GatherArgs { const uint8_t *slab; const uint64_t *slab_off; const uint32_t *nruns; const uint64_t *run_off; uint64_t n; uint8_t *runs; };
template <int GROUP> void gather_body(void *arg)
{
    const GatherArgs *a = static_cast<const GatherArgs *>(arg);
    sg::gather_runs_kernel<GROUP>(a->slab, a->slab_off, a->nruns, a->run_off, a->n, a->runs);
}
struct CheckArgs { const uint8_t *runs; const uint64_t *run_off; uint64_t n; const uint64_t *query_len; const int64_t *edit;
                   const uint64_t *ref_consumed; uint32_t max_count; unsigned long long *n_bad; };
void check_body(void *arg)
{
    const CheckArgs *a = static_cast<const CheckArgs *>(arg);
    sg::check_runs_kernel(a->runs, a->run_off, a->n, a->query_len, a->edit, a->ref_consumed, a->max_count, a->n_bad);
}
}  // namespace

extern "C" {

// check_runs_kernel of the bench library (sg_dev_check_runs): *n_bad += alignments whose runs contradict their other results
int sim_check_runs(const uint8_t *runs, const uint64_t *run_off, uint64_t n, const uint64_t *query_len, const int64_t *edit,
                   const uint64_t *ref_consumed, uint32_t max_count, uint64_t *n_bad, unsigned blocks)
{
    CheckArgs a{runs, run_off, n, query_len, edit, ref_consumed, max_count, (unsigned long long *)n_bad};
    sim::launch(blocks, 256, sg::smem_all, check_body, &a);
    return 0;
}

// order in which the fibers of a CTA are scheduled: 0 = thread order, otherwise pseudo-random per sweep
void sim_set_schedule_seed(uint64_t seed) { sim::schedule_seed = seed; }

// genasm_generic_kernel for any window configuration the product accepts (2 <= W <= 256, 0 <= O < W, W - O <= 128), op planes
// in shared memory (gp = 0) or in a per-CTA global scratch (gp = 1); one-warp CTAs as in the product's launch.
int sim_generic_align(int W, int O, int gp, unsigned ctas, const uint32_t *text, const uint64_t *text_start, const uint64_t *text_len,
                      const uint32_t *query, const uint64_t *query_start, const uint64_t *query_len, uint64_t n, uint32_t flags,
                      uint8_t *slab, const uint64_t *slab_off, int64_t *edit, uint64_t *ref_consumed, uint32_t *nruns, uint8_t *status,
                      uint64_t *dc_entries, uint32_t *windows, const uint32_t *order)
{
    if (W < 2 || W > 256 || O < 0 || O >= W || W - O > 128) return -1;
    unsigned long long counter = 0;
    GenericArgs a;
    sg::AlignParams &P = a.P;
    P.text = text; P.text_start = text_start; P.text_len = text_len;
    P.query = query; P.query_start = query_start; P.query_len = query_len;
    P.n = n; P.flags = flags; P.slab = slab; P.slab_off = slab_off; P.counter = &counter;
    P.edit = edit; P.ref_consumed = ref_consumed; P.nruns = nruns; P.status = status; P.dc_entries = dc_entries; P.windows = windows;
    P.order = order;
    const int NW = (W + 31) / 32, TBL = W - O;
    a.G.W = W; a.G.TBL = TBL; a.G.NWT = (TBL + 31) / 32; a.G.planes = nullptr;
    if ((size_t)sg::generic_smem_words(NW, W, TBL, gp != 0) * 4 > sizeof(sg::smem_all)) return -2;
    std::vector<uint32_t> scratch;
    if (gp) { scratch.resize((size_t)ctas * (size_t)sg::generic_plane_words(TBL)); a.G.planes = scratch.data(); }
    const bool wide = TBL > 63;
    void (*fn)(void *) = nullptr;
    switch (NW) {
        case 1: fn = generic_fn<1>(gp, wide); break;
        case 2: fn = generic_fn<2>(gp, wide); break;
        case 3: fn = generic_fn<3>(gp, wide); break;
        case 4: fn = generic_fn<4>(gp, wide); break;
        case 5: fn = generic_fn<5>(gp, wide); break;
        case 6: fn = generic_fn<6>(gp, wide); break;
        case 7: fn = generic_fn<7>(gp, wide); break;
        default: fn = generic_fn<8>(gp, wide); break;
    }
    sim::launch(ctas, 32, sg::smem_all, fn, &a);
    return 0;
}

// pack_2bit_kernel (the plain ingest kernel) over words [w_first, n_words) of the packed blob
int sim_pack_2bit(const char *ascii, uint64_t n_bases, uint32_t *packed, uint64_t n_words, uint64_t *bad_pos, uint64_t w_first, unsigned blocks)
{
    PackArgs a{ascii, n_bases, packed, n_words, (unsigned long long *)bad_pos, w_first};
    sim::launch(blocks, 256, sg::smem_all, pack_body, &a);
    return 0;
}

// the three scan passes of sg_dev_scan_runs; tmp holds n / 2048 + 2 words
int sim_scan_runs(const uint32_t *nruns, uint64_t n, uint64_t *run_off, uint64_t *tmp)
{
    const uint64_t tiles = (n + sg::kScanTile - 1) / sg::kScanTile;
    ScanArgs a{nruns, n, tmp, run_off, tiles};
    sim::launch((unsigned)tiles, sg::kScanBlock, sg::smem_all, scan1_body, &a);
    sim::launch(1, sg::kScanBlock, sg::smem_all, scan2_body, &a);
    sim::launch((unsigned)tiles, sg::kScanBlock, sg::smem_all, scan3_body, &a);
    return 0;
}

// gather_runs_kernel<group> (group = 32: a warp per alignment, 4: four lanes per alignment)
int sim_gather_runs(int group, const uint8_t *slab, const uint64_t *slab_off, const uint32_t *nruns, const uint64_t *run_off, uint64_t n,
                    uint8_t *runs, unsigned blocks)
{
    GatherArgs a{slab, slab_off, nruns, run_off, n, runs};
    if (group != 4 && group != 32) return -1;
    sim::launch(blocks, 256, sg::smem_all, group == 4 ? gather_body<4> : gather_body<32>, &a);
    return 0;
}

// One launch of genasm_delta_kernel<W, EMIT> over n alignments on `ctas` CTAs; all pointers are host memory laid out as
// sg_dev_align's device buffers (include/scrooge_b200.h).  counters_out[8]: SG_SIM_COUNT events (0 = byte stores of runs,
// 1 = word stores of runs, per lane).  Returns 0, or -1 for an unknown variant.
int sim_delta_align(int W, int emit, unsigned ctas, const uint32_t *text, const uint64_t *text_start, const uint64_t *text_len,
                    const uint32_t *query, const uint64_t *query_start, const uint64_t *query_len, uint64_t n, uint32_t flags,
                    uint8_t *slab, const uint64_t *slab_off, int64_t *edit, uint64_t *ref_consumed, uint32_t *nruns, uint8_t *status,
                    uint64_t *dc_entries, uint32_t *windows, const uint32_t *order, uint64_t *counters_out)
{
    unsigned long long counter = 0;
    sg::AlignParams P;
    P.text = text; P.text_start = text_start; P.text_len = text_len;
    P.query = query; P.query_start = query_start; P.query_len = query_len;
    P.n = n; P.flags = flags; P.slab = slab; P.slab_off = slab_off; P.counter = &counter;
    P.edit = edit; P.ref_consumed = ref_consumed; P.nruns = nruns; P.status = status; P.dc_entries = dc_entries; P.windows = windows;
    P.order = order;
    for (auto &c : sim::counters) c = 0;
    void (*fn)(void *) = nullptr;
    unsigned block = 0;
    if (W == 64) { block = sg::DeltaLayout<64>::WARPS_PER_CTA * 32; fn = emit ? body<64, 1> : body<64, 0>; static_assert(sg::DeltaLayout<64>::BYTES_PER_CTA <= sizeof(sg::smem_all), ""); }
    else if (W == 32) { block = sg::DeltaLayout<32>::WARPS_PER_CTA * 32; fn = emit ? body<32, 1> : body<32, 0>; }
    else return -1;
    if (emit != 0 && emit != 1) return -1;
    sim::launch(ctas, block, sg::smem_all, fn, &P);
    if (counters_out) for (int k = 0; k < 8; k++) counters_out[k] = sim::counters[k];
    return 0;
}

}  // extern "C"
