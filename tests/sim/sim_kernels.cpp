// sim_kernels.cpp -- the alignment kernel's SOURCE (scrooge_b200/csrc/sg_align_delta.cuh) compiled for the host and run
// thread by thread on the fiber scheduler of sim_runtime.cpp.  TEST INFRASTRUCTURE: built and loaded by
// tests/test_kernel_sim.py only; nothing under scrooge_b200/ knows it exists.
#include <vector>
#include "sim_runtime.h"
#include "../../scrooge_b200/csrc/sg_align_delta.cuh"

namespace sg {
__attribute__((aligned(16))) uint32_t smem_all[(227 * 1024) / 4];   // what `extern __shared__ uint32_t smem_all[]` in the kernels resolves to
}

namespace {
template <int W, int EMIT> void body(void *arg) { sg::genasm_delta_kernel<W, EMIT>(*static_cast<const sg::AlignParams *>(arg)); }
}

extern "C" {

// One launch of genasm_delta_kernel<W, EMIT> over n alignments on `ctas` CTAs; all pointers are host memory laid out as
// sg_dev_align's device buffers (include/scrooge_b200.h).  counters_out[8]: SG_SIM_COUNT events (0 = byte stores of runs,
// 1 = word stores, 2 = 64-bit stores of runs, per lane).  Returns 0, or -1 for an unknown variant.
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
    P.k_one = 1u; P.k_two = 2u; P.k_4 = 4u; P.k_16 = 16u; P.k_256 = 256u;
    for (int c = 0; c < 16; c++) P.k_sel[c] = 1u << (30 - 2 * c);
    for (auto &c : sim::counters) c = 0;
    void (*fn)(void *) = nullptr;
    unsigned block = 0;
    if (W == 64) { block = sg::DeltaLayout<64>::WARPS_PER_CTA * 32; fn = emit == 2 ? body<64, 2> : emit ? body<64, 1> : body<64, 0>; static_assert(sg::DeltaLayout<64>::BYTES_PER_CTA <= sizeof(sg::smem_all), ""); }
    else if (W == 32) { block = sg::DeltaLayout<32>::WARPS_PER_CTA * 32; fn = emit == 2 ? body<32, 2> : emit ? body<32, 1> : body<32, 0>; }
    else return -1;
    if (emit < 0 || emit > 2) return -1;
    sim::launch(ctas, block, sg::smem_all, fn, &P);
    if (counters_out) for (int k = 0; k < 8; k++) counters_out[k] = sim::counters[k];
    return 0;
}

}  // extern "C"
