// sg_device_api.cu -- layer 2 of include/scrooge_b200.h: launches of the sm_100a kernels on device
// pointers and an explicit stream.  No CPU fallback: every function needs a CUDA device.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <algorithm>
#include <atomic>
#include <mutex>
#include <cuda_runtime.h>

#include "../../include/scrooge_b200.h"
#include "sg_internal.h"
#include "sg_align.cuh"
#include "sg_align_delta.cuh"
#include "sg_align_generic.cuh"
#include "sg_aux.cuh"

namespace sg {

thread_local std::string g_last_error;

int fail(int code, const std::string &msg)
{
    g_last_error = msg;
    return code;
}

int cuda_fail(cudaError_t e, const char *what)
{
    return fail(SG_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}

struct DeviceInfo {
    int sms = 0;
    int ctas_per_sm[2][2] = {{0, 0}, {0, 0}};  // row-wise kernel: [W=64 | W=32][smem forefront | TMEM forefront]
    int delta_ctas_per_sm[2][2] = {{0, 0}, {0, 0}};   // delta kernel: [W=64 | W=32][runs stored as bytes | as words]
    bool ready = false;
};
static DeviceInfo g_dev_info[64];

// Which kernel variant sg_dev_align launches.  W=64: forefront in tensor memory (16 warps per SM, alu pipe 89 %
// busy) unless SG_FOREFRONT=smem; W=32: forefront in shared memory (it is only 4 KB per warp there, 20 warps per SM)
// unless SG_FOREFRONT=tmem.  All variants are bit-identical; see SmemLayout in sg_align.cuh.
static bool use_tmem(int W)
{
    static const int forced = [] {
        const char *e = std::getenv("SG_FOREFRONT");
        if (e && std::string(e) == "smem") return 0;
        if (e && std::string(e) == "tmem") return 1;
        return -1;
    }();
    return forced >= 0 ? forced == 1 : W == 64;
}

// Which distance-calculation formulation sg_dev_align launches: "delta" (sg_align_delta.cuh, column-wise +-1 deltas,
// the default) or "rows" (sg_align.cuh, the reference's row-wise threshold vectors in G-row chunks).  Bit-identical.
static bool use_delta()
{
    static const bool v = [] {
        const char *e = std::getenv("SG_DC");
        return !(e && std::string(e) == "rows");
    }();
    return v;
}

template <int W, int EMIT> static int setup_delta_kernel(int *ctas_per_sm)
{
    using L = DeltaLayout<W>;
    auto kern = genasm_delta_kernel<W, EMIT>;
    SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES_PER_CTA));
    SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kern, L::WARPS_PER_CTA * 32, L::BYTES_PER_CTA));
    if (std::getenv("SG_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        fprintf(stderr, "[sg] delta W=%d emit=%d: occupancy %d CTAs/SM, %d regs, %d dyn smem, %d threads/CTA\n", W, EMIT, *ctas_per_sm, fa.numRegs,
                L::BYTES_PER_CTA, L::WARPS_PER_CTA * 32);
    }
    // SG_DELTA_RESERVE=1 leaves one CTA slot per SM to the HBM-bound kernels around the aligner (ingest of the next batch,
    // compaction of the previous one run beside it on other streams; bench.py --pipeline 1).  The aligner does not need the
    // slot -- W=64: 12 / 16 / 20 / 24 warps per SM run 1 M pairs in 40.6 / 38.2 / 36.49 / 36.48 ms -- but the overlap buys
    // nothing either: beside the ingest the alignment kernel slows from 36.4 to 40.0 ms (both want the alu pipe; the
    // ingest's SWAR conversion alone is ~2 ms of it) and a pipelined pass takes 42.8 ms against 41.8 ms back to back
    // (profiles/r02_pipeline_ab.md).  Off by default.
    static const bool reserve = [] { const char *e = std::getenv("SG_DELTA_RESERVE"); return e && atoi(e) != 0; }();
    if (reserve) {
        int dev = 0, smem_sm = 0;
        SG_CUDA(cudaGetDevice(&dev));
        SG_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        // the side geometry of the ingest kernel: 32 KB of tiles + its barriers + the 1 KB the system takes per CTA
        const int side = kPackSideTile * kPackSideStages + 1024 + 256;
        const int by_smem = (smem_sm - side) / (L::BYTES_PER_CTA + 1024), by_threads = (2048 - 256) / (L::WARPS_PER_CTA * 32);
        *ctas_per_sm = std::max(1, std::min(*ctas_per_sm, std::min(by_smem, by_threads)));
    }
    if (const char *e = std::getenv("SG_DELTA_CTAS")) *ctas_per_sm = std::max(1, atoi(e));  // experiment knob (the launch fails if it does not fit)
    if (*ctas_per_sm < 1) return fail(SG_ERR_CUDA, "alignment kernel does not fit on this device");
    return SG_OK;
}

template <int W, bool TMEM> static int setup_kernel(int *ctas_per_sm)
{
    using L = SmemLayout<W, TMEM>;
    auto kern = genasm_align_kernel<W, TMEM>;
    SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, L::BYTES_PER_CTA));
    SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kern, L::WARPS_PER_CTA * 32, L::BYTES_PER_CTA));
    if (std::getenv("SG_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        fprintf(stderr, "[sg] W=%d tmem=%d: occupancy %d CTAs/SM, %d regs, %zu static smem, %d dyn smem, %d threads/CTA\n", W, (int)TMEM,
                *ctas_per_sm, fa.numRegs, fa.sharedSizeBytes, L::BYTES_PER_CTA, L::WARPS_PER_CTA * 32);
    }
    if (TMEM) {
        // cudaOccupancyMaxActiveBlocksPerMultiprocessor reports 1 for any kernel that allocates tensor memory, but the
        // hardware co-schedules CTAs as long as their tcgen05.alloc requests fit in the SM's 512 columns (measured:
        // sm__warps_active shows 16 resident warps with 4 CTAs).  Compute the residency from the real limits.
        cudaFuncAttributes fa;
        SG_CUDA(cudaFuncGetAttributes(&fa, kern));
        int dev = 0, smem_sm = 0, regs_sm = 0;
        SG_CUDA(cudaGetDevice(&dev));
        SG_CUDA(cudaDeviceGetAttribute(&smem_sm, cudaDevAttrMaxSharedMemoryPerMultiprocessor, dev));
        SG_CUDA(cudaDeviceGetAttribute(&regs_sm, cudaDevAttrMaxRegistersPerMultiprocessor, dev));
        const int threads = L::WARPS_PER_CTA * 32;
        const int regs_cta = ((fa.numRegs + 7) / 8 * 8) * threads;
        int lim = 512 / L::TMEM_COLS;
        lim = std::min(lim, smem_sm / (L::BYTES_PER_CTA + (int)fa.sharedSizeBytes + 1024));
        lim = std::min(lim, regs_sm / std::max(regs_cta, 1));
        lim = std::min(lim, 2048 / threads);
        if (const char *e = std::getenv("SG_TMEM_CTAS")) lim = std::min(lim, std::max(1, atoi(e)));  // experiment knob
        *ctas_per_sm = std::max(*ctas_per_sm, lim);
    }
    if (*ctas_per_sm < 1) return fail(SG_ERR_CUDA, "alignment kernel does not fit on this device");
    return SG_OK;
}

static int device_info(DeviceInfo **out)
{
    int dev = 0;
    SG_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return fail(SG_ERR_BAD_ARG, "device index out of range");
    DeviceInfo &di = g_dev_info[dev];
    // set up once per device, whichever thread or context comes first (two contexts on one device, or device-API calls
    // from several threads, must not race on the occupancy fields); a failed set-up is retried by the next call
    static std::mutex setup_mu[64];
    static std::atomic<bool> ready[64];
    if (!ready[dev].load(std::memory_order_acquire)) {
        std::lock_guard<std::mutex> g(setup_mu[dev]);
        if (!ready[dev].load(std::memory_order_relaxed)) {
            SG_CUDA(cudaDeviceGetAttribute(&di.sms, cudaDevAttrMultiProcessorCount, dev));
            int rc = setup_kernel<64, false>(&di.ctas_per_sm[0][0]);
            if (!rc) rc = setup_kernel<64, true>(&di.ctas_per_sm[0][1]);
            if (!rc) rc = setup_kernel<32, false>(&di.ctas_per_sm[1][0]);
            if (!rc) rc = setup_kernel<32, true>(&di.ctas_per_sm[1][1]);
            if (!rc) rc = setup_delta_kernel<64, 0>(&di.delta_ctas_per_sm[0][0]);
            if (!rc) rc = setup_delta_kernel<64, 1>(&di.delta_ctas_per_sm[0][1]);
            if (!rc) rc = setup_delta_kernel<32, 0>(&di.delta_ctas_per_sm[1][0]);
            if (!rc) rc = setup_delta_kernel<32, 1>(&di.delta_ctas_per_sm[1][1]);
            if (rc) return rc;
            di.ready = true;
            ready[dev].store(true, std::memory_order_release);
        }
    }
    *out = &di;
    return SG_OK;
}

// Persistent lane count shared by both kernels, see launch_align.
static uint64_t balanced_ctas(uint64_t n, uint64_t max_ctas, uint64_t lanes_per_cta);

// The window configurations genasm_delta_kernel / genasm_align_kernel are built for (src/genasm_cpu.cpp:7-9, README.md:208);
// every other (W, O) runs on genasm_generic_kernel, and so do these two with SG_GENERIC=1 (A/B and parity tests).
static bool tuned_config(int W, int O)
{
    static const bool force_generic = [] {
        const char *e = std::getenv("SG_GENERIC");
        return e && *e && std::string(e) != "0";
    }();
    return !force_generic && ((W == 64 && O == 33) || (W == 32 && O == 17));
}

static int check_window(int W, int O)
{
    if (W < 2 || W > 256 || O < 0 || O >= W || W - O > 128)
        return fail(SG_ERR_BAD_ARG, "window configuration out of range: need 2 <= W <= 256, 0 <= O < W, W - O <= 128");
    return SG_OK;
}

// Where the general kernel keeps its op planes: shared memory, or a per-CTA scratch in global memory that stays in the L2
// cache (SG_GENERIC_PLANES=smem|global forces one).  Measured on 1 M x 10 kbp pairs (profiles/r01_window_sweep.md): global
// planes pay when the window is wide and the walk short relative to it -- 128/65 (32 KB of planes per warp, 6 warps per SM in
// shared memory): 8.2 -> 10.1 M alignments/s -- and cost when the traceback dominates (W = 64 with O <= 24, up to 126 dependent
// plane reads per window: 13-16 -> 11-12 M/s; 96/49: 11.2 -> 10.8), so only four-word windows with more than 24 KB of planes
// use them -- and every configuration whose planes exceed 40 KB per warp (W - O > 63: up to 128 KB).
static bool generic_global_planes(int W, int O)
{
    static const int forced = [] {
        const char *e = std::getenv("SG_GENERIC_PLANES");
        if (e && std::string(e) == "smem") return 0;
        if (e && std::string(e) == "global") return 1;
        return -1;
    }();
    if (forced >= 0) return forced == 1;
    const int plane_bytes = generic_plane_words(W - O) * 4;
    return plane_bytes > 40 * 1024 || ((W + 31) / 32 >= 4 && plane_bytes > 24 * 1024);   // > 40 KB: 5 warps per SM or fewer
}

template <int NW, bool GP, bool WIDE> static int generic_occupancy(int smem_bytes, int *ctas_per_sm)
{
    auto kern = genasm_generic_kernel<NW, GP, WIDE>;
    if (smem_bytes > 48 * 1024) {   // beyond the default limit: opt in (at most cudaDevAttrMaxSharedMemoryPerBlockOptin)
        if (cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes) != cudaSuccess) {
            cudaGetLastError();
            return fail(SG_ERR_CUDA, "generic alignment kernel: " + std::to_string(smem_bytes) + " bytes of shared memory per warp do not fit on this device");
        }
    }
    SG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared));
    SG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(ctas_per_sm, kern, 32, smem_bytes));
    if (std::getenv("SG_DEBUG")) {
        cudaFuncAttributes fa;
        cudaFuncGetAttributes(&fa, kern);
        fprintf(stderr, "[sg] generic NW=%d planes=%s wide=%d: occupancy %d CTAs/SM, %d regs, %d dyn smem\n", NW, GP ? "global" : "smem", (int)WIDE, *ctas_per_sm,
                fa.numRegs, smem_bytes);
    }
    if (*ctas_per_sm < 1) return fail(SG_ERR_CUDA, "generic alignment kernel does not fit on this device");
    return SG_OK;
}

static int generic_geometry(const DeviceInfo &di, int W, int O, int *ctas_per_sm, int *smem_bytes)
{
    const int NW = (W + 31) / 32;
    const bool gp = generic_global_planes(W, O);
    *smem_bytes = generic_smem_words(NW, W, W - O, gp) * 4;
    const bool wide = W - O > 63;
    int rc;
#define SG_GEN_OCC(N) (gp ? (wide ? generic_occupancy<N, true, true>(*smem_bytes, ctas_per_sm) : generic_occupancy<N, true, false>(*smem_bytes, ctas_per_sm)) \
                          : (wide ? generic_occupancy<N, false, true>(*smem_bytes, ctas_per_sm) : generic_occupancy<N, false, false>(*smem_bytes, ctas_per_sm)))
    switch (NW) {
        case 1: rc = SG_GEN_OCC(1); break;
        case 2: rc = SG_GEN_OCC(2); break;
        case 3: rc = SG_GEN_OCC(3); break;
        case 4: rc = SG_GEN_OCC(4); break;
        case 5: rc = SG_GEN_OCC(5); break;
        case 6: rc = SG_GEN_OCC(6); break;
        case 7: rc = SG_GEN_OCC(7); break;
        default: rc = SG_GEN_OCC(8); break;
    }
#undef SG_GEN_OCC
    if (rc) return rc;
    if (gp) {
        // the planes of all resident warps should stay in the L2 cache (126 MB): at most 96 MB of them, at least 8 warps per SM
        const long long per_warp = (long long)generic_plane_words(W - O) * 4;
        const long long fit = (96ll << 20) / (per_warp * (long long)di.sms);
        *ctas_per_sm = (int)std::min<long long>(*ctas_per_sm, std::max<long long>(8, fit));
    }
    return SG_OK;
}

static int launch_generic(const DeviceInfo &di, const AlignParams &P, int W, int O, cudaStream_t st)
{
    int per_sm = 0, smem = 0;
    int rc = generic_geometry(di, W, O, &per_sm, &smem);
    if (rc) return rc;
    const bool gp = generic_global_planes(W, O);
    GenericGeom G;
    G.W = W; G.TBL = W - O; G.NWT = (G.TBL + 31) / 32; G.planes = nullptr;
    const unsigned ctas = (unsigned)balanced_ctas(P.n, (uint64_t)di.sms * (uint64_t)per_sm, 32ull);
    if (gp) {   // stream-ordered scratch: launches on different streams of one device never share it
        static std::atomic<bool> pool_kept[64];
        int dev = 0;
        SG_CUDA(cudaGetDevice(&dev));
        if (dev >= 0 && dev < 64 && !pool_kept[dev].exchange(true)) {
            // keep freed scratch in the device's pool across synchronisations (the default threshold of 0 hands it back to
            // the driver at every sync, and each launch would pay a real allocation)
            cudaMemPool_t pool;
            unsigned long long keep = ~0ull;
            if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
            cudaGetLastError();
        }
        void *scratch = nullptr;
        SG_CUDA(cudaMallocAsync(&scratch, (size_t)ctas * (size_t)generic_plane_words(G.TBL) * 4, st));
        G.planes = (uint32_t *)scratch;
    }
    const bool wide = G.TBL > 63;
#define SG_GEN_LAUNCH(N)                                                                                   \
    do {                                                                                                   \
        if (gp) { if (wide) genasm_generic_kernel<N, true, true><<<ctas, 32, smem, st>>>(P, G); else genasm_generic_kernel<N, true, false><<<ctas, 32, smem, st>>>(P, G); } \
        else { if (wide) genasm_generic_kernel<N, false, true><<<ctas, 32, smem, st>>>(P, G); else genasm_generic_kernel<N, false, false><<<ctas, 32, smem, st>>>(P, G); } \
    } while (0)
    switch ((W + 31) / 32) {
        case 1: SG_GEN_LAUNCH(1); break;
        case 2: SG_GEN_LAUNCH(2); break;
        case 3: SG_GEN_LAUNCH(3); break;
        case 4: SG_GEN_LAUNCH(4); break;
        case 5: SG_GEN_LAUNCH(5); break;
        case 6: SG_GEN_LAUNCH(6); break;
        case 7: SG_GEN_LAUNCH(7); break;
        default: SG_GEN_LAUNCH(8); break;
    }
#undef SG_GEN_LAUNCH
    const cudaError_t le = cudaGetLastError();
    if (gp) cudaFreeAsync(G.planes, st);
    if (le != cudaSuccess) return cuda_fail(le, "genasm_generic_kernel launch");
    return SG_OK;
}

// Persistent lane count shared by both kernels, see launch_align.
static uint64_t balanced_ctas(uint64_t n, uint64_t max_ctas, uint64_t lanes_per_cta)
{
    const uint64_t k = (n + max_ctas * lanes_per_cta - 1) / (max_ctas * lanes_per_cta);
    const uint64_t lanes = (n + k - 1) / k;
    uint64_t ctas = (lanes + lanes_per_cta - 1) / lanes_per_cta;
    if (k >= 8 || ctas > max_ctas) ctas = max_ctas;  // many alignments per lane: the dynamic queue evens things out
    return ctas;
}

template <int W, int EMIT> static int launch_delta(const DeviceInfo &di, const AlignParams &P, cudaStream_t st)
{
    using L = DeltaLayout<W>;
    const uint64_t max_ctas = (uint64_t)di.sms * (uint64_t)di.delta_ctas_per_sm[W == 64 ? 0 : 1][EMIT];
    const uint64_t ctas = balanced_ctas(P.n, max_ctas, 32ull * L::WARPS_PER_CTA);
    genasm_delta_kernel<W, EMIT><<<(unsigned)ctas, L::WARPS_PER_CTA * 32, L::BYTES_PER_CTA, st>>>(P);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

template <int W, bool TMEM> static int launch_align(const DeviceInfo &di, const AlignParams &P, cudaStream_t st)
{
    using L = SmemLayout<W, TMEM>;
    // Persistent CTAs, a lane per alignment at a time.  With n alignments and at most `max_lanes` resident lanes every lane
    // runs k = ceil(n / max_lanes) alignments back to back; launching only ceil(n / k) lanes gives every lane the same
    // count, so that a batch of few, long alignments (100 kbp reads: 1.3 per resident lane) does not end with most of the
    // device idle while a third of the lanes run their second alignment.
    const uint64_t lanes_per_cta = 32ull * L::WARPS_PER_CTA;
    const uint64_t max_ctas = (uint64_t)di.sms * (uint64_t)di.ctas_per_sm[W == 64 ? 0 : 1][TMEM ? 1 : 0];
    const uint64_t k = (P.n + max_ctas * lanes_per_cta - 1) / (max_ctas * lanes_per_cta);
    const uint64_t lanes = (P.n + k - 1) / k;
    uint64_t ctas = (lanes + lanes_per_cta - 1) / lanes_per_cta;
    if (k >= 8 || ctas > max_ctas) ctas = max_ctas;  // many alignments per lane: the dynamic queue evens things out
    genasm_align_kernel<W, TMEM><<<(unsigned)ctas, L::WARPS_PER_CTA * 32, L::BYTES_PER_CTA, st>>>(P);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

template <int TILE, int STAGES>
static int launch_pack_bulk(const DeviceInfo &di, const char *d_ascii, uint64_t n_bases, uint32_t *d_packed, uint64_t *d_bad_pos, cudaStream_t st,
                            uint64_t *done_bases)
{
    static std::once_flag attr_once[64];
    int dev = 0;
    SG_CUDA(cudaGetDevice(&dev));
    constexpr int smem = TILE * STAGES;
    cudaError_t attr_rc = cudaSuccess;
    if (dev >= 0 && dev < 64)
        std::call_once(attr_once[dev], [&] { attr_rc = cudaFuncSetAttribute(pack_2bit_bulk_kernel<TILE, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem); });
    SG_CUDA(attr_rc);
    const uint64_t n_tiles = n_bases / (uint64_t)TILE;
    const int per_sm = std::max(1, (227 * 1024) / (smem + 1024));
    const int blocks = (int)std::min<uint64_t>(n_tiles, (uint64_t)di.sms * (uint64_t)per_sm);
    pack_2bit_bulk_kernel<TILE, STAGES><<<blocks, 256, smem, st>>>(d_ascii, n_tiles, d_packed, (unsigned long long *)d_bad_pos);
    SG_CUDA(cudaGetLastError());
    *done_bases = n_tiles * (uint64_t)TILE;
    return SG_OK;
}

}  // namespace sg

using namespace sg;

extern "C" {

const char *sg_last_error(void) { return g_last_error.c_str(); }

int sg_version(void) { return 1; }

int sg_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

uint64_t sg_packed_words(uint64_t n_bases) { return (n_bases + 15ull) / 16ull + 8ull; }

int sg_dev_pack_2bit(const char *d_ascii, uint64_t n_bases, uint32_t *d_packed, uint64_t *d_bad_pos, void *stream)
{
    return sg_dev_pack_2bit_ex(d_ascii, n_bases, d_packed, d_bad_pos, 0, stream);
}

int sg_dev_pack_2bit_ex(const char *d_ascii, uint64_t n_bases, uint32_t *d_packed, uint64_t *d_bad_pos, uint32_t flags, void *stream)
{
    if (!d_packed || !d_bad_pos || (!d_ascii && n_bases)) return fail(SG_ERR_BAD_ARG, "sg_dev_pack_2bit: null pointer");
    if (((uintptr_t)d_ascii & 15u) != 0) return fail(SG_ERR_BAD_ARG, "sg_dev_pack_2bit: d_ascii must be 16-byte aligned");
    DeviceInfo *di;
    int rc = device_info(&di);
    if (rc) return rc;
    const uint64_t n_words = sg_packed_words(n_bases);  // includes zeroed padding words the aligner may read
    // the whole tiles of the blob through the bulk-copy-staged kernel, the tail through the plain one (SG_PACK=plain:
    // everything through the plain one)
    static const bool bulk = [] { const char *e = std::getenv("SG_PACK"); return !(e && std::string(e) == "plain"); }();
    uint64_t done_bases = 0;
    if (bulk && (flags & SG_PACK_SIDE) && n_bases >= (uint64_t)kPackSideTile)
        rc = launch_pack_bulk<kPackSideTile, kPackSideStages>(*di, d_ascii, n_bases, d_packed, d_bad_pos, (cudaStream_t)stream, &done_bases);
    else if (bulk && n_bases >= (uint64_t)kPackTile)
        rc = launch_pack_bulk<kPackTile, kPackStages>(*di, d_ascii, n_bases, d_packed, d_bad_pos, (cudaStream_t)stream, &done_bases);
    if (rc) return rc;
    const uint64_t rest_words = n_words - done_bases / 16ull;
    const uint64_t want = (rest_words + 255ull) / 256ull;
    const int blocks = (int)std::min<uint64_t>(want, (uint64_t)di->sms * 16ull);
    if (rest_words)
        pack_2bit_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(d_ascii, n_bases, d_packed, n_words, (unsigned long long *)d_bad_pos,
                                                                  done_bases / 16ull);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_default_overlap(int W) { return W / 2 + 1 < W - 1 ? W / 2 + 1 : W - 1; }

int sg_dev_align_geometry(int W, int *warps_per_sm, int *smem_per_warp, int *num_sms)
{
    if (W != 64 && W != 32) return fail(SG_ERR_BAD_ARG, "W must be 64 or 32");
    return sg_dev_align_geometry_wo(W, sg_default_overlap(W), warps_per_sm, smem_per_warp, num_sms);
}

int sg_dev_align_geometry_wo(int W, int O, int *warps_per_sm, int *smem_per_warp, int *num_sms)
{
    int rc = check_window(W, O);
    if (rc) return rc;
    DeviceInfo *di;
    rc = device_info(&di);
    if (rc) return rc;
    if (!tuned_config(W, O)) {
        int per_sm = 0, smem = 0;
        rc = generic_geometry(*di, W, O, &per_sm, &smem);
        if (rc) return rc;
        if (warps_per_sm) *warps_per_sm = per_sm;
        if (smem_per_warp) *smem_per_warp = smem;
        if (num_sms) *num_sms = di->sms;
        return SG_OK;
    }
    if (use_delta()) {
        if (warps_per_sm) *warps_per_sm = di->delta_ctas_per_sm[W == 64 ? 0 : 1][0] * DeltaLayout<64>::WARPS_PER_CTA;
        if (smem_per_warp) *smem_per_warp = W == 64 ? DeltaLayout<64>::BYTES_PER_WARP : DeltaLayout<32>::BYTES_PER_WARP;
        if (num_sms) *num_sms = di->sms;
        return SG_OK;
    }
    const bool t = use_tmem(W);
    if (warps_per_sm) *warps_per_sm = di->ctas_per_sm[W == 64 ? 0 : 1][t ? 1 : 0] * (t ? 4 : 1);
    if (smem_per_warp)
        *smem_per_warp = W == 64 ? (t ? SmemLayout<64, true>::BYTES_PER_WARP : SmemLayout<64, false>::BYTES_PER_WARP)
                                 : (t ? SmemLayout<32, true>::BYTES_PER_WARP : SmemLayout<32, false>::BYTES_PER_WARP);
    if (num_sms) *num_sms = di->sms;
    return SG_OK;
}

int sg_dev_align(int W, const uint32_t *d_text, const uint64_t *d_text_start, const uint64_t *d_text_len,
                 const uint32_t *d_query, const uint64_t *d_query_start, const uint64_t *d_query_len,
                 uint64_t n, uint32_t flags, uint8_t *d_slab, const uint64_t *d_slab_off,
                 uint64_t *d_counter, int64_t *d_edit, uint64_t *d_ref_consumed, uint32_t *d_nruns,
                 uint8_t *d_status, uint64_t *d_dc_entries, uint32_t *d_windows, void *stream)
{
    if (W != 64 && W != 32) return fail(SG_ERR_BAD_ARG, "W must be 64 or 32");
    return sg_dev_align_wo(W, sg_default_overlap(W), d_text, d_text_start, d_text_len, d_query, d_query_start, d_query_len, n, flags,
                           d_slab, d_slab_off, d_counter, d_edit, d_ref_consumed, d_nruns, d_status, d_dc_entries, d_windows, stream);
}

int sg_dev_align_wo(int W, int O, const uint32_t *d_text, const uint64_t *d_text_start, const uint64_t *d_text_len,
                    const uint32_t *d_query, const uint64_t *d_query_start, const uint64_t *d_query_len,
                    uint64_t n, uint32_t flags, uint8_t *d_slab, const uint64_t *d_slab_off,
                    uint64_t *d_counter, int64_t *d_edit, uint64_t *d_ref_consumed, uint32_t *d_nruns,
                    uint8_t *d_status, uint64_t *d_dc_entries, uint32_t *d_windows, void *stream)
{
    return sg_dev_align_ordered(W, O, d_text, d_text_start, d_text_len, d_query, d_query_start, d_query_len, n, flags, d_slab, d_slab_off,
                                d_counter, d_edit, d_ref_consumed, d_nruns, d_status, d_dc_entries, d_windows, nullptr, stream);
}

int sg_dev_align_ordered(int W, int O, const uint32_t *d_text, const uint64_t *d_text_start, const uint64_t *d_text_len,
                         const uint32_t *d_query, const uint64_t *d_query_start, const uint64_t *d_query_len,
                         uint64_t n, uint32_t flags, uint8_t *d_slab, const uint64_t *d_slab_off,
                         uint64_t *d_counter, int64_t *d_edit, uint64_t *d_ref_consumed, uint32_t *d_nruns,
                         uint8_t *d_status, uint64_t *d_dc_entries, uint32_t *d_windows, const uint32_t *d_order, void *stream)
{
    {
        const int wrc = check_window(W, O);
        if (wrc) return wrc;
    }
    if (n == 0) return SG_OK;
    if (d_order && n > 0xFFFFFFFFull) return fail(SG_ERR_BAD_ARG, "sg_dev_align_ordered: more than 2^32 alignments in one launch");
    if (!d_text || !d_text_start || !d_text_len || !d_query || !d_query_start || !d_query_len || !d_counter ||
        !d_edit || !d_ref_consumed || !d_nruns || !d_status)
        return fail(SG_ERR_BAD_ARG, "sg_dev_align: null pointer");
    if (!(flags & SG_FLAG_DISTANCE_ONLY) && (!d_slab || !d_slab_off))
        return fail(SG_ERR_BAD_ARG, "sg_dev_align: CIGAR output needs d_slab and d_slab_off");
    DeviceInfo *di;
    int rc = device_info(&di);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    SG_CUDA(cudaMemsetAsync(d_counter, 0, sizeof(uint64_t), st));
    AlignParams P;
    P.text = d_text; P.text_start = d_text_start; P.text_len = d_text_len;
    P.query = d_query; P.query_start = d_query_start; P.query_len = d_query_len;
    P.n = n; P.flags = flags; P.slab = d_slab; P.slab_off = d_slab_off;
    P.counter = (unsigned long long *)d_counter;
    P.edit = d_edit; P.ref_consumed = d_ref_consumed; P.nruns = d_nruns; P.status = d_status; P.dc_entries = d_dc_entries; P.windows = d_windows;
    P.order = d_order;
    if (!tuned_config(W, O)) return launch_generic(*di, P, W, O, st);
    if (use_delta()) {
        // runs as whole words: only with CIGAR output, and only on the caller's promise of 4-byte aligned slots
        const bool words = (flags & SG_FLAG_RUN_WORDS) && !(flags & SG_FLAG_DISTANCE_ONLY);
        if (words && ((uintptr_t)d_slab & 3u) != 0) return fail(SG_ERR_BAD_ARG, "sg_dev_align: SG_FLAG_RUN_WORDS needs a 4-byte aligned d_slab (and slab offsets)");
        if (words) return W == 64 ? launch_delta<64, 1>(*di, P, st) : launch_delta<32, 1>(*di, P, st);
        return W == 64 ? launch_delta<64, 0>(*di, P, st) : launch_delta<32, 0>(*di, P, st);
    }
    if (use_tmem(W)) return W == 64 ? launch_align<64, true>(*di, P, st) : launch_align<32, true>(*di, P, st);
    return W == 64 ? launch_align<64, false>(*di, P, st) : launch_align<32, false>(*di, P, st);
}

#ifdef SG_STATS
int sg_dev_debug_stats(uint64_t *out4, int reset)
{
    SG_CUDA(cudaMemcpyFromSymbol(out4, g_delta_stats, 32));
    if (reset) { uint64_t z[4] = {0, 0, 0, 0}; SG_CUDA(cudaMemcpyToSymbol(g_delta_stats, z, 32)); }
    return SG_OK;
}
#endif

uint64_t sg_scan_tmp_bytes(uint64_t n) { return ((n + kScanTile - 1) / kScanTile + 1) * sizeof(uint64_t); }

int sg_dev_scan_runs(const uint32_t *d_nruns, uint64_t n, uint64_t *d_run_off, void *d_scan_tmp, void *stream)
{
    if (!d_run_off) return fail(SG_ERR_BAD_ARG, "sg_dev_scan_runs: null pointer");
    cudaStream_t st = (cudaStream_t)stream;
    if (n == 0) {
        SG_CUDA(cudaMemsetAsync(d_run_off, 0, sizeof(uint64_t), st));
        return SG_OK;
    }
    if (!d_nruns || !d_scan_tmp) return fail(SG_ERR_BAD_ARG, "sg_dev_scan_runs: null pointer");
    const uint64_t tiles = (n + kScanTile - 1) / kScanTile;
    uint64_t *tmp = (uint64_t *)d_scan_tmp;
    scan_tile_sums_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(d_nruns, n, tmp);
    scan_tile_offsets_kernel<<<1, kScanBlock, 0, st>>>(tmp, tiles);
    scan_finish_kernel<<<(unsigned)tiles, kScanBlock, 0, st>>>(d_nruns, n, tmp, d_run_off);
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

int sg_dev_gather_runs(const uint8_t *d_slab, const uint64_t *d_slab_off, const uint32_t *d_nruns,
                       const uint64_t *d_run_off, uint64_t n, uint8_t *d_runs, void *stream)
{
    return sg_dev_gather_runs_sized(d_slab, d_slab_off, d_nruns, d_run_off, n, d_runs, 0, stream);
}

int sg_dev_gather_runs_sized(const uint8_t *d_slab, const uint64_t *d_slab_off, const uint32_t *d_nruns,
                             const uint64_t *d_run_off, uint64_t n, uint8_t *d_runs, uint64_t runs_per_alignment_hint, void *stream)
{
    if (n == 0) return SG_OK;
    if (!d_slab || !d_slab_off || !d_nruns || !d_run_off || !d_runs)
        return fail(SG_ERR_BAD_ARG, "sg_dev_gather_runs: null pointer");
    DeviceInfo *di;
    int rc = device_info(&di);
    if (rc) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    // a warp per alignment when alignments have hundreds of runs or more (or nothing is known: hint 0); four lanes per
    // alignment when they have a handful (10 M short reads: the warp-wide version leaves 28 of 32 lanes idle)
    if (runs_per_alignment_hint != 0 && runs_per_alignment_hint <= 1024) {
        const uint64_t want = (n * 4ull + 255ull) / 256ull;
        const unsigned blocks = (unsigned)std::min<uint64_t>(want, (uint64_t)di->sms * 32ull);
        gather_runs_kernel<4><<<blocks, 256, 0, st>>>(d_slab, d_slab_off, d_nruns, d_run_off, n, d_runs);
    } else {
        const uint64_t want = (n * 32ull + 255ull) / 256ull;
        const unsigned blocks = (unsigned)std::min<uint64_t>(want, (uint64_t)di->sms * 32ull);
        gather_runs_kernel<32><<<blocks, 256, 0, st>>>(d_slab, d_slab_off, d_nruns, d_run_off, n, d_runs);
    }
    SG_CUDA(cudaGetLastError());
    return SG_OK;
}

}  // extern "C"
