// sim_runtime.cpp -- the fiber scheduler behind tests/sim/shim/cuda_runtime.h.  TEST INFRASTRUCTURE.
// One CTA at a time, its threads as ucontext fibers run round-robin by one host thread; a fiber leaves the CPU only at
// a rendezvous (warp vote / shuffle, CTA barrier), so everything between two rendezvous is atomic -- a legal schedule of
// the CUDA model (independent thread scheduling promises no more).
#include <ucontext.h>
#include <stdlib.h>
#include <stdio.h>
#include <vector>
#include <utility>
#include "sim_runtime.h"

namespace sim {

ThreadCtx *cur = nullptr;
uint3 block_idx, block_dim, grid_dim;
unsigned char *smem_base = nullptr;
uint64_t counters[8];
uint64_t schedule_seed = 0;

namespace {
struct Fiber {
    ThreadCtx t;
    ucontext_t ctx;
    char *stack = nullptr;
    bool done = false;
};
struct Warp {
    unsigned gen = 0, arrived = 0, votes = 0, result = 0, alive = 0, exited = 0;
    uint64_t xch[32];
};
constexpr size_t kStack = 512 * 1024;
std::vector<Fiber> fibers;
std::vector<Warp> warps;
unsigned cta_gen, cta_arrived, cta_alive;
ucontext_t sched_ctx;
void (*g_body)(void *);
void *g_arg;

void yield() { swapcontext(&static_cast<Fiber *>(static_cast<void *>(cur))->ctx, &sched_ctx); }

void trampoline()
{
    g_body(g_arg);
    Fiber *f = static_cast<Fiber *>(static_cast<void *>(cur));   // ThreadCtx is the first member
    f->done = true;
    Warp &w = warps[f->t.tid.x >> 5];
    w.alive--;
    w.exited |= 1u << (f->t.tid.x & 31u);
    if (w.alive && w.arrived == w.alive) {   // the others were waiting for this lane only
        w.result = w.votes; w.votes = 0; w.arrived = 0; w.gen++;
    }
    cta_alive--;
    if (cta_alive && cta_arrived == cta_alive) { cta_arrived = 0; cta_gen++; }
    swapcontext(&f->ctx, &sched_ctx);
}
}  // namespace

unsigned ballot(bool pred)
{
    const unsigned lane = cur->tid.x & 31u;
    Warp &w = warps[cur->tid.x >> 5];
    const unsigned g = w.gen;
    if (pred) w.votes |= 1u << lane;
    if (++w.arrived == w.alive) {
        w.result = w.votes;
        w.votes = 0; w.arrived = 0; w.gen++;
    } else {
        while (w.gen == g) yield();
    }
    return w.result;
}

unsigned alive_mask() { return ~warps[cur->tid.x >> 5].exited; }

uint64_t exchange(uint64_t v, int src_lane)
{
    Warp &w = warps[cur->tid.x >> 5];
    w.xch[cur->tid.x & 31u] = v;
    (void)ballot(true);
    const uint64_t r = w.xch[src_lane & 31];
    (void)ballot(true);   // nobody overwrites its slot before everybody has read
    return r;
}

void cta_barrier()
{
    const unsigned g = cta_gen;
    if (++cta_arrived == cta_alive) { cta_arrived = 0; cta_gen++; }
    else while (cta_gen == g) yield();
}

void launch(unsigned grid, unsigned block, void *smem, void (*body)(void *), void *arg)
{
    if (block % 32u) { fprintf(stderr, "sim::launch: block size must be a multiple of 32\n"); abort(); }
    g_body = body;
    g_arg = arg;
    grid_dim = uint3{grid, 1, 1};
    block_dim = uint3{block, 1, 1};
    smem_base = static_cast<unsigned char *>(smem);
    if (fibers.size() < block) fibers.resize(block);
    for (unsigned t = 0; t < block; t++)
        if (!fibers[t].stack) fibers[t].stack = (char *)malloc(kStack);
    for (unsigned b = 0; b < grid; b++) {
        block_idx = uint3{b, 0, 0};
        warps.assign(block / 32, Warp());
        for (auto &w : warps) w.alive = 32;
        cta_gen = 0; cta_arrived = 0; cta_alive = block;
        for (unsigned t = 0; t < block; t++) {
            Fiber &f = fibers[t];
            f.t.tid = uint3{t, 0, 0};
            f.done = false;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
        }
        // round-robin sweeps over the fibers; with schedule_seed != 0 every sweep visits them in a fresh pseudo-random order, so
        // that lanes reach the work queue and the rendezvous in orders a GPU might produce as well
        std::vector<unsigned> order(block);
        for (unsigned t = 0; t < block; t++) order[t] = t;
        uint64_t rs = schedule_seed * 0x9E3779B97F4A7C15ull + b + 1;
        unsigned left = block;
        while (left) {
            left = 0;
            if (schedule_seed)
                for (unsigned t = block - 1; t > 0; t--) {
                    rs ^= rs << 13; rs ^= rs >> 7; rs ^= rs << 17;
                    std::swap(order[t], order[rs % (t + 1)]);
                }
            for (unsigned k = 0; k < block; k++) {
                Fiber &f = fibers[order[k]];
                if (f.done) continue;
                cur = &f.t;
                swapcontext(&sched_ctx, &f.ctx);
                if (!f.done) left++;
            }
        }
    }
    cur = nullptr;
    smem_base = nullptr;
}

}  // namespace sim
