// cuda_emu.cpp -- cooperative SIMT emulator behind TC_EMU builds (tests only).
//
// Blocks run one after another on the calling OS thread.  Inside a block every
// CUDA thread is a ucontext fiber; fibers run round-robin and yield at
// __syncthreads(), __syncwarp() and warp shuffles/ballots, which are
// implemented as block- / warp-level generation barriers.
#ifdef TC_EMU
#include "tc_rt.h"

#include <ucontext.h>
#include <stdio.h>
#include <vector>

uint3_emu threadIdx, blockIdx;
dim3 blockDim, gridDim;
unsigned char *tc_emu_dyn_smem = nullptr;

namespace {

struct Fiber {
    ucontext_t ctx;
    unsigned char *stack = nullptr;
    bool done = false;
    uint3_emu tid;
};

const size_t kStack = 256 * 1024;

ucontext_t g_sched;
std::vector<Fiber> g_fibers;
int g_cur = -1;
bool g_in_fiber = false;
const std::function<void()> *g_body = nullptr;

// block barrier
int g_live = 0, g_arrived = 0;
unsigned g_gen = 0;
// warp barriers
std::vector<int> g_warp_live, g_warp_arrived;
std::vector<unsigned> g_warp_gen;
std::vector<uint64_t> g_exch;
std::vector<int> g_pred;

inline int lin_tid() { return (int)(threadIdx.x + blockDim.x * (threadIdx.y + blockDim.y * threadIdx.z)); }

void yield_fiber()
{
    Fiber &f = g_fibers[g_cur];
    swapcontext(&f.ctx, &g_sched);
    threadIdx = f.tid;
}

void fiber_entry()
{
    (*g_body)();
    Fiber &f = g_fibers[g_cur];
    f.done = true;
    int w = g_cur / 32;
    g_live--;
    g_warp_live[w]--;
    // release barriers that were only waiting for this thread
    if (g_live > 0 && g_arrived == g_live) { g_arrived = 0; g_gen++; }
    if (g_warp_live[w] > 0 && g_warp_arrived[w] == g_warp_live[w]) { g_warp_arrived[w] = 0; g_warp_gen[w]++; }
    swapcontext(&f.ctx, &g_sched);
}

void warp_barrier()
{
    int w = lin_tid() / 32;
    unsigned gen = g_warp_gen[w];
    if (++g_warp_arrived[w] == g_warp_live[w]) { g_warp_arrived[w] = 0; g_warp_gen[w]++; return; }
    while (g_warp_gen[w] == gen) yield_fiber();
}

}  // namespace

void __syncthreads()
{
    if (!g_in_fiber) { fprintf(stderr, "tc_emu: __syncthreads in a NOSYNC launch\n"); abort(); }
    unsigned gen = g_gen;
    if (++g_arrived == g_live) { g_arrived = 0; g_gen++; return; }
    while (g_gen == gen) yield_fiber();
}

void __syncwarp(unsigned)
{
    if (!g_in_fiber) return;
    warp_barrier();
}

uint64_t tc_emu_shfl_u64(uint64_t v, int arg, int mode, int width)
{
    if (!g_in_fiber) { fprintf(stderr, "tc_emu: shuffle in a NOSYNC launch\n"); abort(); }
    int t = lin_tid();
    int lane = t & 31, base = t - lane;
    g_exch[t] = v;
    warp_barrier();
    int seg = lane - (lane % width);
    int src;
    switch (mode) {
    case 0: src = seg + (arg % width); break;                                  // idx
    case 1: src = lane ^ arg; if (src < seg || src >= seg + width) src = lane; break;  // xor
    case 2: src = lane - arg; if (src < seg) src = lane; break;                // up
    default: src = lane + arg; if (src >= seg + width) src = lane; break;      // down
    }
    uint64_t out = v;
    int st = base + src;
    if (st < (int)g_fibers.size() && !g_fibers[st].done) out = g_exch[st];
    warp_barrier();
    return out;
}

unsigned __ballot_sync(unsigned, int pred)
{
    if (!g_in_fiber) { fprintf(stderr, "tc_emu: ballot in a NOSYNC launch\n"); abort(); }
    int t = lin_tid();
    int lane = t & 31, base = t - lane;
    g_pred[t] = pred != 0;
    warp_barrier();
    unsigned out = 0;
    for (int l = 0; l < 32; l++) {
        int st = base + l;
        if (st < (int)g_fibers.size() && !g_fibers[st].done && g_pred[st]) out |= 1u << l;
    }
    (void)lane;
    warp_barrier();
    return out;
}

void tc_emu_run_grid(dim3 grid, dim3 block, size_t smem, bool nosync,
                     const std::function<void()> &body)
{
    gridDim = grid;
    blockDim = block;
    size_t nthreads = (size_t)block.x * block.y * block.z;
    std::vector<unsigned char> dyn(smem + 64);
    tc_emu_dyn_smem = dyn.data();
    if (!nosync && g_fibers.size() < nthreads) {
        size_t old = g_fibers.size();
        g_fibers.resize(nthreads);
        for (size_t i = old; i < nthreads; i++) g_fibers[i].stack = (unsigned char *)malloc(kStack);
    }
    for (unsigned bz = 0; bz < grid.z; bz++)
    for (unsigned by = 0; by < grid.y; by++)
    for (unsigned bx = 0; bx < grid.x; bx++) {
        blockIdx.x = bx; blockIdx.y = by; blockIdx.z = bz;
        if (nosync) {
            g_in_fiber = false;
            for (unsigned tz = 0; tz < block.z; tz++)
            for (unsigned ty = 0; ty < block.y; ty++)
            for (unsigned tx = 0; tx < block.x; tx++) {
                threadIdx.x = tx; threadIdx.y = ty; threadIdx.z = tz;
                body();
            }
            continue;
        }
        g_in_fiber = true;
        g_body = &body;
        size_t nw = (nthreads + 31) / 32;
        g_warp_live.assign(nw, 0);
        g_warp_arrived.assign(nw, 0);
        g_warp_gen.assign(nw, 0);
        g_exch.assign(nthreads, 0);
        g_pred.assign(nthreads, 0);
        g_live = (int)nthreads;
        g_arrived = 0;
        size_t i = 0;
        for (unsigned tz = 0; tz < block.z; tz++)
        for (unsigned ty = 0; ty < block.y; ty++)
        for (unsigned tx = 0; tx < block.x; tx++, i++) {
            Fiber &f = g_fibers[i];
            f.done = false;
            f.tid.x = tx; f.tid.y = ty; f.tid.z = tz;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &g_sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
            g_warp_live[i / 32]++;
        }
        size_t remaining = nthreads;
        while (remaining) {
            remaining = 0;
            for (size_t k = 0; k < nthreads; k++) {
                Fiber &f = g_fibers[k];
                if (f.done) continue;
                g_cur = (int)k;
                threadIdx = f.tid;
                swapcontext(&g_sched, &f.ctx);
                if (!f.done) remaining++;
            }
        }
        g_in_fiber = false;
    }
    tc_emu_dyn_smem = nullptr;
}
#endif  // TC_EMU
