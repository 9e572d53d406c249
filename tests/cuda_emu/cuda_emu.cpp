/* cuda_emu.cpp -- fiber scheduler for the TEST-ONLY CUDA emulation (see cuda_emu.h). */
#include "cuda_emu.h"

dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace cuemu {

CtaState g_cta;
static const size_t kStack = 192 * 1024;

[[noreturn]] void die(const char *msg)
{
    fprintf(stderr, "cuda_emu: %s (block %u thread %u)\n", msg, blockIdx.x, threadIdx.x);
    abort();
}

void yield()
{
    Fiber *f = g_cta.cur;
    swapcontext(&f->ctx, &g_cta.sched);
    threadIdx = dim3(f->tid);            /* restored on resume */
}

void cta_barrier()
{
    CtaState &c = g_cta;
    unsigned gen = c.bar_gen;
    c.bar_arrived++;
    if (c.bar_arrived >= c.live) {
        c.bar_arrived = 0; c.bar_gen++; c.progress = true;
        return;
    }
    while (c.bar_gen == gen) yield();
    c.progress = true;
}

void warp_barrier(unsigned mask)
{
    CtaState &c = g_cta;
    WarpState &w = c.warps[threadIdx.x >> 5];
    /* lanes beyond blockDim do not exist */
    unsigned first = threadIdx.x & ~31u;
    unsigned have = c.nthreads - first >= 32 ? 0xffffffffu : ((1u << (c.nthreads - first)) - 1u);
    unsigned need = (unsigned)__builtin_popcount(mask & have);
    if (!((mask >> (threadIdx.x & 31u)) & 1u)) die("lane not in its own sync mask");
    unsigned gen = w.gen;
    w.arrived++;
    if (w.arrived >= need) {
        w.arrived = 0; w.gen++; c.progress = true;
        return;
    }
    while (w.gen == gen) yield();
    c.progress = true;
}

static void fiber_entry()
{
    Fiber *f = g_cta.cur;
    threadIdx = dim3(f->tid);
    g_cta.body();
    f->done = true;
    g_cta.live--;
    g_cta.progress = true;
    /* a thread that exits counts as arrived for a pending CTA barrier */
    if (g_cta.live && g_cta.bar_arrived >= g_cta.live) {
        g_cta.bar_arrived = 0; g_cta.bar_gen++;
    }
    swapcontext(&f->ctx, &g_cta.sched);
}

void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> &body)
{
    CtaState &c = g_cta;
    if (block.y != 1 || block.z != 1 || grid.y != 1 || grid.z != 1) die("only 1-D launches emulated");
    unsigned T = block.x;
    gridDim = grid; blockDim = block;
    c.body = body;
    c.nthreads = T;
    if (c.fibers.size() < T) {
        size_t old = c.fibers.size();
        c.fibers.resize(T);
        for (size_t i = old; i < T; i++) c.fibers[i].stack = (char *)malloc(kStack);
    }
    c.warps.assign((T + 31) / 32, WarpState());
    unsigned char *dyn = smem ? (unsigned char *)aligned_alloc(1024, (smem + 1023) & ~(size_t)1023) : nullptr;
    c.dyn_smem = dyn;

    for (unsigned b = 0; b < grid.x; b++) {
        blockIdx = dim3(b);
        if (dyn) memset(dyn, 0xcd, smem);          /* shared memory starts as garbage */
        c.live = T; c.bar_arrived = 0; c.bar_gen = 0;
        for (auto &w : c.warps) { w.arrived = 0; w.gen = 0; }
        for (unsigned t = 0; t < T; t++) {
            Fiber &f = c.fibers[t];
            f.done = false; f.tid = t;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack;
            f.ctx.uc_stack.ss_size = kStack;
            f.ctx.uc_link = &c.sched;
            makecontext(&f.ctx, (void (*)())fiber_entry, 0);
        }
        while (c.live) {
            c.progress = false;
            for (unsigned t = 0; t < T; t++) {
                Fiber &f = c.fibers[t];
                if (f.done) continue;
                c.cur = &f;
                threadIdx = dim3(t);
                swapcontext(&c.sched, &f.ctx);
            }
            if (!c.progress && c.live) die("deadlock: threads wait at a barrier/collective that cannot complete");
        }
    }
    free(dyn);
    c.dyn_smem = nullptr;
}

} /* namespace cuemu */
