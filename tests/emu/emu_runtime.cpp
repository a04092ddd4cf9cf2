// tests/emu/emu_runtime.cpp -- the fiber scheduler behind tests/emu/cuda_runtime.h (SIMT emulation, TEST INFRASTRUCTURE ONLY).
#include <cuda_runtime.h>
#include <setjmp.h>
#include <sys/mman.h>
#include <time.h>
#include <ucontext.h>

#include <vector>

cudaError_t cudaEventRecord(cudaEvent_t e, cudaStream_t) {
    timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts);
    *(double*)e = (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
    return cudaSuccess;
}

namespace emu {
namespace {
constexpr size_t STACK_BYTES = 512 << 10;

struct Fiber {
    ucontext_t ctx;
    char* stack = nullptr;
    bool done = false, started = false;
    jmp_buf jb;                             // switches after the first entry use _setjmp / _longjmp: no signal-mask system call
    int waitKind = 0;                       // 0 runnable | 1 waits for its warp's collective | 2 waits at the block barrier
    unsigned waitGen = 0;
    ThreadInfo info;
};
struct Warp {
    unsigned arrived = 0, exited = 0, gen = 0;
    unsigned pendingMask = 0;               // the collective the arrived lanes wait for (completed by the last arrival or by an exit)
    CollectiveFn pendingFn = nullptr;
    unsigned long long vals[32], out[32];
    unsigned aux[32];
};
struct Block {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    unsigned live = 0, syncArrived = 0, syncGen = 0;
    unsigned long long events = 0;          // collectives completed + fibers finished: progress indicator
};
Block* g_block = nullptr;
Fiber* g_cur = nullptr;
ucontext_t g_sched;
jmp_buf g_schedJb;
const std::function<void()>* g_body = nullptr;
ThreadInfo g_hostInfo;

void yield() { if (!_setjmp(g_cur->jb)) _longjmp(g_schedJb, 1); }

void release_barrier_if_complete(Block& b) {
    if (b.syncArrived > 0 && b.syncArrived >= b.live) { b.syncArrived = 0; b.syncGen++; b.events++; }
}
void complete_if_ready(Warp& w, unsigned mask, CollectiveFn out) {
    if (((w.arrived | w.exited) & mask) != mask) return;
    for (unsigned l = 0; l < 32; l++)
        if ((mask >> l) & 1u & ~(w.exited >> l)) w.out[l] = out(l, w.vals, w.aux, mask & ~w.exited);
    w.arrived &= ~mask;
    w.gen++;
    g_block->events++;
}
void fiber_main() {
    (*g_body)();
    Fiber& f = *g_cur;
    Block& b = *g_block;
    f.done = true;
    b.live--;
    b.events++;
    Warp& w = b.warps[f.info.warp];
    w.exited |= 1u << f.info.lane;                           // an exited lane never blocks the others (as on the hardware)
    if (w.arrived && w.pendingFn) complete_if_ready(w, w.pendingMask, w.pendingFn);
    release_barrier_if_complete(b);
    _longjmp(g_schedJb, 1);
}
}  // namespace

ThreadInfo& cur() { return g_cur ? g_cur->info : g_hostInfo; }

unsigned long long collective(unsigned mask, unsigned long long mine, unsigned aux, CollectiveFn out) {
    Fiber& f = *g_cur;
    const unsigned lane = f.info.lane, bit = 1u << lane;
    Warp& w = g_block->warps[f.info.warp];
    if ((mask & ~bit) == 0u) {                               // a collective of one
        unsigned long long v[32] = { 0 }; unsigned a[32] = { 0 };
        v[lane] = mine; a[lane] = aux;
        return out(lane, v, a, bit);
    }
    w.vals[lane] = mine; w.aux[lane] = aux; w.arrived |= bit;
    w.pendingMask = mask; w.pendingFn = out;
    g_block->events++;
    const unsigned gen = w.gen;
    // the last lane to arrive evaluates the collective for everybody; lanes that left the kernel count as arrived (an exit
    // re-evaluates the pending collective, see fiber_main), so a waiting lane is simply skipped by the scheduler until gen moves
    complete_if_ready(w, mask, out);
    while (w.gen == gen) {
        f.waitKind = 1; f.waitGen = gen;
        yield();
    }
    f.waitKind = 0;
    return w.out[lane];
}

void syncthreads() {
    Block& b = *g_block;
    const unsigned gen = b.syncGen;
    b.syncArrived++;
    b.events++;
    release_barrier_if_complete(b);
    while (b.syncGen == gen) {
        g_cur->waitKind = 2; g_cur->waitGen = gen;
        yield();
    }
    g_cur->waitKind = 0;
}

void launch(dim3 grid, dim3 block, const std::function<void()>& body) {
    const unsigned nThreads = block.x * block.y * block.z;
    if (nThreads == 0 || grid.x * grid.y * grid.z == 0) return;
    Block b;
    b.fibers.resize(nThreads);
    b.warps.resize((nThreads + 31) / 32);
    static std::vector<char*> stackPool;                       // fiber stacks are kept for the next launch
    while (stackPool.size() < nThreads) {
        char* st = (char*)mmap(nullptr, STACK_BYTES, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
        if (st == MAP_FAILED) { fprintf(stderr, "emu: cannot map a fiber stack\n"); abort(); }
        stackPool.push_back(st);
    }
    for (unsigned t = 0; t < nThreads; t++) b.fibers[t].stack = stackPool[t];
    g_body = &body;
    for (unsigned bz = 0; bz < grid.z; bz++) for (unsigned by = 0; by < grid.y; by++) for (unsigned bx = 0; bx < grid.x; bx++) {
        b.live = nThreads; b.syncArrived = 0; b.events = 0;
        for (auto& w : b.warps) { w.arrived = 0; w.exited = 0; w.pendingMask = 0; w.pendingFn = nullptr; }
        // lanes beyond the block size do not exist: they count as exited
        if (nThreads % 32) b.warps.back().exited = ~0u << (nThreads % 32);
        for (unsigned t = 0; t < nThreads; t++) {
            Fiber& f = b.fibers[t];
            f.done = false; f.started = false; f.waitKind = 0;
            f.info.tIdx = uint3{ t % block.x, (t / block.x) % block.y, t / (block.x * block.y) };
            f.info.bIdx = uint3{ bx, by, bz };
            f.info.bDim = block; f.info.gDim = grid;
            f.info.lane = t & 31; f.info.warp = t >> 5;
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = STACK_BYTES; f.ctx.uc_link = &g_sched;
            makecontext(&f.ctx, fiber_main, 0);
        }
        g_block = &b;
        while (b.live > 0) {
            const unsigned long long before = b.events;
            for (unsigned t = 0; t < nThreads; t++) {
                Fiber& f = b.fibers[t];
                if (f.done) continue;
                if (f.waitKind == 1 && b.warps[f.info.warp].gen == f.waitGen) continue;      // still waiting: no context switch
                if (f.waitKind == 2 && b.syncGen == f.waitGen) continue;
                g_cur = &f;
                if (!_setjmp(g_schedJb)) {
                    if (f.started) _longjmp(f.jb, 1);
                    f.started = true;
                    swapcontext(&g_sched, &f.ctx);               // first entry: onto the fiber's own stack
                }
            }
            g_cur = nullptr;
            if (b.live > 0 && b.events == before) {
                fprintf(stderr, "emu: dead-lock in block (%u,%u,%u): %u threads wait at collectives that cannot complete\n", bx, by, bz, b.live);
                abort();
            }
        }
        g_block = nullptr;
    }
    g_body = nullptr;
}
}  // namespace emu
