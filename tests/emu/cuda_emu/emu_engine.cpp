// emu_engine.cpp -- TEST INFRASTRUCTURE ONLY: fiber scheduler behind tests/emu/cuda_emu/cuda_runtime.h.
#include <stdio.h>
#include <ucontext.h>

#include <mutex>
#include <vector>

#include "cuda_runtime.h"

// Fiber switch: on x86-64 (outside AddressSanitizer builds) a dozen instructions that save the callee-saved registers and
// swap the stack pointer; swapcontext() does the same plus two sigprocmask system calls, which dominated the run time.
#if defined(__x86_64__) && !defined(__SANITIZE_ADDRESS__)
#define EMU_FAST_SWITCH 1
extern "C" void emu_switch(void **save_sp, void *new_sp);
asm(".text\n"
    ".globl emu_switch\n"
    ".type emu_switch,@function\n"
    "emu_switch:\n"
    "  pushq %rbp\n  pushq %rbx\n  pushq %r12\n  pushq %r13\n  pushq %r14\n  pushq %r15\n"
    "  movq %rsp, (%rdi)\n"
    "  movq %rsi, %rsp\n"
    "  popq %r15\n  popq %r14\n  popq %r13\n  popq %r12\n  popq %rbx\n  popq %rbp\n"
    "  ret\n"
    ".size emu_switch,.-emu_switch\n");
#else
#define EMU_FAST_SWITCH 0
#endif

namespace emu {

enum { READY = 0, WAIT_WARP = 1, WAIT_BLOCK = 2, DONE = 3 };

struct Barrier {          // rendezvous of the lanes named by one mask
    uint32_t mask = 0;
    uint64_t slot[2][32];
    uint32_t live_snap[2];
    int arrived = 0;
    unsigned gen = 0;
};

struct Warp {
    Barrier bars[40];     // one per distinct mask seen in this block (full warp, 2 x 16, 4 x 8, 8 x 4 lanes, ...)
    int n_bars = 0;
    uint32_t live_mask = 0;
    int live = 0;
};

struct Thread {
#if EMU_FAST_SWITCH
    void *sp = nullptr;
#else
    ucontext_t ctx;
#endif
    int tid = 0, state = READY;
    unsigned wait_gen = 0;
    Warp *warp = nullptr;
    Barrier *bar = nullptr;
};

struct Engine {
#if EMU_FAST_SWITCH
    void *sched_sp = nullptr;
#else
    ucontext_t sched;
#endif
    std::vector<Thread> threads;
    std::vector<Warp> warps;
    std::vector<char *> stacks;
    const std::function<void()> *body = nullptr;
    unsigned grid = 0, block = 0, bidx = 0;
    int block_live = 0, block_arrived = 0;
    unsigned block_gen = 0;
};

static constexpr size_t kStack = 256 * 1024;
thread_local Thread *cur = nullptr;
static thread_local Engine *E = nullptr;

uint3 thread_idx() { return {(unsigned)cur->tid, 0, 0}; }
uint3 block_idx() { return {E->bidx, 0, 0}; }
dim3 block_dim() { return dim3(E->block); }
dim3 grid_dim() { return dim3(E->grid); }
int lane_id() { return cur->tid & 31; }

#if EMU_FAST_SWITCH
static void yield_to_scheduler() { emu_switch(&cur->sp, E->sched_sp); }
#else
static void yield_to_scheduler() { swapcontext(&cur->ctx, &E->sched); }
#endif

static void release(Warp *w, Barrier *b) {
    b->live_snap[b->gen & 1] = w->live_mask & b->mask;
    b->arrived = 0;
    b->gen++;
}

static Barrier *barrier_for(Warp *w, uint32_t mask) {
    for (int k = 0; k < w->n_bars; k++)
        if (w->bars[k].mask == mask) return &w->bars[k];
    if (w->n_bars == 40) { fprintf(stderr, "cuda_emu: too many distinct warp masks\n"); abort(); }
    Barrier *b = &w->bars[w->n_bars++];
    *b = Barrier();
    b->mask = mask;
    return b;
}

uint64_t warp_exchange(unsigned mask, uint64_t mine, uint64_t out[32], uint32_t *live_mask) {
    Thread *t = cur;
    Warp *w = t->warp;
    const int lane = t->tid & 31;
    if (!((mask >> lane) & 1u)) { fprintf(stderr, "cuda_emu: lane %d is not in its own mask %08x\n", lane, mask); abort(); }
    Barrier *b = barrier_for(w, mask);
    const unsigned g = b->gen, buf = g & 1;
    b->slot[buf][lane] = mine;
    b->arrived++;
    if (b->arrived >= __builtin_popcount(mask & w->live_mask)) {
        release(w, b);
    } else {
        t->state = WAIT_WARP;
        t->bar = b;
        t->wait_gen = g;
        while (b->gen == g) yield_to_scheduler();
        t->state = READY;
    }
    memcpy(out, b->slot[buf], sizeof(b->slot[buf]));
    *live_mask = b->live_snap[buf];
    return mine;
}

void sync_block() {
    Thread *t = cur;
    const unsigned g = E->block_gen;
    E->block_arrived++;
    if (E->block_arrived == E->block_live) {
        E->block_arrived = 0;
        E->block_gen++;
        return;
    }
    t->state = WAIT_BLOCK;
    t->wait_gen = g;
    while (E->block_gen == g) yield_to_scheduler();
    t->state = READY;
}

static void trampoline() {
    Thread *t = cur;
    (*E->body)();
    // thread exit: it no longer takes part in barriers
    t->state = DONE;
    Warp *w = t->warp;
    w->live--;
    w->live_mask &= ~(1u << (t->tid & 31));
    for (int k = 0; k < w->n_bars; k++) {   // lanes that wait for this one at some barrier no longer have to
        Barrier *b = &w->bars[k];
        if (b->arrived > 0 && b->arrived >= __builtin_popcount(b->mask & w->live_mask)) release(w, b);
    }
    E->block_live--;
    if (E->block_live > 0 && E->block_arrived == E->block_live) { E->block_arrived = 0; E->block_gen++; }
    yield_to_scheduler();   // a finished thread is never resumed
    abort();
}

static bool resumable(const Thread &t) {
    if (t.state == READY) return true;
    if (t.state == WAIT_WARP) return t.bar->gen != t.wait_gen;
    if (t.state == WAIT_BLOCK) return E->block_gen != t.wait_gen;
    return false;
}

static void run_block(unsigned b) {
    Engine &e = *E;
    e.bidx = b;
    const unsigned T = e.block, nw = (T + 31) / 32;
    e.threads.resize(T);
    e.warps.clear();
    e.warps.resize(nw);
    while (e.stacks.size() < T) e.stacks.push_back((char *)malloc(kStack));
    e.block_live = (int)T;
    e.block_arrived = 0;
    for (unsigned i = 0; i < T; i++) {
        Thread &t = e.threads[i];
        t.tid = (int)i;
        t.state = READY;
        t.warp = &e.warps[i / 32];
        t.warp->live++;
        t.warp->live_mask |= 1u << (i & 31);
#if EMU_FAST_SWITCH
        {   // initial frame: six zeroed callee-saved registers, then the entry point as the return address
            uintptr_t top = ((uintptr_t)e.stacks[i] + kStack) & ~(uintptr_t)15;
            void **f = (void **)(top - 64);
            for (int q = 0; q < 6; q++) f[q] = nullptr;
            f[6] = (void *)trampoline;
            f[7] = nullptr;
            t.sp = f;
        }
#else
        getcontext(&t.ctx);
        t.ctx.uc_stack.ss_sp = e.stacks[i];
        t.ctx.uc_stack.ss_size = kStack;
        t.ctx.uc_link = nullptr;
        makecontext(&t.ctx, trampoline, 0);
#endif
    }
    while (e.block_live > 0) {
        bool progress = false;
        for (unsigned wi = 0; wi < nw; wi++) {
            bool again = true;
            while (again) {  // run the lanes of this warp until all of them are finished or blocked
                again = false;
                for (unsigned i = wi * 32; i < std::min(T, wi * 32 + 32); i++) {
                    Thread &t = e.threads[i];
                    if (!resumable(t)) continue;
                    cur = &t;
#if EMU_FAST_SWITCH
                    emu_switch(&e.sched_sp, t.sp);
#else
                    swapcontext(&e.sched, &t.ctx);
#endif
                    cur = nullptr;
                    again = progress = true;
                }
            }
        }
        if (!progress) {
            fprintf(stderr, "cuda_emu: deadlock in block %u (threads wait for a barrier that cannot complete)\n", b);
            abort();
        }
    }
}

// `__shared__` variables are function-level statics here, so only one kernel may run at a time in the process: host
// threads that launch concurrently (the library's two-lane mode) take turns
static std::mutex g_launch_mutex;

void Launch::operator<<(const std::function<void()> &body) const {
    if (grid == 0 || block == 0) return;
    std::lock_guard<std::mutex> guard(g_launch_mutex);
    Engine engine;
    Engine *outer = E;
    E = &engine;
    engine.body = &body;
    engine.grid = grid;
    engine.block = block;
    for (unsigned b = 0; b < grid; b++) run_block(b);
    for (char *s : engine.stacks) free(s);
    E = outer;
}

}  // namespace emu
