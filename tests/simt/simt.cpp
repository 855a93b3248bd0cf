/* simt.cpp -- fibre scheduler for the CUDA emulator in simt.h (test infrastructure). */
#include "simt.h"

#include <sys/mman.h>

namespace simt {

Block *g_block = nullptr;
uint3  g_threadIdx, g_blockIdx;
dim3   g_blockDim, g_gridDim;

static const size_t kStackBytes = 256 * 1024;

#ifdef LZS_SIMT_FAST_SWITCH
/* Save the callee-saved registers of the System V x86-64 ABI on the current stack, store the
 * stack pointer, load the other fibre's and return into it. */
asm(R"(
    .text
    .globl simt_switch
    .type simt_switch,@function
simt_switch:
    pushq %rbp
    pushq %rbx
    pushq %r12
    pushq %r13
    pushq %r14
    pushq %r15
    movq %rsp, (%rdi)
    movq %rsi, %rsp
    popq %r15
    popq %r14
    popq %r13
    popq %r12
    popq %rbx
    popq %rbp
    ret
    .size simt_switch,.-simt_switch
)");
#endif

static void fiber_entry()
{
    Block *b = g_block;
    b->body();
    b->fibers[b->current].done = true;
#ifdef LZS_SIMT_FAST_SWITCH
    simt_switch(&b->fibers[b->current].sp, b->sched_sp);
#else
    swapcontext(&b->fibers[b->current].ctx, &b->sched);
#endif
    abort();                            /* a finished fibre is never resumed */
}

uint8_t *dyn_smem() { return g_block->dyn_smem.data(); }

static void run_block(Block &b, unsigned nthreads)
{
    b.nthreads = nthreads;
    b.fibers.resize(nthreads);
    b.rdv.assign((nthreads + 31) / 32, {});
    b.bar_arrived = 0;
    b.named.clear();
    for (unsigned t = 0; t < nthreads; t++) {
        Fiber &f = b.fibers[t];
        f.done = false;
        if (!f.stack) {
            f.stack = (char *)mmap(nullptr, kStackBytes, PROT_READ | PROT_WRITE,
                                   MAP_PRIVATE | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
            if (f.stack == MAP_FAILED) { perror("mmap"); abort(); }
        }
        f.tid.x = t % g_blockDim.x;
        f.tid.y = (t / g_blockDim.x) % g_blockDim.y;
        f.tid.z = t / (g_blockDim.x * g_blockDim.y);
#ifdef LZS_SIMT_FAST_SWITCH
        /* first switch: six zeroed registers are popped, then `ret` enters fiber_entry with the
         * stack as after a call (return-address slot on top, 16-byte aligned above it) */
        void **top = reinterpret_cast<void **>(f.stack + kStackBytes);
        *--top = nullptr;                                   /* fiber_entry's (unused) return address */
        *--top = reinterpret_cast<void *>(&fiber_entry);
        for (int r = 0; r < 6; r++) *--top = nullptr;
        f.sp = top;
#else
        getcontext(&f.ctx);
        f.ctx.uc_stack.ss_sp = f.stack;
        f.ctx.uc_stack.ss_size = kStackBytes;
        f.ctx.uc_link = nullptr;
        makecontext(&f.ctx, fiber_entry, 0);
#endif
    }
    unsigned remaining = nthreads;
    uint64_t idle_passes = 0;
    while (remaining) {
        unsigned before = remaining;
        for (unsigned t = 0; t < nthreads; t++) {
            Fiber &f = b.fibers[t];
            if (f.done) continue;
            b.current = (int)t;
            g_threadIdx = f.tid;
#ifdef LZS_SIMT_FAST_SWITCH
            simt_switch(&b.sched_sp, f.sp);
#else
            swapcontext(&b.sched, &f.ctx);
#endif
            if (f.done) remaining--;
        }
        if (remaining == before) {
            if (++idle_passes > 400000000ull) {
                fprintf(stderr, "simt: no thread finished for a very long time; deadlock?\n");
                abort();
            }
        } else {
            idle_passes = 0;
        }
    }
}

void launch(dim3 grid, dim3 block, size_t dyn_smem_bytes, const std::function<void()> &body)
{
    static Block b;                 /* stacks are reused across launches */
    g_block = &b;
    g_blockDim = block;
    g_gridDim = grid;
    b.body = body;
    b.dyn_smem.assign(dyn_smem_bytes + 64, 0xCD);
    unsigned nthreads = block.x * block.y * block.z;
    for (unsigned z = 0; z < grid.z; z++)
        for (unsigned y = 0; y < grid.y; y++)
            for (unsigned x = 0; x < grid.x; x++) {
                g_blockIdx.x = x; g_blockIdx.y = y; g_blockIdx.z = z;
                run_block(b, nthreads);
            }
    g_block = nullptr;
}

}  // namespace simt

namespace simt {
int g_scramble_exchanges = 0;
}  // namespace simt
