// cuda_emu.cpp -- TEST INFRASTRUCTURE.  The execution model behind ddc_host_emu.h: runs a CUDA kernel that was
// compiled as host C++ (-DDDC_HOST_EMU) the way a GPU would, only slowly.
//
//   * every thread of a block is a fiber (ucontext) with its own stack; the block's fibers are resumed
//     round-robin by a scheduler and give up the processor only inside __syncthreads(), __syncthreads_or()
//     and the warp collectives -- there is no preemption, so plain memory accesses model atomics;
//   * __syncthreads(): a fiber waits until every fiber of the block that has not returned has arrived;
//   * warp collectives (ddc_emu_warp_gather, on which __shfl_*_sync, __ballot_sync, __reduce_or_sync build):
//     the lanes named in the mask deposit their values; when all of them (that have not returned) are in, the
//     32 values are published and every lane reads what its intrinsic needs;
//   * `__shared__` variables are function-local statics (one block runs at a time), dynamic shared memory is
//     one aligned buffer;
//   * blocks of a grid run one after the other, in x-fastest order -- or, with set_schedule_seed(), blocks in a
//     random order and the threads of a block in a new random order every scheduling round: a kernel whose
//     result depends on the interleaving (a missing barrier, an assumption about block order) then gives
//     different results for different seeds.
// Divergence needs no modelling: every fiber simply executes its own path.  A kernel that would deadlock on a
// GPU (a barrier some live threads never reach) is reported instead of hanging.
#include "cuda_emu.h"

#include <setjmp.h>
#include <ucontext.h>

#include <cstdio>
#include <cstdlib>
#include <mutex>
#include <stdexcept>
#include <string>
#include <vector>

ddc_emu_dim threadIdx = { 0, 0, 0 }, blockIdx = { 0, 0, 0 }, blockDim = { 1, 1, 1 }, gridDim = { 1, 1, 1 };
int ddc_emu_fdividef_ulps = 0;

namespace {
constexpr size_t STACK_BYTES = 256 * 1024;
constexpr size_t DYN_SMEM_BYTES = 256 * 1024;

struct Warp {
    unsigned alive = 0; // lanes that have not returned
    unsigned arrived = 0; // lanes that have deposited for the collective being gathered
    unsigned long long vals[32];
    // published results of the last completed collective
    unsigned long long out[32];
    unsigned out_present = 0;
    unsigned to_read = 0; // lanes that still have to pick the results up
    unsigned long done_gen = 0, gen = 1;
};

// A fiber is entered the first time through its ucontext (that is what puts it on its own stack); every later
// switch is _setjmp / _longjmp, which -- unlike swapcontext -- makes no system call to save the signal mask.
struct Fiber {
    ucontext_t ctx;
    jmp_buf jb;
    ddc_emu_dim tid;
    bool started = false, done = false;
};

struct BlockState {
    std::vector<Fiber> fibers;
    std::vector<Warp> warps;
    ucontext_t sched;
    jmp_buf sched_jb;
    int current = -1;
    int alive = 0;
    // block barrier
    int bar_arrived = 0;
    unsigned long bar_gen = 0;
    int or_acc[2] = { 0, 0 }, or_res[2] = { 0, 0 };
    unsigned long progress = 0; // bumped whenever any fiber gets past a wait: deadlock detection
    const std::function<void()>* body = nullptr;
};

BlockState* g_blk = nullptr;
unsigned long long g_sched_seed = 0; // 0: threads and blocks in index order; else: a different random order every round
unsigned long long sched_rand()
{
    g_sched_seed ^= g_sched_seed << 13;
    g_sched_seed ^= g_sched_seed >> 7;
    g_sched_seed ^= g_sched_seed << 17;
    return g_sched_seed;
}
void shuffle(std::vector<int>& v)
{
    for (size_t i = v.size(); i > 1; i--)
        std::swap(v[i - 1], v[sched_rand() % i]);
}
std::vector<char> g_stacks; // fibers * STACK_BYTES, reused by every launch
alignas(64) char g_dyn_smem[DYN_SMEM_BYTES];
std::string g_error;

void yield()
{
    BlockState& b = *g_blk;
    Fiber& f = b.fibers[b.current];
    if (!_setjmp(f.jb))
        _longjmp(b.sched_jb, 1);
}
// scheduler side: run fiber t until it yields or returns
void resume(BlockState& b, int t)
{
    Fiber& f = b.fibers[t];
    b.current = t;
    threadIdx = f.tid;
    if (!_setjmp(b.sched_jb)) {
        if (!f.started) {
            f.started = true;
            setcontext(&f.ctx);
        } else
            _longjmp(f.jb, 1);
    }
}

int linear_tid() { return (int)(threadIdx.x + threadIdx.y * blockDim.x + threadIdx.z * blockDim.x * blockDim.y); }

void release_barrier_if_complete(BlockState& b)
{
    if (b.alive > 0 && b.bar_arrived == b.alive) {
        const int g = (int)(b.bar_gen & 1);
        b.or_res[g] = b.or_acc[g];
        b.or_acc[g] = 0;
        b.bar_arrived = 0;
        b.bar_gen++;
        b.progress++;
    }
}

void fiber_main()
{
    BlockState& b = *g_blk;
    Fiber& f = b.fibers[b.current];
    (*b.body)();
    // the thread returns: it no longer takes part in barriers or collectives
    f.done = true;
    b.alive--;
    const int t = b.current;
    Warp& w = b.warps[t >> 5];
    w.alive &= ~(1u << (t & 31));
    release_barrier_if_complete(b);
    b.progress++;
    _longjmp(b.sched_jb, 1);
}
} // namespace

void __syncthreads() { (void)__syncthreads_or(0); }

int __syncthreads_or(int predicate)
{
    BlockState& b = *g_blk;
    const unsigned long gen = b.bar_gen;
    const int g = (int)(gen & 1);
    if (predicate)
        b.or_acc[g] = 1;
    b.bar_arrived++;
    b.progress++;
    release_barrier_if_complete(b);
    while (b.bar_gen == gen)
        yield();
    return b.or_res[g];
}

const unsigned long long* ddc_emu_warp_gather(unsigned mask, unsigned long long v, unsigned* present)
{
    BlockState& b = *g_blk;
    const int t = linear_tid();
    Warp& w = b.warps[t >> 5];
    const unsigned bit = 1u << (t & 31);
    // the results of the previous collective must have been picked up by everyone before values are overwritten
    while (w.to_read != 0)
        yield();
    const unsigned long my_gen = w.gen;
    w.vals[t & 31] = v;
    w.arrived |= bit;
    b.progress++;
    for (;;) {
        const unsigned need = mask & w.alive;
        if (w.done_gen == my_gen)
            break; // somebody completed it
        if ((w.arrived & need) == need) { // I am the last one in: publish
            for (int l = 0; l < 32; l++)
                w.out[l] = (w.arrived >> l & 1u) ? w.vals[l] : 0ull;
            w.out_present = w.arrived;
            w.to_read = w.arrived;
            w.arrived = 0;
            w.done_gen = my_gen;
            w.gen++;
            b.progress++;
            break;
        }
        yield();
    }
    *present = w.out_present;
    // copy out: the lane may run on while later arrivals overwrite nothing (to_read guards the buffer)
    // (one buffer for all fibers: the caller consumes it before it can yield)
    static unsigned long long mine[32];
    for (int l = 0; l < 32; l++)
        mine[l] = w.out[l];
    w.to_read &= ~bit;
    b.progress++;
    return mine;
}

void* ddc_emu_dyn_smem() { return g_dyn_smem; }

namespace cuda_emu {

const char* last_error() { return g_error.c_str(); }
void set_schedule_seed(unsigned long long seed) { g_sched_seed = seed; }

bool launch(Dim3 grid, Dim3 block, size_t dyn_smem, const std::function<void()>& body)
{
    // one launch at a time: the built-in variables, the `__shared__` statics and the fiber stacks are process-wide
    // (callers may be several OS threads, e.g. the thread-ranks of oracle/ref_hostpath_shim.cpp)
    static std::mutex launch_mutex;
    std::lock_guard<std::mutex> lock(launch_mutex);
    const int nthreads = (int)(block.x * block.y * block.z);
    if (nthreads < 1 || nthreads > 1024 || dyn_smem > DYN_SMEM_BYTES) {
        g_error = "cuda_emu::launch: bad block size or too much dynamic shared memory";
        return false;
    }
    if (g_stacks.size() < (size_t)nthreads * STACK_BYTES)
        g_stacks.resize((size_t)nthreads * STACK_BYTES);
    gridDim = { grid.x, grid.y, grid.z };
    blockDim = { block.x, block.y, block.z };
    BlockState b;
    b.body = &body;
    g_blk = &b;
    std::vector<int> block_order((size_t)grid.x * grid.y * grid.z), thread_order((size_t)nthreads);
    for (size_t i = 0; i < block_order.size(); i++)
        block_order[i] = (int)i;
    for (int t = 0; t < nthreads; t++)
        thread_order[t] = t;
    if (g_sched_seed)
        shuffle(block_order);
    for (int blk : block_order) {
                const unsigned bx = (unsigned)blk % grid.x, by = ((unsigned)blk / grid.x) % grid.y,
                               bz = (unsigned)blk / (grid.x * grid.y);
                blockIdx = { bx, by, bz };
                b.fibers.assign((size_t)nthreads, Fiber());
                b.warps.assign((size_t)(nthreads + 31) / 32, Warp());
                b.alive = nthreads;
                b.bar_arrived = 0;
                b.bar_gen = 0;
                b.or_acc[0] = b.or_acc[1] = b.or_res[0] = b.or_res[1] = 0;
                for (int t = 0; t < nthreads; t++) {
                    Fiber& f = b.fibers[t];
                    f.tid = { (unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y) };
                    b.warps[t >> 5].alive |= 1u << (t & 31);
                    getcontext(&f.ctx);
                    f.ctx.uc_stack.ss_sp = g_stacks.data() + (size_t)t * STACK_BYTES;
                    f.ctx.uc_stack.ss_size = STACK_BYTES;
                    f.ctx.uc_link = &b.sched;
                    makecontext(&f.ctx, fiber_main, 0);
                }
                // round-robin until every fiber has returned; a full round without any progress is a deadlock
                int remaining = nthreads;
                while (remaining > 0) {
                    const unsigned long before = b.progress;
                    bool ran = false;
                    if (g_sched_seed)
                        shuffle(thread_order);
                    for (int t : thread_order) {
                        Fiber& f = b.fibers[t];
                        if (f.done)
                            continue;
                        resume(b, t);
                        ran = true;
                        if (f.done)
                            remaining--;
                    }
                    if (ran && remaining > 0 && b.progress == before) {
                        // second chance: a round in which fibers only re-checked their conditions
                        const unsigned long again = b.progress;
                        for (int t : thread_order) {
                            Fiber& f = b.fibers[t];
                            if (f.done)
                                continue;
                            resume(b, t);
                            if (f.done)
                                remaining--;
                        }
                        if (remaining > 0 && b.progress == again) {
                            char msg[256];
                            std::snprintf(msg, sizeof msg,
                                "cuda_emu: deadlock in block (%u,%u,%u): %d thread(s) wait at a barrier or a warp "
                                "collective that the others never reach",
                                bx, by, bz, remaining);
                            g_error = msg;
                            g_blk = nullptr;
                            return false;
                        }
                    }
                }
            }
    g_blk = nullptr;
    return true;
}

} // namespace cuda_emu
