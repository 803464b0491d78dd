// tests/emu/emu.cpp -- TEST INFRASTRUCTURE ONLY: the fiber scheduler behind tests/emu/include/cuda_runtime.h, plus the host-side pieces of
// the library the emulated translation units expect (error string, workspace arena, stubs for the models that are not emulated).
#include "kokoro.h"
#include "dac.h"

#include <cstdarg>
#include <algorithm>
#include <cstdio>
#include <vector>

extern "C" void b2emu_switch(void ** save_sp, void * load_sp);
// x86-64 SysV: callee-saved registers on the old stack, swap stack pointers, restore from the new stack
asm(R"(
.text
.globl b2emu_switch
.type b2emu_switch,@function
b2emu_switch:
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
.size b2emu_switch,.-b2emu_switch
)");

namespace b2emu {

Fiber * g_cur = nullptr;
dim3 g_blockDim, g_gridDim;
uint64_t g_launches = 0, g_blocks = 0;

namespace {
constexpr size_t STACK = 96 * 1024;
void * g_sched_sp = nullptr;
const std::function<void()> * g_body = nullptr;
std::vector<Fiber> g_fibers;
std::vector<char *> g_stacks;
// per-block barrier state.  Ordinary launches run their blocks one after another (one live Block); a cooperative launch (launch_coop) keeps every block of the grid
// alive at once -- all their threads are fibers of one scheduler -- so that a grid-wide barrier in global memory can complete.
struct NamedBar { int arrived = 0; uint64_t gen = 0; };
struct Block { int alive = 0, arrived = 0; uint64_t gen = 0; std::vector<char> dyn; NamedBar nb[16]; };
std::vector<Block> g_blk;
int g_nt = 0;                                                     // threads per block of the running launch
uint64_t g_progress = 0;
struct Warp { int alive = 0, arrived = 0; uint64_t gen = 0; uint64_t slot[2][32]; unsigned wide[2][32][8]; };
std::vector<Warp> g_warps;

void yield() { Fiber * f = g_cur; b2emu_switch(&f->sp, g_sched_sp); }
int linear_tid(const Fiber * f) { return (int) (f - g_fibers.data()); }

// A thread that has to wait hands the CPU straight to the next live thread of its group (warp or block) instead of going through the scheduler: in lockstep code
// every lane then costs one context switch per barrier.  After a few fruitless rounds (divergent code) it falls back to the round-robin scheduler.
// B2EMU_REVERSE=1 runs the threads of a block (and hands over at barriers) in descending order instead of ascending: a kernel whose result depends on the
// order in which threads reach a barrier-free region -- a missing __syncthreads, a race on shared memory -- gives a different answer under the two schedules.
const bool g_reverse = [] { const char * e = getenv("B2EMU_REVERSE"); return e && e[0] == '1'; }();

void wait_pass_on(int lo, int hi, int & tries) {
    Fiber * f = g_cur;
    if (tries++ < 3 * (hi - lo)) {
        int t = linear_tid(f);
        for (int k = 1; k < hi - lo; k++) {
            const int c = lo + (t - lo + (g_reverse ? (hi - lo) - k : k)) % (hi - lo);
            Fiber * n = &g_fibers[(size_t) c];
            if (!n->done) { g_cur = n; b2emu_switch(&f->sp, n->sp); return; }
        }
    }
    yield();
}

void release_if_complete(Block & b) {      // a thread that exits while others wait at a barrier completes that barrier (CUDA leaves this undefined; be lenient)
    if (b.alive > 0 && b.arrived == b.alive) { b.arrived = 0; b.gen++; }
}

void fiber_entry() {
    (*g_body)();
    Fiber * f = g_cur;
    f->done = true; g_progress++;
    Block & b = g_blk[(size_t) f->blk];
    b.alive--;
    Warp & w = g_warps[(size_t) linear_tid(f) >> 5];
    w.alive--;
    if (w.alive > 0 && w.arrived == w.alive) { w.arrived = 0; w.gen++; }
    release_if_complete(b);
    yield();
    abort();   // a finished fiber is never resumed
}

void warp_barrier() {
    Warp & w = g_warps[(size_t) linear_tid(g_cur) >> 5];
    const uint64_t gen = w.gen; g_progress++;
    if (++w.arrived == w.alive) { w.arrived = 0; w.gen++; return; }
    const int lo = linear_tid(g_cur) & ~31, hi = std::min(lo + 32, (int) g_fibers.size());
    int tries = 0;
    while (w.gen == gen) wait_pass_on(lo, hi, tries);
}
}  // namespace

void sync_block() {
    Block & b = g_blk[(size_t) g_cur->blk];
    const uint64_t gen = b.gen; g_progress++;
    if (++b.arrived == b.alive) { b.arrived = 0; b.gen++; return; }
    const int lo = g_cur->blk * g_nt;
    int tries = 0;
    while (b.gen == gen) wait_pass_on(lo, lo + g_nt, tries);
}

// bar.sync id, nthreads: the first `nthreads` arrivals at barrier `id` of this block release each other (which threads take part is the kernel's business)
void named_bar(int id, int nthreads) {
    if (id < 0 || id >= 16) { fprintf(stderr, "b2emu: named barrier %d\n", id); abort(); }
    Block & b = g_blk[(size_t) g_cur->blk];
    NamedBar & nb = b.nb[id];
    const uint64_t gen = nb.gen; g_progress++;
    if (++nb.arrived == nthreads) { nb.arrived = 0; nb.gen++; return; }
    const int lo = g_cur->blk * g_nt;
    int tries = 0;
    while (nb.gen == gen) wait_pass_on(lo, lo + g_nt, tries);
}

// a thread that polls memory another thread will write (mbarrier phase, grid-barrier counter) gives the CPU away; writers call note_progress so that the
// scheduler's deadlock detector can tell "everybody polls and nothing changes" from work in flight
void yield_spin() { yield(); }
void note_progress() { g_progress++; }

uint64_t shfl(uint64_t v, int src_lane) {
    const int t = linear_tid(g_cur);
    Warp & w = g_warps[(size_t) t >> 5];
    // double-buffered by barrier generation: a lane can only overwrite a buffer two shuffles later, after every lane has passed the barrier in between
    // (and read its value before arriving there), so one barrier per shuffle is enough
    const int buf = (int) (w.gen & 1);
    w.slot[buf][t & 31] = v;
    warp_barrier();
    return w.slot[buf][src_lane & 31];
}

void * dyn_smem() { return g_blk[(size_t) g_cur->blk].dyn.data(); }

void warp_exchange(const unsigned * mine, int nwords, unsigned * all) {
    if (nwords > 8) { fprintf(stderr, "b2emu: warp_exchange of %d words\n", nwords); abort(); }
    const int t = linear_tid(g_cur);
    Warp & w = g_warps[(size_t) t >> 5];
    const int buf = (int) (w.gen & 1);
    for (int i = 0; i < nwords; i++) w.wide[buf][t & 31][i] = mine[i];
    warp_barrier();
    for (int l = 0; l < 32; l++) for (int i = 0; i < nwords; i++) all[l * nwords + i] = w.wide[buf][l][i];
}

struct Recorded { dim3 grid, block; size_t smem; std::function<void()> body; };
}  // namespace b2emu
struct b2emu_graph { std::vector<b2emu::Recorded> nodes; int refs = 1; };
namespace b2emu {
static b2emu_graph * g_capture = nullptr;
uint64_t g_replays = 0;

static void launch_impl(dim3 grid, dim3 block, size_t smem, const std::function<void()> & body, bool coop) {
    if (g_cur) { fprintf(stderr, "b2emu: nested launch\n"); abort(); }
    if (g_capture) { if (coop) { fprintf(stderr, "b2emu: cooperative launch during capture\n"); abort(); } g_capture->nodes.push_back(Recorded{grid, block, smem, body}); return; }   // captured, not executed -- like a real stream capture
    const int nt = (int) (block.x * block.y * block.z);
    if (nt <= 0 || nt > 1024) { fprintf(stderr, "b2emu: %d threads per block\n", nt); abort(); }
    if (smem > 227 * 1024) { fprintf(stderr, "b2emu: %zu bytes of dynamic shared memory\n", smem); abort(); }
    const int nblocks = (int) (grid.x * grid.y * grid.z);
    const int live_blocks = coop ? nblocks : 1;                     // blocks alive at once
    if (coop && nt % 32) { fprintf(stderr, "b2emu: cooperative launch with %d threads per block\n", nt); abort(); }
    g_launches++;
    g_body = &body; g_blockDim = block; g_gridDim = grid; g_nt = nt;
    const int nf = nt * live_blocks;
    while ((int) g_stacks.size() < nf) g_stacks.push_back((char *) aligned_alloc(64, STACK));
    g_fibers.assign((size_t) nf, Fiber{});
    g_blk.assign((size_t) live_blocks, Block{});
    for (int first = 0; first < nblocks; first += live_blocks) {
        g_warps.assign((size_t) (nf + 31) / 32, Warp{});
        for (int lb = 0; lb < live_blocks; lb++) {
            const int bi = first + lb;
            const dim3 bidx((unsigned) bi % grid.x, (unsigned) bi / grid.x % grid.y, (unsigned) bi / (grid.x * grid.y));
            g_blocks++;
            Block & b = g_blk[(size_t) lb];
            b.alive = nt; b.arrived = 0; b.gen = 0; b.dyn.assign(smem + 64, 0);
            for (auto & nb : b.nb) nb = NamedBar{};
            for (int t = 0; t < nt; t++) {
                Fiber & f = g_fibers[(size_t) lb * nt + t];
                f.tid = uint3{(unsigned) t % block.x, (unsigned) t / block.x % block.y, (unsigned) t / (block.x * block.y)};
                f.bidx = bidx; f.blk = lb;
                f.done = false; f.stack = g_stacks[(size_t) lb * nt + t];
                g_warps[((size_t) lb * nt + t) >> 5].alive++;
                void ** sp = (void **) (f.stack + STACK - 64);   // 6 callee-saved slots, then the entry address `ret` pops: rsp % 16 == 8 at entry
                for (int i = 0; i < 6; i++) sp[i] = nullptr;
                sp[6] = (void *) &fiber_entry; sp[7] = nullptr;
                f.sp = sp;
            }
        }
        int live = nf, stalls = 0;
        while (live > 0) {
            const uint64_t before = g_progress;
            for (int tt = 0; tt < nf; tt++) {
                const int t = g_reverse ? nf - 1 - tt : tt;
                Fiber & f = g_fibers[(size_t) t];
                if (f.done) continue;
                g_cur = &f;
                b2emu_switch(&g_sched_sp, f.sp);
                g_cur = nullptr;
            }
            live = 0;                                           // threads hand the CPU to each other directly, so any of them may have finished in this round
            for (int t = 0; t < nf; t++) if (!g_fibers[(size_t) t].done) live++;
            stalls = (live > 0 && g_progress == before) ? stalls + 1 : 0;
            if (stalls > (coop ? 4 : 0)) {
                const Fiber & f0 = *std::find_if(g_fibers.begin(), g_fibers.end(), [](const Fiber & f) { return !f.done; });
                fprintf(stderr, "b2emu: deadlock -- %d threads wait at a barrier not all reach (first: block (%u,%u,%u) thread %u)\n", live, f0.bidx.x, f0.bidx.y, f0.bidx.z, f0.tid.x);
                abort();
            }
        }
    }
    g_body = nullptr;
}
void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()> & body) { launch_impl(grid, block, smem, body, false); }
void launch_coop(dim3 grid, dim3 block, size_t smem, const std::function<void()> & body) { launch_impl(grid, block, smem, body, true); }

}  // namespace b2emu

cudaError_t cudaStreamBeginCapture(cudaStream_t, cudaStreamCaptureMode) { if (b2emu::g_capture) return 1; b2emu::g_capture = new b2emu_graph(); return 0; }
cudaError_t cudaStreamEndCapture(cudaStream_t, cudaGraph_t * g) { if (!b2emu::g_capture) return 1; *g = b2emu::g_capture; b2emu::g_capture = nullptr; return 0; }
cudaError_t cudaGraphInstantiate(cudaGraphExec_t * e, cudaGraph_t g, unsigned long long) { g->refs++; *e = g; return 0; }
cudaError_t cudaGraphLaunch(cudaGraphExec_t e, cudaStream_t) {
    b2emu::g_replays++;
    for (const auto & n : e->nodes) b2emu::launch(n.grid, n.block, n.smem, n.body);
    return 0;
}
cudaError_t cudaGraphExecDestroy(cudaGraphExec_t e) { if (--e->refs == 0) delete e; return 0; }
cudaError_t cudaGraphDestroy(cudaGraph_t g) { if (--g->refs == 0) delete g; return 0; }

namespace b2 {

static char g_err[1024] = "";
void set_error(const char * fmt, ...) { va_list ap; va_start(ap, fmt); vsnprintf(g_err, sizeof g_err, fmt, ap); va_end(ap); }
const char * emu_last_error() { return g_err; }

// The emulated workspace arena hands out separately malloc'd blocks (NaN-filled) instead of slices of one allocation: under AddressSanitizer
// (tests/test_emu_cpu.py::test_address_sanitizer_*) every workspace buffer then has its own red zones, so a kernel that indexes past the end of one is caught
// even when the neighbouring slice would have absorbed the access on a real device.
static std::vector<void *> g_arena_blocks;
static void arena_drop() { for (void * p : g_arena_blocks) free(p); g_arena_blocks.clear(); }
int Arena::reserve(size_t bytes) {
    arena_drop();
    base = (char *) 1; cap = bytes; off = 0;                      // only the budget is tracked; `base` is never dereferenced
    return 0;
}
void * Arena::alloc(size_t bytes) {
    if (off + bytes > cap) { set_error("workspace arena exhausted (%zu + %zu > %zu)", off, bytes, cap); return nullptr; }
    off += (bytes + 255) & ~(size_t) 255;
    void * p = nullptr;                                           // exact size: the sanitizer's red zone starts right behind the last element
    if (posix_memalign(&p, 256, bytes ? bytes : 1)) { set_error("emu: workspace allocation failed"); return nullptr; }
    memset(p, 0xff, bytes);                                       // NaN-fill: reads of never-written workspace show up in the comparison
    g_arena_blocks.push_back(p);
    return p;
}
void Arena::release() { arena_drop(); base = nullptr; cap = off = 0; }

// models whose kernels need hardware features the emulation does not have: the GGUF reader references them, nothing calls them here
#ifndef B2EMU_HAVE_KOKORO
int Kokoro::assign(const char *, int, int, const int64_t *, const void *, size_t) { return 1; }
int Kokoro::prepare() { return 1; }
#endif
#ifndef B2EMU_HAVE_DAC
int Dac::assign(const char *, int, int, const int64_t *, const void *, size_t) { return 1; }
int Dac::prepare() { return 1; }
int Snac::assign(const char *, int, int, const int64_t *, const void *, size_t) { return 1; }
int Snac::prepare() { return 1; }
#endif

}  // namespace b2
