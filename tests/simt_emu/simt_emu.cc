// simt_emu.cc -- TEST INFRASTRUCTURE ONLY (see simt_emu.h): fiber scheduler, barriers, warp collectives,
// guard-paged allocations.
#include "simt_emu.h"

#include <sys/mman.h>
#include <ucontext.h>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <vector>

simt::Dim3 threadIdx, blockIdx, blockDim, gridDim;

namespace simt {
namespace {

constexpr size_t kStack = 256 * 1024;
constexpr int kNamed = 16;

struct Bar { int arrived = 0; unsigned gen = 0; };
struct Warp {
    Bar bar;
    unsigned bar_mask = 0;          // mask of the collective currently gathering (all arrivals must agree)
    uint64_t buf[32];
    int pred[32];
};
struct PendingCopy { void *dst; const void *src; size_t n; };
struct Fiber {
    ucontext_t ctx;
    unsigned char *stack = nullptr;
    bool done = false;
    unsigned tid = 0;
    const Bar *wait = nullptr;      // barrier the fiber sleeps on, released when its generation changes
    unsigned wait_gen = 0;
    std::vector<PendingCopy> copies;
};

ucontext_t g_sched;
std::vector<Fiber> g_fibers;
std::vector<Warp> g_warps;
Bar g_named[kNamed];
int g_named_expect[kNamed];
int g_live = 0;
Fiber *g_cur = nullptr;
const std::function<void()> *g_body = nullptr;
unsigned char *g_smem = nullptr, *g_smem_base = nullptr;
size_t g_smem_map = 0;
long g_launches = 0;
std::vector<unsigned char *> g_stack_pool;

[[noreturn]] void die(const char *msg) {
    fprintf(stderr, "simt_emu: %s (block %u, thread %u)\n", msg, blockIdx.x, g_cur ? g_cur->tid : 0u);
    abort();
}

void yield_to_sched() {
    Fiber *f = g_cur;
    swapcontext(&f->ctx, &g_sched);
    threadIdx.x = f->tid;           // the scheduler resumed us: restore the per-thread built-in
}

void wait_on(Bar &b) {
    Fiber *f = g_cur;
    f->wait = &b; f->wait_gen = b.gen;
    while (b.gen == f->wait_gen) yield_to_sched();
    f->wait = nullptr;
}

void arrive(Bar &b, int expected) {
    if (++b.arrived >= expected) { b.arrived = 0; ++b.gen; return; }
    wait_on(b);
}

void trampoline() {
    (*g_body)();
    Fiber *f = g_cur;
    if (!f->copies.empty()) die("thread exited with cp.async copies it never waited for");
    f->done = true;
    --g_live;
    // a thread that exits no longer takes part in __syncthreads (hardware counts only live warps' threads)
    if (g_named[0].arrived > 0 && g_named[0].arrived >= g_live) { g_named[0].arrived = 0; ++g_named[0].gen; }
    swapcontext(&f->ctx, &g_sched);
}

}  // namespace

unsigned char *dyn_smem() { return g_smem; }
long kernel_launches() { return g_launches; }

void syncthreads() { arrive(g_named[0], g_live); }

void bar_sync(int id, int nthreads) {
    if (id <= 0 || id >= kNamed) die("bar_sync: id out of range (0 is __syncthreads)");
    if (g_named[id].arrived == 0) g_named_expect[id] = nthreads;
    else if (g_named_expect[id] != nthreads) die("bar_sync: threads disagree on the thread count of a named barrier");
    arrive(g_named[id], nthreads);
}

// PTX bar.arrive: counts the thread in, does not wait (the thread may even exit afterwards)
void bar_arrive(int id, int nthreads) {
    if (id <= 0 || id >= kNamed) die("bar_arrive: id out of range");
    if (g_named[id].arrived == 0) g_named_expect[id] = nthreads;
    else if (g_named_expect[id] != nthreads) die("bar_arrive: threads disagree on the thread count of a named barrier");
    Bar &b = g_named[id];
    if (++b.arrived >= nthreads) { b.arrived = 0; ++b.gen; }
}

static void warp_gather(unsigned mask) {
    Warp &w = g_warps[g_cur->tid >> 5];
    const unsigned lane = g_cur->tid & 31u;
    if (!(mask & (1u << lane))) die("warp collective: calling lane is not in its own mask");
    if (w.bar.arrived == 0) w.bar_mask = mask;
    else if (w.bar_mask != mask) die("warp collective: lanes arrived with different masks (divergent collective)");
    arrive(w.bar, __builtin_popcount(mask));
}

void syncwarp(unsigned mask) { warp_gather(mask); }

uint64_t shfl(unsigned mask, uint64_t bits, int src_lane) {
    Warp &w = g_warps[g_cur->tid >> 5];
    const unsigned lane = g_cur->tid & 31u;
    w.buf[lane] = bits;
    warp_gather(mask);
    const uint64_t r = (src_lane >= 0 && src_lane < 32 && (mask & (1u << src_lane))) ? w.buf[src_lane] : bits;
    warp_gather(mask);              // nobody overwrites buf before everyone has read it
    return r;
}

unsigned ballot(unsigned mask, int pred) {
    Warp &w = g_warps[g_cur->tid >> 5];
    const unsigned lane = g_cur->tid & 31u;
    if (mask == (1u << lane)) return pred ? mask : 0u;       // a collective of one (__activemask() of this emulator): nothing to gather
    w.pred[lane] = pred != 0;
    warp_gather(mask);
    unsigned r = 0;
    for (int l = 0; l < 32; ++l) if ((mask & (1u << l)) && w.pred[l]) r |= 1u << l;
    warp_gather(mask);
    return r;
}

void cp_async(void *dst, const void *src, size_t n) {
    if (((uintptr_t)dst % n) || ((uintptr_t)src % n)) die("cp.async: source or destination not aligned to the copy size");
    g_cur->copies.push_back({dst, src, n});
}
void cp_async_wait_all() {
    for (const PendingCopy &c : g_cur->copies) memcpy(c.dst, c.src, c.n);
    g_cur->copies.clear();
}

// memory that ends flush (to 16 B) against an inaccessible page
static std::map<void *, std::pair<void *, size_t>> g_allocs;
void *dev_alloc(size_t bytes) {
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t body = (bytes + 15) / 16 * 16;
    const size_t map = (body + page - 1) / page * page + page;
    unsigned char *base = (unsigned char *)mmap(nullptr, map, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    if (base == MAP_FAILED) return nullptr;
    mprotect(base + map - page, page, PROT_NONE);
    unsigned char *p = base + map - page - body;
    memset(p, 0xCD, body);          // uninitialised device memory is not zero
    g_allocs[p] = {base, map};
    return p;
}
void dev_free(void *p) {
    auto it = g_allocs.find(p);
    if (it == g_allocs.end()) return;
    munmap(it->second.first, it->second.second);
    g_allocs.erase(it);
}

void launch(unsigned grid, unsigned block, size_t smem_bytes, const std::function<void()> &body) {
    if (block == 0 || block > 1024) die("launch: bad block size");
    if (smem_bytes > 227 * 1024) die("launch: more than 227 KB of dynamic shared memory");
    ++g_launches;
    static std::mt19937 rng;
    static int seeded = -1;
    if (seeded < 0) { const char *s = getenv("SIMT_EMU_SEED"); seeded = s ? atoi(s) : 0; rng.seed((unsigned)seeded); }
    gridDim.x = grid; blockDim.x = block;
    g_body = &body;
    const size_t page = (size_t)sysconf(_SC_PAGESIZE);
    const size_t sbody = (smem_bytes + 15) / 16 * 16;
    const size_t smap = (sbody + page - 1) / page * page + page;
    g_smem_base = (unsigned char *)mmap(nullptr, smap, PROT_READ | PROT_WRITE, MAP_PRIVATE | MAP_ANONYMOUS, -1, 0);
    mprotect(g_smem_base + smap - page, page, PROT_NONE);
    g_smem = g_smem_base + smap - page - sbody;
    g_smem_map = smap;
    while (g_stack_pool.size() < block) g_stack_pool.push_back((unsigned char *)malloc(kStack));
    g_fibers.resize(block);
    std::vector<unsigned> order(block);
    for (unsigned b = 0; b < grid; ++b) {
        blockIdx.x = b;
        memset(g_smem, 0xAB, sbody);                 // shared memory starts with garbage, like the hardware's
        g_warps.assign((block + 31) / 32, Warp());
        for (int i = 0; i < kNamed; ++i) { g_named[i] = Bar(); g_named_expect[i] = 0; }
        g_live = (int)block;
        for (unsigned t = 0; t < block; ++t) {
            Fiber &f = g_fibers[t];
            f.done = false; f.tid = t; f.wait = nullptr; f.copies.clear(); f.stack = g_stack_pool[t];
            getcontext(&f.ctx);
            f.ctx.uc_stack.ss_sp = f.stack; f.ctx.uc_stack.ss_size = kStack; f.ctx.uc_link = nullptr;
            makecontext(&f.ctx, trampoline, 0);
            order[t] = t;
        }
        for (;;) {
            if (seeded > 0) std::shuffle(order.begin(), order.end(), rng);
            bool progress = false, any_left = false;
            for (unsigned k = 0; k < block; ++k) {
                Fiber &f = g_fibers[order[k]];
                if (f.done) continue;
                any_left = true;
                if (f.wait && f.wait->gen == f.wait_gen) continue;      // still blocked
                g_cur = &f; threadIdx.x = f.tid;
                swapcontext(&g_sched, &f.ctx);
                progress = true;
            }
            if (!any_left) break;
            if (!progress) {                         // say who waits where before giving up
                for (unsigned t = 0; t < block; ++t) {
                    const Fiber &f = g_fibers[t];
                    if (f.done) { fprintf(stderr, "simt_emu:   thread %u exited\n", t); continue; }
                    int named = -1, warp = -1;
                    for (int i = 0; i < kNamed; ++i) if (f.wait == &g_named[i]) named = i;
                    for (size_t w = 0; w < g_warps.size(); ++w) if (f.wait == &g_warps[w].bar) warp = (int)w;
                    if (named >= 0) fprintf(stderr, "simt_emu:   thread %u waits on CTA barrier %d (%d arrived)\n", t, named, g_named[named].arrived);
                    else fprintf(stderr, "simt_emu:   thread %u waits on a collective of warp %d (mask %08x, %d arrived)\n", t, warp,
                                 g_warps[warp < 0 ? 0 : warp].bar_mask, g_warps[warp < 0 ? 0 : warp].bar.arrived);
                }
                g_cur = nullptr;
                die("deadlock: every live thread waits on a barrier that cannot complete");
            }
        }
    }
    munmap(g_smem_base, g_smem_map);
    g_smem = g_smem_base = nullptr;
    g_cur = nullptr;
}

}  // namespace simt
