// Fiber scheduler of the kernel emulator (see shim/cuda_runtime.h): every CUDA thread of a CTA is a ucontext fiber on ONE
// OS thread; CTAs of a grid run one after the other.  A fiber runs until it reaches a barrier or a shuffle, then yields;
// the scheduler resumes fibers round-robin and releases a barrier when all expected threads have arrived.  Semantics
// modelled: __syncthreads (all live threads of the CTA), bar.sync id, n (named barriers), __syncwarp, __shfl_xor_sync
// (all live lanes of the warp take part), shared memory (one buffer per CTA), cp.async (immediate copy).  NOT modelled:
// memory-ordering hazards between barriers (fibers interleave only at yield points, so a missing barrier is caught only
// if it changes results under this particular interleaving), bank conflicts, timing.
#include <ucontext.h>

#include <cstdio>
#include <stdexcept>
#include <vector>

#include "cuda_runtime.h"

namespace kemu {
namespace {
constexpr size_t kStack = 256 * 1024;
constexpr int kMaxBar = 16;

struct Bar {
  int arrived = 0;
  uint64_t gen = 0;
};
struct Fiber {
  ucontext_t ctx;
  std::vector<unsigned char> stack;
  ThreadCtx tc;
  bool done = false;
  // what the fiber waits for: nullptr = runnable
  Bar* wait = nullptr;
  uint64_t wait_gen = 0;
};
struct Warp {
  Bar bar;
  uint64_t slot[32];
  int live = 0;
};
struct Cta {
  std::vector<Fiber> fibers;
  std::vector<Warp> warps;
  Bar bars[kMaxBar];
  int live = 0;
  std::vector<unsigned char> smem_store;
  unsigned char* smem = nullptr;
  const std::function<void()>* body = nullptr;
  ucontext_t sched;
  int running = -1;
};
Cta* g_cta = nullptr;
ThreadCtx g_idle;

void arrive_and_wait(Bar& b, int expected) {
  Cta& c = *g_cta;
  Fiber& f = c.fibers[c.running];
  if (++b.arrived >= expected) {
    b.arrived = 0;
    ++b.gen;  // releases everyone recorded on the previous generation
    return;
  }
  f.wait = &b;
  f.wait_gen = b.gen;
  swapcontext(&f.ctx, &c.sched);
}

void fiber_main() {
  Cta& c = *g_cta;
  (*c.body)();
  Fiber& f = c.fibers[c.running];
  f.done = true;
  --c.live;
  Warp& w = c.warps[c.running >> 5];
  --w.live;
  // an exited thread no longer counts for "all live threads" barriers: release those that are now complete
  Bar* open[2] = {&c.bars[0], &w.bar};
  const int expected[2] = {c.live, w.live};
  for (int i = 0; i < 2; ++i)
    if (open[i]->arrived > 0 && open[i]->arrived >= expected[i]) {
      open[i]->arrived = 0;
      ++open[i]->gen;
    }
  swapcontext(&f.ctx, &c.sched);
}
}  // namespace

ThreadCtx& cur() { return g_cta && g_cta->running >= 0 ? g_cta->fibers[g_cta->running].tc : g_idle; }
unsigned char* dyn_smem() { return g_cta ? g_cta->smem : nullptr; }

void barrier(int id, int count) {
  if (!g_cta || id < 0 || id >= kMaxBar) throw std::runtime_error("kernel_emu: bad barrier");
  arrive_and_wait(g_cta->bars[id], count > 0 ? count : g_cta->live);
}
void warp_barrier() {
  Warp& w = g_cta->warps[g_cta->running >> 5];
  arrive_and_wait(w.bar, w.live);
}
uint64_t shfl_xor_bits(uint64_t v, int lane_mask) {
  Warp& w = g_cta->warps[g_cta->running >> 5];
  const int lane = g_cta->running & 31;
  w.slot[lane] = v;
  arrive_and_wait(w.bar, w.live);
  const uint64_t r = w.slot[(lane ^ lane_mask) & 31];
  arrive_and_wait(w.bar, w.live);  // nobody overwrites a slot before every lane has read
  return r;
}

void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body) {
  const int nthr = (int)(block.x * block.y * block.z);
  if (nthr <= 0 || nthr > 1024) throw std::runtime_error("kernel_emu: bad block size");
  Cta cta;
  cta.smem_store.assign(smem_bytes + 64, 0);
  cta.smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(cta.smem_store.data()) + 63) & ~uintptr_t(63));
  cta.fibers.resize(nthr);
  for (auto& f : cta.fibers) f.stack.resize(kStack);
  cta.body = &body;
  const uint64_t n_blocks = (uint64_t)grid.x * grid.y * grid.z;
  for (uint64_t b = 0; b < n_blocks; ++b) {
    // fresh (poisoned) shared memory for every CTA: reads of uninitialised shared memory must not go unnoticed
    std::memset(cta.smem, 0xCB, smem_bytes);
    cta.warps.assign((nthr + 31) / 32, Warp());
    for (auto& bar : cta.bars) bar = Bar();
    cta.live = nthr;
    for (int t = 0; t < nthr; ++t) {
      Fiber& f = cta.fibers[t];
      f.done = false;
      f.wait = nullptr;
      f.tc.tid = uint3{(unsigned)t % block.x, ((unsigned)t / block.x) % block.y, (unsigned)t / (block.x * block.y)};
      f.tc.bid = uint3{(unsigned)(b % grid.x), (unsigned)((b / grid.x) % grid.y), (unsigned)(b / ((uint64_t)grid.x * grid.y))};
      f.tc.bdim = block;
      f.tc.gdim = grid;
      getcontext(&f.ctx);
      f.ctx.uc_stack.ss_sp = f.stack.data();
      f.ctx.uc_stack.ss_size = f.stack.size();
      f.ctx.uc_link = nullptr;
      makecontext(&f.ctx, fiber_main, 0);
      ++cta.warps[t >> 5].live;
    }
    g_cta = &cta;
    // KEMU_ORDER=reverse resumes the fibers from the highest thread index down: results must not depend on the order in which
    // threads run between two barriers, so a test that passes in both orders has no store/load pair that relies on it
    static const bool reverse = [] {
      const char* e = std::getenv("KEMU_ORDER");
      return e && e[0] == 'r';
    }();
    while (cta.live > 0) {
      bool progressed = false;
      for (int k = 0; k < nthr; ++k) {
        const int t = reverse ? nthr - 1 - k : k;
        Fiber& f = cta.fibers[t];
        if (f.done) continue;
        if (f.wait) {
          if (f.wait->gen == f.wait_gen) continue;  // barrier not released yet
          f.wait = nullptr;
        }
        cta.running = t;
        swapcontext(&cta.sched, &f.ctx);
        cta.running = -1;
        progressed = true;
      }
      if (!progressed) {
        g_cta = nullptr;
        throw std::runtime_error("kernel_emu: deadlock (threads wait at a barrier that can never complete)");
      }
    }
    g_cta = nullptr;
  }
}
}  // namespace kemu
