// C ABI of the engine (include/qandle_b200.h): plan management + kernel launchers.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "../../include/qandle_b200.h"
#include "kernels.cuh"
#include "packed64.cuh"
#include "flat64.cuh"
#include "flat128.cuh"
#ifndef QB_KERNEL_EMU  // (TMA bulk copies: not modelled by the CPU kernel emulator)
#include "exchange.cuh"
#endif
#include "plan.h"

using namespace qb;

struct qb_plan {
  Plan p;
  int device = -1;
  int num_sms = 148;
};

namespace {

thread_local std::string g_err;

int fail(const std::string& msg) {
  g_err = msg;
  return 1;
}

#define QB_CUDA(call)                                                                         \
  do {                                                                                        \
    cudaError_t _e = (call);                                                                  \
    if (_e != cudaSuccess)                                                                    \
      return fail(std::string(#call) + " failed: " + cudaGetErrorString(_e) + " (" __FILE__ ":" + \
                  std::to_string(__LINE__) + ")");                                            \
  } while (0)

#define QB_REQUIRE(cond, msg) \
  do {                        \
    if (!(cond)) return fail(msg); \
  } while (0)

inline size_t align256(size_t x) { return (x + 255) & ~size_t(255); }

struct Workspace {
  size_t mats_shared, mats_batch, k_shared, k_batch, partials, probs_part, dyn_counter, total;
};

// Test hooks (environment, read once).  They select between code paths that all ship and are all covered by the parity suite;
// none of them is needed in normal use:
//   QB_DYN_GRID=n     cap the persistent grid of the full-tile complex64 kernels at n CTAs (the kernel emulator runs CTAs one after
//                     the other: with n = 2 the work queue is exercised at small sizes)
//   QB_DYN_TARGET=k   work items per resident slot the persistent launch aims for (default 16)
//   QB_ADJ_STREAM=0   adjoint sweeps on the generic kernel (psi and lambda both in registers, 2 CTAs / SM) instead of the
//                     streaming kernel -- the path partial tiles, parametrised diagonals and sharded plans take anyway
//   QB_EXCHANGE_TMA=0 the push exchange of sharded states with 16-byte loads / stores instead of TMA bulk copies (A/B)
//   QB_FUSE=0         separate |0...0> / MeasureProbability / adjoint-seed passes instead of the ones fused into the first and last
//                     sweeps (A/B of the HBM traffic)
struct Hooks {
  int64_t dyn_grid_cap = int64_t(1) << 40;
  int64_t dyn_target = 16;
  bool adj_stream = true;
  bool fuse = true;
  bool exchange_tma = true;
  Hooks() {
    if (const char* e = std::getenv("QB_DYN_GRID")) dyn_grid_cap = std::max<int64_t>(1, std::atoll(e));
    if (const char* e = std::getenv("QB_DYN_TARGET")) dyn_target = std::max<int64_t>(1, std::atoll(e));
    if (const char* e = std::getenv("QB_ADJ_STREAM")) adj_stream = e[0] != '0';
    if (const char* e = std::getenv("QB_FUSE")) fuse = e[0] != '0';
    if (const char* e = std::getenv("QB_EXCHANGE_TMA")) exchange_tma = e[0] != '0';
  }
};
const Hooks& hooks() {
  static const Hooks h;
  return h;
}

}  // namespace

namespace qb {
bool flat_stream_fits(int m, int L, int n_ops, int n_kslots) {
  // three CTAs per SM for the streaming adjoint kernel AND the full-tile forward kernel (+ 1 KB of static shared memory when it also
  // reduces the probabilities); 1 KB per CTA is reserved by the system
  return 3 * (fl::flat_smem_bytes(m, L, n_ops, n_kslots, fl::kStreamStages, true, true) + 1024) <= 228 * 1024 &&
         3 * (fl::flat_smem_bytes(m, L, n_ops, 0, fl::kStreamStages, false, true) + 2048) <= 228 * 1024;
}
}  // namespace qb

namespace {
constexpr int kDynMinCap = 8;  // the persistent mode may always split a sample into this many work items

// CTAs per sample of a static sweep launch.  A CTA walks the tiles c, c + cps, ... of ONE sample (its fused 2x2s are per sample), so
// the launch is `B * cps` CTAs of ceil(n_tiles / cps) tiles each on `num_sms * resident` slots: the sweep takes
// ceil(grid / slots) waves.  Pick the cps with the smallest modelled makespan waves * (tiles per CTA + per-CTA setup), the setup
// (stage tables, matrices) counted as a fraction of a tile.  ("About two waves", cps = ceil(2 slots / B), ignored the rounding:
// 896 CTAs on 888 slots run THREE waves)
int choose_cps(const qb_plan* plan, int64_t B, int n_tiles_log2, int resident_per_sm, int64_t cap = int64_t(1) << 30) {
  const int64_t n_tiles = int64_t(1) << n_tiles_log2;
  const int64_t slots = (int64_t)plan->num_sms * resident_per_sm;
  const int64_t hi = std::max<int64_t>(1, std::min(std::min(n_tiles, cap), (8 * slots + B - 1) / B));  // at most ~8 waves of CTAs
  const double setup = 0.35;
  int64_t best = 1;
  double best_cost = 1e300;
  for (int64_t cps = 1; cps <= hi; ++cps) {
    const int64_t waves = (B * cps + slots - 1) / slots;
    const int64_t tiles = (n_tiles + cps - 1) / cps;
    const double cost = (double)waves * ((double)tiles + setup);
    if (cost < best_cost * 0.999) {  // ties go to fewer, longer CTAs
      best_cost = cost;
      best = cps;
    }
  }
  return (int)best;
}

int64_t max_tiles_per_sample(const qb_plan* plan) {
  int64_t t = 1;
  for (const Sweep& sw : plan->p.sweeps) t = std::max<int64_t>(t, int64_t(1) << (plan->p.n_local - (int)sw.tile_bits.size()));
  return t;
}

// upper bound of the CTAs / work items per sample of any launch: sizes the per-item partial buffers
int max_cps(const qb_plan* plan, int64_t B) {
  const int64_t stat = std::max<int64_t>(1, ((int64_t)plan->num_sms * 8 * 2 + B - 1) / B);
  const int64_t dyn = std::min(max_tiles_per_sample(plan),
                               std::max<int64_t>(kDynMinCap, (hooks().dyn_target * plan->num_sms * 3 + B - 1) / B));
  return (int)std::max(stat, dyn);
}

// persistent mode: work items of about 1/32 of a slot's share of the sweep (at least one tile), but enough of them -- about
// `dyn_target` per slot -- that a small batch of large states still fills the machine (config 3: 16 samples x 4096 tiles)
int dyn_cps(const qb_plan* plan, int64_t B, int n_tiles_log2, int resident_per_sm) {
  const int64_t n_tiles = int64_t(1) << n_tiles_log2;
  const int64_t slots = (int64_t)plan->num_sms * resident_per_sm;
  const int64_t per_slot = std::max<int64_t>(1, B * n_tiles / slots);
  const int64_t tiles_per_item = std::max<int64_t>(1, per_slot / 32);
  int64_t cps = (n_tiles + tiles_per_item - 1) / tiles_per_item;
  const int64_t cap = std::max<int64_t>(kDynMinCap, (hooks().dyn_target * slots + B - 1) / B);
  cps = std::min(std::min(cps, cap), std::min<int64_t>(n_tiles, max_cps(plan, B)));
  return (int)std::max<int64_t>(1, cps);
}

Workspace layout(const qb_plan* plan, int64_t B) {
  const Plan& p = plan->p;
  const size_t szT = p.dtype == QB_C64 ? 4 : 8;
  Workspace w{};
  size_t off = 0;
  w.mats_shared = off;
  off += align256((size_t)std::max(p.n_groups_shared, 1) * 8 * szT);
  w.mats_batch = off;
  off += align256((size_t)std::max(p.n_groups_batch, 1) * 8 * szT * B);
  w.k_shared = off;
  off += align256((size_t)std::max(p.n_k_shared, 1) * kAcc * szT);
  w.k_batch = off;
  off += align256((size_t)std::max(p.n_k_batch, 1) * kAcc * szT * B);
  w.partials = off;
  off += align256((size_t)B * max_cps(plan, B) * std::max(p.max_kslots, 1) * kAcc * szT);
  w.probs_part = off;
  off += align256((size_t)B * max_cps(plan, B) * kProbPartStride * sizeof(double));
  w.dyn_counter = off;
  off += 256;
  w.total = off;
  return w;
}

void fill_args(const qb_plan* plan, const Sweep& sw, SweepArgs& A, int64_t B, void* state, void* lam, void* ws_base,
               int rank, bool backward) {
  const Plan& p = plan->p;
  Workspace w = layout(plan, B);
  char* ws = reinterpret_cast<char*>(ws_base);
  std::memset(&A, 0, sizeof(A));
  A.psi = state;
  A.lam = lam;
  A.ops = sw.d_ops;
  A.mats_shared = ws + w.mats_shared;
  A.mats_batch = ws + w.mats_batch;
  A.partials = ws + w.partials;
  A.rank_bits = (uint64_t)rank << p.n_local;
  A.n_ops = (int)sw.ops.size();
  A.n_groups_batch = p.n_groups_batch;
  A.n_kslots = (int)sw.kslots.size();
  A.m = (int)sw.tile_bits.size();
  A.L = std::min(p.low_bits, A.m);
  A.n_local = p.n_local;
  A.need_tile_dot = backward ? sw.has_ext_diag_param : 0;
  for (size_t i = 0; i < sw.tile_bits.size(); ++i) A.tile_bits[i] = (int8_t)sw.tile_bits[i];
  for (size_t i = 0; i < sw.nontile_bits.size(); ++i) A.nontile_bits[i] = (int8_t)sw.nontile_bits[i];
}

#ifdef QB_KERNEL_EMU  // tests/kernel_emu only: let a test assert which path a call took
static int g_emu_fused_inits = 0, g_emu_fused_probs = 0, g_emu_fused_seeds = 0, g_emu_stream_launches = 0, g_emu_dyn_launches = 0;
extern "C" int qb_emu_fused_inits() { return g_emu_fused_inits; }
extern "C" int qb_emu_fused_probs() { return g_emu_fused_probs; }
extern "C" int qb_emu_fused_seeds() { return g_emu_fused_seeds; }
extern "C" int qb_emu_stream_launches() { return g_emu_stream_launches; }
extern "C" int qb_emu_dyn_launches() { return g_emu_dyn_launches; }
#define QB_EMU_COUNT(x) (++(x))
#else
#define QB_EMU_COUNT(x) ((void)0)
#endif

bool flat64_sweep(const Plan& p, const Sweep& sw) { return p.dtype == QB_C64 && p.packed && !sw.stages.empty() && sw.stages[0].flat; }

// A forward that starts from |0...0> skips the init pass when its first sweep runs on a flat complex64 kernel, which then builds
// its tiles in shared memory (flat64.cuh: prefetch_tile): one HBM write and one HBM read of the whole state less.
bool fuse_init_ok(const Plan& p) {
  if (!hooks().fuse || p.n_local != p.n_qubits || p.steps.empty() || p.steps[0].type != QB_STEP_SWEEP) return false;
  return flat64_sweep(p, p.sweeps[p.steps[0].index]);
}
// MeasureProbability's reduction runs inside the last forward sweep when that sweep is on a flat complex64 kernel (flat64.cuh
// FUSE); probs_partial_kernel's read of the state goes away, probs_finalize_kernel is unchanged.
bool fuse_probs_ok(const Plan& p) {
  if (!hooks().fuse || p.n_local != p.n_qubits || p.n_qubits > 48 || p.steps.empty() || p.steps.back().type != QB_STEP_SWEEP) return false;
  const Sweep& sw = p.sweeps[p.steps.back().index];
  return flat64_sweep(p, sw) && sw.tile_bits.size() <= 12;
}
// After MeasureProbability the first adjoint sweep computes lambda from the psi tile it loads (flat64.cuh FUSE):
// seed_probs_kernel's read of psi and write of lambda, and the sweep's own read of lambda, go away.
bool fuse_seed_ok(const Plan& p) { return fuse_probs_ok(p); }

template <typename T>
int launch_sweep_fwd(const qb_plan* plan, const Sweep& sw, int64_t B, void* state, void* ws, int rank, cudaStream_t st,
                     bool zero_init = false, int* probs_cps_out = nullptr) {
  StagedArgs SA;
  SweepArgs& A = SA.s;
  fill_args(plan, sw, A, B, state, nullptr, ws, rank, false);
  const bool staged = !sw.stages.empty();
  SA.stages = sw.d_stages;
  SA.n_stages = (int)sw.stages.size();
  const bool use_packed = plan->p.packed && sizeof(T) == 4;
  const bool flat = staged && sw.stages[0].flat;
  QB_REQUIRE((!zero_init && !probs_cps_out) || (flat && sizeof(T) == 4), "fused |0...0> / probabilities need a flat complex64 sweep");
  // full complex64 tiles run on the persistent kernel with the 16-stage tables (flat64.cuh); more stages: the generic flat kernel
  const bool full_fwd = flat && sizeof(T) == 4 && use_packed && A.m == 12 && SA.n_stages <= fl::kStreamStages;
  const size_t smem = !staged                    ? sweep_smem_bytes(A.m, A.L, A.n_ops, 0, false, sizeof(T))
                      : flat && sizeof(T) == 4   ? fl::flat_smem_bytes(A.m, A.L, A.n_ops, 0, SA.n_stages, false, full_fwd)
                      : flat                     ? fd::flat128_smem_bytes(A.m, A.L, A.n_ops, 0, SA.n_stages, false)
                      : use_packed ? pk::packed_smem_bytes(A.m, A.L, A.n_ops, 0, SA.n_stages, false)
                                   : staged_smem_bytes(A.m, A.L, A.n_ops, 0, SA.n_stages, false, sizeof(T));
  QB_REQUIRE(smem <= 226 * 1024, "sweep needs more than 226 KB of shared memory");
  // (228 KB per SM, 1 KB reserved per CTA; the fused probability reduction adds 816 B of static shared memory)
  const int resident = (int)std::max<size_t>(1, std::min<size_t>(staged ? 3 : 8, (228 * 1024) / (smem + 1024 + (probs_cps_out ? 1024 : 0))));
  A.cps = probs_cps_out ? choose_cps(plan, B, A.n_local - A.m, resident, max_cps(plan, B))  // one row of probs_part per CTA
                        : choose_cps(plan, B, A.n_local - A.m, resident);
  int64_t grid = B * A.cps;
  QB_REQUIRE(grid < (int64_t(1) << 31), "grid too large");
  if (flat && sizeof(T) == 8) {
    pk::PackedArgs PA;
    PA.s = A;
    PA.stages = sw.d_stages;
    PA.n_stages = SA.n_stages;
    if (A.m == 11)
      fd::sweep_flat128_kernel<false, true><<<(unsigned)grid, fd::flat128_threads(A.m, A.L), smem, st>>>(PA);
    else
      fd::sweep_flat128_kernel<false><<<(unsigned)grid, fd::flat128_threads(A.m, A.L), smem, st>>>(PA);
  } else if (staged && use_packed) {
    pk::PackedArgs PA;
    PA.s = A;
    PA.stages = sw.d_stages;
    PA.n_stages = SA.n_stages;
    PA.zero_init = zero_init ? 1 : 0;
    if (probs_cps_out) PA.probs_part = reinterpret_cast<double*>(static_cast<char*>(ws) + layout(plan, B).probs_part);
    if (full_fwd) {
      // full tiles: persistent CTAs, one per resident slot, work items from the queue
      A.cps = PA.s.cps = dyn_cps(plan, B, A.n_local - A.m, resident);
      grid = B * A.cps;
      QB_REQUIRE(grid < (int64_t(1) << 31), "too many work items");
      PA.dyn_items = (int32_t)grid;
      PA.dyn_counter = reinterpret_cast<int32_t*>(static_cast<char*>(ws) + layout(plan, B).dyn_counter);
      QB_CUDA(cudaMemsetAsync(PA.dyn_counter, 0, sizeof(int32_t), st));
      const int64_t pgrid = std::min(std::min<int64_t>(grid, (int64_t)plan->num_sms * resident), hooks().dyn_grid_cap);
      QB_EMU_COUNT(g_emu_dyn_launches);
      if (probs_cps_out)
        fl::sweep_flat_kernel<false, true, false, true, true><<<(unsigned)pgrid, fl::flat_threads(A.m, A.L), smem, st>>>(PA);
      else
        fl::sweep_flat_kernel<false, true, false, true><<<(unsigned)pgrid, fl::flat_threads(A.m, A.L), smem, st>>>(PA);
    } else if (flat) {  // one thread per 16 amplitudes of the tile (at most 256: the planner keeps flat tiles at <= 2^12)
      fl::sweep_flat_kernel<false><<<(unsigned)grid, fl::flat_threads(A.m, A.L), smem, st>>>(PA);
    } else {
      pk::sweep_packed_kernel<false><<<(unsigned)grid, kSweepThreads, smem, st>>>(PA);
    }
  } else if (staged) {
    sweep_staged_kernel<T, false><<<(unsigned)grid, kSweepThreads, smem, st>>>(SA);
  } else {
    sweep_forward_kernel<T><<<(unsigned)grid, kSweepThreads, smem, st>>>(A);
  }
  QB_CUDA(cudaGetLastError());
  if (probs_cps_out) *probs_cps_out = A.cps;
  return 0;
}

template <typename T>
int launch_sweep_bwd(const qb_plan* plan, const Sweep& sw, int64_t B, void* state, void* lam, void* ws_base, int rank,
                     cudaStream_t st, const void* seed_grad = nullptr, const void* psi_src = nullptr) {
  const Plan& p = plan->p;
  StagedArgs SA;
  SweepArgs& A = SA.s;
  fill_args(plan, sw, A, B, state, lam, ws_base, rank, true);
  const bool staged = !sw.stages.empty();
  SA.stages = sw.d_stages;
  SA.n_stages = (int)sw.stages.size();
  const bool use_packed = plan->p.packed && sizeof(T) == 4;
  const bool flat = staged && sw.stages[0].flat;
  if (flat) {  // the adjoint sweep's own linearisation (plan.h: Sweep::ops_bwd)
    A.ops = sw.d_ops_bwd;
    SA.stages = sw.d_stages_bwd;
    SA.n_stages = (int)sw.stages_bwd.size();
  }
  QB_REQUIRE(!seed_grad || (flat && sizeof(T) == 4 && A.m <= 12), "fused adjoint seed needs a flat complex64 sweep");
  QB_REQUIRE(!psi_src || (flat && sizeof(T) == 4 && use_packed), "an out-of-place adjoint sweep needs a flat complex64 sweep");
  // streaming adjoint kernel: full tiles, small stage tables, no gradient-carrying diagonal (flat64.cuh), and the
  // shared memory of three CTAs must fit one SM -- otherwise the generic kernel runs
  // (amplitude-sharded plans too: the rank bits enter through gbase like any out-of-tile bit; tests/test_kernel_emu_sharded.py)
  bool stream = flat && sizeof(T) == 4 && use_packed && hooks().adj_stream && A.m == 12 && SA.n_stages <= fl::kStreamStages &&
                !A.need_tile_dot;
  if (stream) {
    for (const KOp& o : sw.ops_bwd)
      if ((o.kind == K_D1 || o.kind == K_D1_EXT) && o.kslot >= 0) stream = false;
    if (!flat_stream_fits(A.m, A.L, A.n_ops, A.n_kslots)) stream = false;
  }
  const size_t smem = !staged                    ? sweep_smem_bytes(A.m, A.L, A.n_ops, A.n_kslots, true, sizeof(T))
                      : stream                   ? fl::flat_smem_bytes(A.m, A.L, A.n_ops, A.n_kslots, SA.n_stages, true, true)
                      : flat && sizeof(T) == 4   ? fl::flat_smem_bytes(A.m, A.L, A.n_ops, A.n_kslots, SA.n_stages, true)
                      : flat                     ? fd::flat128_smem_bytes(A.m, A.L, A.n_ops, A.n_kslots, SA.n_stages, true)
                      : use_packed ? pk::packed_smem_bytes(A.m, A.L, A.n_ops, A.n_kslots, SA.n_stages, true)
                                   : staged_smem_bytes(A.m, A.L, A.n_ops, A.n_kslots, SA.n_stages, true, sizeof(T));
  QB_REQUIRE(smem <= 227 * 1024, "backward sweep needs more than 227 KB of shared memory");
  const int resident = (int)std::max<size_t>(1, std::min<size_t>(stream ? 3 : (staged ? 2 : 8), (227 * 1024) / (smem + 1024)));
  A.cps = choose_cps(plan, B, A.n_local - A.m, resident, max_cps(plan, B));
  int64_t grid = B * A.cps;
  QB_REQUIRE(grid < (int64_t(1) << 31), "grid too large");
  if (flat && sizeof(T) == 8) {
    pk::PackedArgs PA;
    PA.s = A;
    PA.stages = SA.stages;
    PA.n_stages = SA.n_stages;
    if (A.m == 11)
      fd::sweep_flat128_kernel<true, true><<<(unsigned)grid, fd::flat128_threads(A.m, A.L), smem, st>>>(PA);
    else
      fd::sweep_flat128_kernel<true><<<(unsigned)grid, fd::flat128_threads(A.m, A.L), smem, st>>>(PA);
  } else if (staged && use_packed) {
    pk::PackedArgs PA;
    PA.s = A;
    PA.stages = SA.stages;
    PA.n_stages = SA.n_stages;
    PA.psi_src = psi_src;
    QB_REQUIRE(!psi_src || !stream || seed_grad, "an out-of-place streaming adjoint sweep runs on the seed-fused instantiation");
    if (seed_grad) {
      PA.seed_grad = static_cast<const float*>(seed_grad);
      PA.seed_final_pos = p.d_final_pos;
      PA.seed_n_qubits = p.n_qubits;
      for (int j = 0; j < A.m; ++j) {
        int q = 0;
        while (q < p.n_qubits && p.final_pos[q] != sw.tile_bits[j]) ++q;
        QB_REQUIRE(q < p.n_qubits, "tile bit without a qubit in the final layout");
        PA.seed_tile_q[j] = (int8_t)q;
      }
    }
    if (stream) {
      QB_EMU_COUNT(g_emu_stream_launches);
      QB_EMU_COUNT(g_emu_dyn_launches);
      A.cps = PA.s.cps = dyn_cps(plan, B, A.n_local - A.m, resident);
      grid = B * A.cps;
      QB_REQUIRE(grid < (int64_t(1) << 31), "too many work items");
      PA.dyn_items = (int32_t)grid;
      PA.dyn_counter = reinterpret_cast<int32_t*>(static_cast<char*>(ws_base) + layout(plan, B).dyn_counter);
      QB_CUDA(cudaMemsetAsync(PA.dyn_counter, 0, sizeof(int32_t), st));
      const int64_t pgrid = std::min(std::min<int64_t>(grid, (int64_t)plan->num_sms * resident), hooks().dyn_grid_cap);
      if (seed_grad)
        fl::sweep_flat_kernel<true, true, true, true, true><<<(unsigned)pgrid, fl::flat_threads(A.m, A.L), smem, st>>>(PA);
      else
        fl::sweep_flat_kernel<true, true, true, true><<<(unsigned)pgrid, fl::flat_threads(A.m, A.L), smem, st>>>(PA);
    } else if (flat) {
      fl::sweep_flat_kernel<true><<<(unsigned)grid, fl::flat_threads(A.m, A.L), smem, st>>>(PA);
    } else {
      pk::sweep_packed_kernel<true><<<(unsigned)grid, kSweepThreads, smem, st>>>(PA);
    }
  } else if (staged) {
    sweep_staged_kernel<T, true><<<(unsigned)grid, kSweepThreads, smem, st>>>(SA);
  } else {
    sweep_backward_kernel<T><<<(unsigned)grid, kSweepThreads, smem, st>>>(A);
  }
  QB_CUDA(cudaGetLastError());
  if (A.n_kslots > 0) {
    Workspace w = layout(plan, B);
    char* ws = reinterpret_cast<char*>(ws_base);
    reduce_partials_kernel<T><<<A.n_kslots, 256, 0, st>>>(reinterpret_cast<const T*>(ws + w.partials), sw.d_kslots,
                                                          A.n_kslots, (int)B, A.cps, reinterpret_cast<T*>(ws + w.k_shared),
                                                          reinterpret_cast<T*>(ws + w.k_batch), p.n_k_batch);
    QB_CUDA(cudaGetLastError());
  }
  return 0;
}

inline unsigned ew_grid(uint64_t total, int num_sms) {
  uint64_t g = (total + 255) / 256;
  uint64_t cap = (uint64_t)num_sms * 16;
  return (unsigned)std::max<uint64_t>(1, std::min(g, cap));
}

template <typename K>
int opt_in(K kernel) {
  cudaFuncAttributes fa;
  QB_CUDA(cudaFuncGetAttributes(&fa, kernel));
  QB_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024 - (int)fa.sharedSizeBytes));
  return 0;
}

int opt_in_shared_memory() {
#define QB_OPT_IN(...)                        \
  do {                                        \
    if (int rc = opt_in(__VA_ARGS__)) return rc; \
  } while (0)
  QB_OPT_IN(sweep_forward_kernel<float>);
  QB_OPT_IN(sweep_forward_kernel<double>);
  QB_OPT_IN(sweep_backward_kernel<float>);
  QB_OPT_IN(sweep_backward_kernel<double>);
  QB_OPT_IN(sweep_staged_kernel<float, false>);
  QB_OPT_IN(sweep_staged_kernel<float, true>);
  QB_OPT_IN(sweep_staged_kernel<double, false>);
  QB_OPT_IN(sweep_staged_kernel<double, true>);
  QB_OPT_IN(pk::sweep_packed_kernel<false>);
  QB_OPT_IN(pk::sweep_packed_kernel<true>);
  // the six instantiations of the flat complex64 kernel: forward / adjoint x {generic, full-tile persistent, the same + fused measurement}
  QB_OPT_IN(fl::sweep_flat_kernel<false>);
  QB_OPT_IN(fl::sweep_flat_kernel<false, true, false, true>);
  QB_OPT_IN(fl::sweep_flat_kernel<false, true, false, true, true>);
  QB_OPT_IN(fl::sweep_flat_kernel<true>);
  QB_OPT_IN(fl::sweep_flat_kernel<true, true, true, true>);
  QB_OPT_IN(fl::sweep_flat_kernel<true, true, true, true, true>);
  QB_OPT_IN(fd::sweep_flat128_kernel<false>);
  QB_OPT_IN(fd::sweep_flat128_kernel<true>);
  QB_OPT_IN(fd::sweep_flat128_kernel<false, true>);
  QB_OPT_IN(fd::sweep_flat128_kernel<true, true>);
#undef QB_OPT_IN
  return 0;
}

int upload_plan(qb_plan* plan) {
  Plan& p = plan->p;
  QB_CUDA(cudaGetDevice(&plan->device));
  cudaDeviceProp prop;
  QB_CUDA(cudaGetDeviceProperties(&prop, plan->device));
  plan->num_sms = prop.multiProcessorCount;
  QB_REQUIRE(prop.major >= 10, "qandle_b200 kernels are built for sm_100a (Blackwell); found an older device");
  if (!p.members.empty()) {
    QB_CUDA(cudaMalloc(&p.d_members, p.members.size() * sizeof(Member)));
    QB_CUDA(cudaMemcpy(p.d_members, p.members.data(), p.members.size() * sizeof(Member), cudaMemcpyHostToDevice));
  }
  if (!p.groups.empty()) {
    QB_CUDA(cudaMalloc(&p.d_groups, p.groups.size() * sizeof(Group)));
    QB_CUDA(cudaMemcpy(p.d_groups, p.groups.data(), p.groups.size() * sizeof(Group), cudaMemcpyHostToDevice));
  }
  QB_CUDA(cudaMalloc(&p.d_final_pos, p.final_pos.size() * sizeof(int32_t)));
  QB_CUDA(cudaMemcpy(p.d_final_pos, p.final_pos.data(), p.final_pos.size() * sizeof(int32_t), cudaMemcpyHostToDevice));
  for (Sweep& sw : p.sweeps) {
    if (!sw.ops.empty()) {
      QB_CUDA(cudaMalloc(&sw.d_ops, sw.ops.size() * sizeof(KOp)));
      QB_CUDA(cudaMemcpy(sw.d_ops, sw.ops.data(), sw.ops.size() * sizeof(KOp), cudaMemcpyHostToDevice));
    }
    if (!sw.kslots.empty()) {
      QB_CUDA(cudaMalloc(&sw.d_kslots, sw.kslots.size() * sizeof(KSlot)));
      QB_CUDA(cudaMemcpy(sw.d_kslots, sw.kslots.data(), sw.kslots.size() * sizeof(KSlot), cudaMemcpyHostToDevice));
    }
    if (!sw.stages.empty()) {
      QB_CUDA(cudaMalloc(&sw.d_stages, sw.stages.size() * sizeof(Stage)));
      QB_CUDA(cudaMemcpy(sw.d_stages, sw.stages.data(), sw.stages.size() * sizeof(Stage), cudaMemcpyHostToDevice));
    }
    if (!sw.ops_bwd.empty()) {
      QB_CUDA(cudaMalloc(&sw.d_ops_bwd, sw.ops_bwd.size() * sizeof(KOp)));
      QB_CUDA(cudaMemcpy(sw.d_ops_bwd, sw.ops_bwd.data(), sw.ops_bwd.size() * sizeof(KOp), cudaMemcpyHostToDevice));
      QB_CUDA(cudaMalloc(&sw.d_stages_bwd, sw.stages_bwd.size() * sizeof(Stage)));
      QB_CUDA(cudaMemcpy(sw.d_stages_bwd, sw.stages_bwd.data(), sw.stages_bwd.size() * sizeof(Stage), cudaMemcpyHostToDevice));
    }
  }
  // opt in to > 48 KB dynamic shared memory once per device (the limit covers static + dynamic shared memory)
  if (int rc = opt_in_shared_memory()) return rc;
  return 0;
}

template <typename T>
int prepare_T(const qb_plan* plan, int64_t B, const void* sa, const void* ba, int ncols, const void* fm, void* ws_base,
              cudaStream_t st) {
  const Plan& p = plan->p;
  if (p.groups.empty()) return 0;
  Workspace w = layout(plan, B);
  char* ws = reinterpret_cast<char*>(ws_base);
  const int64_t Beff = p.n_groups_batch > 0 ? B : 1;
  const int64_t total = (int64_t)p.groups.size() * Beff;
  build_mats_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
      p.d_groups, p.d_members, (int)p.groups.size(), reinterpret_cast<const T*>(sa), reinterpret_cast<const T*>(ba), ncols,
      reinterpret_cast<const T*>(fm), reinterpret_cast<T*>(ws + w.mats_shared), reinterpret_cast<T*>(ws + w.mats_batch),
      p.n_groups_batch, Beff);
  QB_CUDA(cudaGetLastError());
  return 0;
}

template <typename T>
int finalize_T(const qb_plan* plan, int64_t B, const void* sa, const void* ba, int ncols, const void* fm, void* ws_base,
               void* gs, int n_shared, void* gb, cudaStream_t st) {
  const Plan& p = plan->p;
  if (gs && n_shared > 0) QB_CUDA(cudaMemsetAsync(gs, 0, (size_t)n_shared * sizeof(T), st));
  if (gb && ncols > 0) QB_CUDA(cudaMemsetAsync(gb, 0, (size_t)B * ncols * sizeof(T), st));
  if (p.groups.empty()) return 0;
  Workspace w = layout(plan, B);
  char* ws = reinterpret_cast<char*>(ws_base);
  const int64_t Beff = p.n_groups_batch > 0 ? B : 1;
  const int64_t total = (int64_t)p.groups.size() * Beff;
  finalize_grads_kernel<T><<<(unsigned)((total + 127) / 128), 128, 0, st>>>(
      p.d_groups, p.d_members, (int)p.groups.size(), reinterpret_cast<const T*>(sa), reinterpret_cast<const T*>(ba), ncols,
      reinterpret_cast<const T*>(fm), reinterpret_cast<const T*>(ws + w.k_shared), reinterpret_cast<const T*>(ws + w.k_batch),
      p.n_k_batch, reinterpret_cast<T*>(gs), reinterpret_cast<T*>(gb), Beff);
  QB_CUDA(cudaGetLastError());
  return 0;
}

int check_plan(const qb_plan* plan, int64_t B) {
  QB_REQUIRE(plan != nullptr, "plan is NULL");
  QB_REQUIRE(!plan->p.host_only, "plan was created with host_only=1: it cannot launch kernels (no CPU fallback exists)");
  QB_REQUIRE(B >= 1, "batch must be >= 1");
  // a plan owns device-resident tables and per-device kernel attributes: it must run on the device it was created on
  int dev = -1;
  QB_CUDA(cudaGetDevice(&dev));
  if (dev != plan->device)
    return fail("plan was created on CUDA device " + std::to_string(plan->device) + " but the current device is " + std::to_string(dev) +
                " (create one plan per device)");
  return 0;
}

int probs_cps(const qb_plan* plan, int64_t B) {
  const Plan& p = plan->p;
  const int64_t n_seg = p.n_local > kProbSegBits ? (int64_t(1) << (p.n_local - kProbSegBits)) : 1;
  int64_t cps = ((int64_t)plan->num_sms * 8 + B - 1) / B;
  cps = std::max<int64_t>(1, std::min(cps, n_seg));
  return (int)std::min<int64_t>(cps, max_cps(plan, B));
}

}  // namespace

// =====================================================================================================
extern "C" {

const char* qb_last_error(void) { return g_err.c_str(); }
const char* qb_version(void) { return "qandle_b200 0.1 (sm_100a)"; }

int qb_plan_create(const int32_t* program, int32_t n_gates, int32_t n_qubits, int32_t dtype, const qb_plan_opts* opts,
                   qb_plan** out) {
  QB_REQUIRE(out != nullptr, "out is NULL");
  *out = nullptr;
  QB_REQUIRE(n_gates >= 0, "n_gates < 0");
  QB_REQUIRE(program != nullptr || n_gates == 0, "program is NULL");
  std::vector<GateIn> gates((size_t)n_gates);
  for (int i = 0; i < n_gates; ++i) {
    const int32_t* r = program + 4 * (size_t)i;
    gates[i].kind = r[0] & QB_OP_MASK;
    gates[i].batch = (r[0] & QB_FLAG_BATCH) ? 1 : 0;
    gates[i].q0 = r[1];
    gates[i].q1 = r[2];
    gates[i].slot = r[3];
  }
  PlanOptions po;
  if (opts) {
    po.tile_bits = opts->tile_bits;
    po.low_bits = opts->low_bits;
    po.fuse = opts->fuse < 0 ? 0 : 1;
    po.n_local = opts->n_local;
    po.host_only = opts->host_only;
    po.swap_relabel = opts->swap_relabel < 0 ? 0 : 1;
    po.final_layout = opts->final_layout;
    po.max_ops_per_sweep = opts->max_ops_per_sweep;
    po.staged = opts->staged < 0 ? 0 : 1;
    po.packed = opts->packed < 0 ? 0 : 1;
    po.flat = opts->flat < 0 ? 0 : 1;
    po.narrow_sync = opts->narrow_sync < 0 ? 0 : 1;
    po.exchange_any_bit = opts->exchange_any_bit > 0 ? 1 : 0;
    po.trim_search = opts->sweep_search < 0 ? 0 : 1;
  }
  qb_plan* plan = new qb_plan();
  try {
    build_plan(gates, n_qubits, dtype, po, plan->p);
  } catch (const std::exception& e) {
    delete plan;
    return fail(std::string("qb_plan_create: ") + e.what());
  }
  if (!po.host_only) {
    int rc = upload_plan(plan);
    if (rc) {
      qb_plan_destroy(plan);
      return rc;
    }
  }
  *out = plan;
  return 0;
}

void qb_plan_destroy(qb_plan* plan) {
  if (!plan) return;
  Plan& p = plan->p;
  if (!p.host_only) {
    cudaFree(p.d_members);
    cudaFree(p.d_groups);
    cudaFree(p.d_final_pos);
    for (Sweep& sw : p.sweeps) {
      cudaFree(sw.d_ops);
      cudaFree(sw.d_kslots);
      cudaFree(sw.d_stages);
      cudaFree(sw.d_ops_bwd);
      cudaFree(sw.d_stages_bwd);
    }
  }
  delete plan;
}

int32_t qb_plan_num_steps(const qb_plan* plan) { return plan ? (int32_t)plan->p.steps.size() : -1; }
int32_t qb_plan_num_sweeps(const qb_plan* plan) { return plan ? (int32_t)plan->p.sweeps.size() : -1; }
int32_t qb_plan_num_groups(const qb_plan* plan) { return plan ? (int32_t)plan->p.groups.size() : -1; }
int32_t qb_plan_step_type(const qb_plan* plan, int32_t step) {
  if (!plan || step < 0 || step >= (int32_t)plan->p.steps.size()) return -1;
  return plan->p.steps[step].type;
}
int qb_plan_final_pos(const qb_plan* plan, int32_t* pos_out) {
  QB_REQUIRE(plan && pos_out, "NULL argument");
  for (size_t i = 0; i < plan->p.final_pos.size(); ++i) pos_out[i] = plan->p.final_pos[i];
  return 0;
}
int64_t qb_plan_dump(const qb_plan* plan, int64_t* buf, int64_t cap) {
  if (!plan) return -1;
  std::vector<int64_t> v;
  dump_plan(plan->p, v);
  for (int64_t i = 0; i < (int64_t)v.size() && i < cap; ++i) buf[i] = v[i];
  return (int64_t)v.size();
}
int64_t qb_workspace_bytes(const qb_plan* plan, int64_t batch) {
  if (!plan || batch < 1) return -1;
  return (int64_t)layout(plan, batch).total;
}
int64_t qb_plan_algorithmic_bytes(const qb_plan* plan, int64_t batch, int32_t backward) {
  if (!plan) return -1;
  const Plan& p = plan->p;
  const int64_t S = (int64_t(1) << p.n_local) * (p.dtype == QB_C64 ? 8 : 16) * batch;
  return (int64_t)p.sweeps.size() * S * (backward ? 4 : 2);
}
int32_t qb_plan_num_launches(const qb_plan* plan, int32_t backward, int32_t measure) {
  if (!plan) return -1;
  const Plan& p = plan->p;
  int n = 0;
  if (!backward) {
    n += p.groups.empty() ? 0 : 1;          // build_mats
    n += (int)p.sweeps.size();              // sweeps
    if (p.sweeps.empty() || !fuse_init_ok(p)) n += 1;  // |0...0> pass (or built inside the first sweep)
    if (measure == QB_MEASURE_PROBS) n += fuse_probs_ok(p) ? 1 : 2;  // (partial sums: own pass, or inside the last sweep) + finalize
    if (measure == QB_MEASURE_JOINT) n += 1;
  } else {
    if (!(measure == QB_MEASURE_PROBS && fuse_seed_ok(p))) n += 1;  // seed pass (or built inside the first adjoint sweep)
    for (const Sweep& sw : p.sweeps) n += sw.kslots.empty() ? 1 : 2;
    n += p.groups.empty() ? 0 : 1;  // finalize
  }
  return n;
}

#define DISPATCH(plan, fn, ...) ((plan)->p.dtype == QB_C64 ? fn<float>(__VA_ARGS__) : fn<double>(__VA_ARGS__))

int qb_prepare_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                   int32_t n_batch_cols, const void* fixed_mats, void* workspace, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(p.n_shared_slots == 0 || shared_angles, "shared_angles is NULL but the program uses shared slots");
  QB_REQUIRE(p.n_batch_slots == 0 || (batch_angles && n_batch_cols >= p.n_batch_slots), "batch_angles missing / too few columns");
  QB_REQUIRE(p.n_fixed_mats == 0 || fixed_mats, "fixed_mats is NULL but the program uses U gates");
  QB_REQUIRE(workspace, "workspace is NULL");
  return DISPATCH(plan, prepare_T, plan, batch, shared_angles, batch_angles, n_batch_cols, fixed_mats, workspace,
                  (cudaStream_t)stream);
}

int qb_init_zero_dev(const qb_plan* plan, int64_t batch, void* state, int32_t rank, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  const uint64_t total = (uint64_t)batch << p.n_local;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.dtype == QB_C64)
    init_zero_kernel<float><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((float2*)state, p.n_local, batch, rank == 0);
  else
    init_zero_kernel<double><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((double2*)state, p.n_local, batch, rank == 0);
  QB_CUDA(cudaGetLastError());
  return 0;
}

int qb_apply_forward_dev(const qb_plan* plan, int32_t step_begin, int32_t step_end, int64_t batch, void* state,
                         void* workspace, int32_t rank, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(step_begin >= 0 && step_end <= (int)p.steps.size() && step_begin <= step_end, "bad step range");
  for (int s = step_begin; s < step_end; ++s) {
    QB_REQUIRE(p.steps[s].type == QB_STEP_SWEEP, "step range contains an exchange step: the caller must perform it");
    const Sweep& sw = p.sweeps[p.steps[s].index];
    int rc = DISPATCH(plan, launch_sweep_fwd, plan, sw, batch, state, workspace, rank, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int qb_measure_probs_dev(const qb_plan* plan, int64_t batch, const void* state, void* probs_out, void* workspace,
                         int32_t rank, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  Workspace w = layout(plan, batch);
  double* part = reinterpret_cast<double*>(reinterpret_cast<char*>(workspace) + w.probs_part);
  const int cps = probs_cps(plan, batch);
  cudaStream_t st = (cudaStream_t)stream;
  QB_REQUIRE(batch * cps < (int64_t(1) << 31), "grid too large");
  if (p.dtype == QB_C64) {
    probs_partial_kernel<float><<<(unsigned)(batch * cps), 256, 0, st>>>((const float2*)state, p.n_local, cps, part);
    probs_finalize_kernel<float><<<(unsigned)batch, 64, 0, st>>>(part, cps, p.n_qubits, p.n_local, p.d_final_pos, rank,
                                                                 (float*)probs_out);
  } else {
    probs_partial_kernel<double><<<(unsigned)(batch * cps), 256, 0, st>>>((const double2*)state, p.n_local, cps, part);
    probs_finalize_kernel<double><<<(unsigned)batch, 64, 0, st>>>(part, cps, p.n_qubits, p.n_local, p.d_final_pos, rank,
                                                                  (double*)probs_out);
  }
  QB_CUDA(cudaGetLastError());
  return 0;
}

int qb_measure_joint_dev(const qb_plan* plan, int64_t batch, const void* state, void* joint_out, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  const uint64_t total = (uint64_t)batch << p.n_local;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.dtype == QB_C64)
    joint_kernel<float><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((const float2*)state, (float*)joint_out, total);
  else
    joint_kernel<double><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((const double2*)state, (double*)joint_out, total);
  QB_CUDA(cudaGetLastError());
  return 0;
}

int qb_seed_probs_dev(const qb_plan* plan, int64_t batch, const void* state, const void* grad_probs, void* lambda,
                      int32_t rank, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  cudaStream_t st = (cudaStream_t)stream;
  const int64_t per_sample_ctas = std::max<int64_t>(1, (int64_t(1) << p.n_local) / (256 * 8));
  int64_t cps = ((int64_t)plan->num_sms * 8 + batch - 1) / batch;
  cps = std::max<int64_t>(1, std::min(cps, per_sample_ctas));
  QB_REQUIRE(batch * cps < (int64_t(1) << 31), "grid too large");
  if (p.dtype == QB_C64)
    seed_probs_kernel<float><<<(unsigned)(batch * cps), 256, 0, st>>>((const float2*)state, (const float*)grad_probs,
                                                                     (float2*)lambda, p.n_qubits, p.n_local, p.d_final_pos,
                                                                     rank, (int)cps);
  else
    seed_probs_kernel<double><<<(unsigned)(batch * cps), 256, 0, st>>>((const double2*)state, (const double*)grad_probs,
                                                                      (double2*)lambda, p.n_qubits, p.n_local,
                                                                      p.d_final_pos, rank, (int)cps);
  QB_CUDA(cudaGetLastError());
  return 0;
}

int qb_seed_joint_dev(const qb_plan* plan, int64_t batch, const void* state, const void* grad_joint, void* lambda,
                      void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  const uint64_t total = (uint64_t)batch << p.n_local;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.dtype == QB_C64)
    seed_joint_kernel<float><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((const float2*)state, (const float*)grad_joint,
                                                                           (float2*)lambda, total);
  else
    seed_joint_kernel<double><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((const double2*)state, (const double*)grad_joint,
                                                                            (double2*)lambda, total);
  QB_CUDA(cudaGetLastError());
  return 0;
}

int qb_seed_state_dev(const qb_plan* plan, int64_t batch, const void* grad_state, void* lambda, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  const uint64_t total = (uint64_t)batch << p.n_local;
  cudaStream_t st = (cudaStream_t)stream;
  if (p.dtype == QB_C64)
    seed_state_kernel<float><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((const float2*)grad_state, (float2*)lambda, total);
  else
    seed_state_kernel<double><<<ew_grid(total, plan->num_sms), 256, 0, st>>>((const double2*)grad_state, (double2*)lambda,
                                                                            total);
  QB_CUDA(cudaGetLastError());
  return 0;
}

int qb_backward_begin_dev(const qb_plan* plan, int64_t batch, void* workspace, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  Workspace w = layout(plan, batch);
  char* ws = reinterpret_cast<char*>(workspace);
  // k_shared and k_batch are contiguous
  QB_CUDA(cudaMemsetAsync(ws + w.k_shared, 0, w.partials - w.k_shared, (cudaStream_t)stream));
  return 0;
}

int qb_apply_backward_dev(const qb_plan* plan, int32_t step_begin, int32_t step_end, int64_t batch, void* state,
                          void* lambda, void* workspace, int32_t rank, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(step_begin >= 0 && step_end <= (int)p.steps.size() && step_begin <= step_end, "bad step range");
  for (int s = step_end - 1; s >= step_begin; --s) {
    QB_REQUIRE(p.steps[s].type == QB_STEP_SWEEP, "step range contains an exchange step: the caller must perform it");
    const Sweep& sw = p.sweeps[p.steps[s].index];
    int rc = DISPATCH(plan, launch_sweep_bwd, plan, sw, batch, state, lambda, workspace, rank, (cudaStream_t)stream);
    if (rc) return rc;
  }
  return 0;
}

int qb_finalize_grads_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                          int32_t n_batch_cols, const void* fixed_mats, void* workspace, void* grad_shared,
                          int32_t n_shared, void* grad_batch, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(n_shared >= p.n_shared_slots, "grad_shared has fewer entries than the program's shared slots");
  QB_REQUIRE(p.n_shared_slots == 0 || grad_shared, "grad_shared is NULL");
  QB_REQUIRE(p.n_batch_slots == 0 || grad_batch, "grad_batch is NULL");
  return DISPATCH(plan, finalize_T, plan, batch, shared_angles, batch_angles, n_batch_cols, fixed_mats, workspace,
                  grad_shared, n_shared, grad_batch, (cudaStream_t)stream);
}

int qb_convert_layout_dev(const qb_plan* plan, int64_t batch, void* state, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  if (p.dtype != QB_C64) return 0;  // complex128 states are interleaved internally
  const uint64_t n_units = ((uint64_t)batch << p.n_local) >> 1;
  QB_REQUIRE(p.n_local >= 1, "state too small");
  convert_layout_kernel<<<ew_grid(n_units, plan->num_sms), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<int4*>(state), n_units);
  QB_CUDA(cudaGetLastError());
  return 0;
}

int qb_forward_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                   int32_t n_batch_cols, const void* fixed_mats, int32_t init_kind, void* state, int32_t measure,
                   void* measure_out, void* workspace, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(p.n_local == p.n_qubits, "qb_forward_dev is for unsharded plans; drive sharded plans step by step");
  QB_REQUIRE(state, "state is NULL");
  if (int rc = qb_prepare_dev(plan, batch, shared_angles, batch_angles, n_batch_cols, fixed_mats, workspace, stream)) return rc;
  int first = 0;
  bool zero_first = false;  // the first sweep builds |0...0> itself
  if (init_kind == QB_INIT_ZERO) {
    if (fuse_init_ok(p)) {
      QB_EMU_COUNT(g_emu_fused_inits);
      zero_first = true;  // step 0 is launched below with the flag
    } else if (int rc = qb_init_zero_dev(plan, batch, state, 0, stream))
      return rc;
  } else {
    QB_REQUIRE(init_kind == QB_INIT_STATE, "bad init_kind");
    if (int rc = qb_convert_layout_dev(plan, batch, state, stream)) return rc;  // caller's state is interleaved
  }
  if (measure == QB_MEASURE_PROBS && fuse_probs_ok(p)) {
    QB_REQUIRE(measure_out, "measure_out is NULL");
    QB_EMU_COUNT(g_emu_fused_probs);
    const int last = (int)p.steps.size() - 1;
    int cps = 0;
    if (zero_first && last > 0) {
      if (int rc = launch_sweep_fwd<float>(plan, p.sweeps[p.steps[0].index], batch, state, workspace, 0, (cudaStream_t)stream, true)) return rc;
      first = 1;
      zero_first = false;
    }
    if (int rc = qb_apply_forward_dev(plan, first, last, batch, state, workspace, 0, stream)) return rc;
    if (int rc = launch_sweep_fwd<float>(plan, p.sweeps[p.steps[last].index], batch, state, workspace, 0, (cudaStream_t)stream, zero_first, &cps)) return rc;
    probs_finalize_kernel<float><<<(unsigned)batch, 64, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const double*>(static_cast<const char*>(workspace) + layout(plan, batch).probs_part), cps, p.n_qubits, p.n_local,
        p.d_final_pos, 0, (float*)measure_out);
    QB_CUDA(cudaGetLastError());
    return 0;
  }
  if (zero_first) {
    if (int rc = launch_sweep_fwd<float>(plan, p.sweeps[p.steps[0].index], batch, state, workspace, 0, (cudaStream_t)stream, true)) return rc;
    first = 1;
  }
  if (int rc = qb_apply_forward_dev(plan, first, (int)p.steps.size(), batch, state, workspace, 0, stream)) return rc;
  if (measure == QB_MEASURE_PROBS) {
    QB_REQUIRE(measure_out, "measure_out is NULL");
    return qb_measure_probs_dev(plan, batch, state, measure_out, workspace, 0, stream);
  } else if (measure == QB_MEASURE_JOINT) {
    QB_REQUIRE(measure_out, "measure_out is NULL");
    return qb_measure_joint_dev(plan, batch, state, measure_out, stream);
  }
  QB_REQUIRE(measure == QB_MEASURE_STATE, "bad measure kind");
  return qb_convert_layout_dev(plan, batch, state, stream);  // the state IS the result: hand it back interleaved
}

int qb_backward_from_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                         int32_t n_batch_cols, const void* fixed_mats, const void* state_in, void* state, void* lambda,
                         int32_t measure, const void* grad_out, void* grad_shared, int32_t n_shared, void* grad_batch,
                         void* workspace, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(p.n_local == p.n_qubits, "qb_backward_dev is for unsharded plans; drive sharded plans step by step");
  QB_REQUIRE(state_in && state && lambda && grad_out, "state / lambda / grad_out is NULL");
  // the fused matrices must be in the workspace (they are if the same workspace was used by qb_forward_dev;
  // rebuilding them is cheap and makes the call self-contained)
  if (int rc = qb_prepare_dev(plan, batch, shared_angles, batch_angles, n_batch_cols, fixed_mats, workspace, stream)) return rc;
  int rc = 0;
  cudaStream_t st = (cudaStream_t)stream;
  const size_t state_bytes = ((size_t)batch << p.n_local) * (p.dtype == QB_C64 ? 8 : 16);
  int last = (int)p.steps.size();
  const bool fused_seed = measure == QB_MEASURE_PROBS && fuse_seed_ok(p);
  // out of place (state_in != state): the first adjoint sweep reads the forward's final state and writes the un-computed one when it
  // runs on a kernel that can (the seed-fused flat complex64 sweep: the hot path); otherwise one device-to-device copy comes first
  const bool oop = state_in != state;
  const bool oop_in_sweep = oop && fused_seed;
  if (oop && !oop_in_sweep) QB_CUDA(cudaMemcpyAsync(state, state_in, state_bytes, cudaMemcpyDeviceToDevice, st));
  const void* src = oop_in_sweep ? state_in : state;
  if (measure == QB_MEASURE_STATE) {  // qb_forward_dev returned the state interleaved
    if ((rc = qb_convert_layout_dev(plan, batch, state, stream))) return rc;
  }
  if (measure == QB_MEASURE_PROBS) {
    if (fused_seed) {
      QB_EMU_COUNT(g_emu_fused_seeds);
      if ((rc = qb_backward_begin_dev(plan, batch, workspace, stream))) return rc;
      --last;
      if ((rc = launch_sweep_bwd<float>(plan, p.sweeps[p.steps[last].index], batch, state, lambda, workspace, 0, st, grad_out,
                                        oop_in_sweep ? src : nullptr)))
        return rc;
      if ((rc = qb_apply_backward_dev(plan, 0, last, batch, state, lambda, workspace, 0, stream))) return rc;
      return qb_finalize_grads_dev(plan, batch, shared_angles, batch_angles, n_batch_cols, fixed_mats, workspace, grad_shared,
                                   n_shared, grad_batch, stream);
    }
    rc = qb_seed_probs_dev(plan, batch, state, grad_out, lambda, 0, stream);
  } else if (measure == QB_MEASURE_JOINT)
    rc = qb_seed_joint_dev(plan, batch, state, grad_out, lambda, stream);
  else if (measure == QB_MEASURE_STATE)
    rc = qb_seed_state_dev(plan, batch, grad_out, lambda, stream);
  else
    return fail("bad measure kind");
  if (rc) return rc;
  if ((rc = qb_backward_begin_dev(plan, batch, workspace, stream))) return rc;
  if ((rc = qb_apply_backward_dev(plan, 0, last, batch, state, lambda, workspace, 0, stream))) return rc;
  return qb_finalize_grads_dev(plan, batch, shared_angles, batch_angles, n_batch_cols, fixed_mats, workspace, grad_shared,
                               n_shared, grad_batch, stream);
}

int qb_backward_dev(const qb_plan* plan, int64_t batch, const void* shared_angles, const void* batch_angles,
                    int32_t n_batch_cols, const void* fixed_mats, void* state, void* lambda, int32_t measure,
                    const void* grad_out, void* grad_shared, int32_t n_shared, void* grad_batch, void* workspace,
                    void* stream) {
  return qb_backward_from_dev(plan, batch, shared_angles, batch_angles, n_batch_cols, fixed_mats, state, state, lambda, measure, grad_out,
                              grad_shared, n_shared, grad_batch, workspace, stream);
}

int32_t qb_plan_exchange_bits(const qb_plan* plan, int32_t step, int32_t* pos_out) {
  if (!plan || step < 0 || step >= (int32_t)plan->p.steps.size() || plan->p.steps[step].type != QB_STEP_EXCHANGE) return -1;
  const Exchange& ex = plan->p.exchanges[plan->p.steps[step].index];
  if (pos_out)
    for (int j = 0; j < ex.g; ++j) pos_out[j] = ex.pos[j];
  return ex.g;
}

int qb_exchange_p2p_dev(const qb_plan* plan, int64_t batch, const void* const* peer_state_ptrs, int32_t rank,
                        int32_t world, int32_t step, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(world >= 2 && world <= 16 && (world & (world - 1)) == 0, "world must be a power of two in [2, 16]");
  QB_REQUIRE(rank >= 0 && rank < world, "bad rank");
  QB_REQUIRE((1 << (p.n_qubits - p.n_local)) == world, "plan was not built for this world size");
  QB_REQUIRE(step >= 0 && step < (int)p.steps.size() && p.steps[step].type == QB_STEP_EXCHANGE, "not an exchange step");
  const Exchange& ex = p.exchanges[p.steps[step].index];
  const int amp_shift = p.dtype == QB_C64 ? 1 : 0;  // amplitudes per 16-byte vector: 2 (complex64) or 1
  ExchangeBits E{};
  E.g = ex.g;
  E.vbits = p.n_local - amp_shift;
  QB_REQUIRE(E.vbits >= E.g + 1, "shard too small for the peer exchange");
  uint32_t used = 0;
  for (int j = 0; j < ex.g; ++j) {
    QB_REQUIRE(ex.pos[j] >= amp_shift && ex.pos[j] < p.n_local, "exchange bit below the 16-byte vector");
    E.vpos[j] = ex.pos[j] - amp_shift;
    used |= 1u << E.vpos[j];
  }
  // the bit that splits a pair's work between its two ranks: low, but >= 256 B runs when the shard allows it
  int h = E.vbits >= 8 + E.g ? 4 : 0;
  while (used & (1u << h)) ++h;
  QB_REQUIRE(h < E.vbits, "no free bit to split the exchange work");
  E.hbit = h;
  int k = 0;
  for (int b = 0; b < E.vbits; ++b)
    if ((used | (1u << h)) & (1u << b)) E.ins[k++] = b;
  PeerPtrs pp{};
  for (int i = 0; i < world; ++i) {
    QB_REQUIRE(peer_state_ptrs[i] != nullptr, "NULL peer pointer");
    pp.p[i] = const_cast<void*>(peer_state_ptrs[i]);
  }
  const unsigned grid = (unsigned)plan->num_sms * 8;
  exchange_p2p_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(pp, E, rank, world, batch);
  QB_CUDA(cudaGetLastError());
  return 0;
}

#ifndef QB_KERNEL_EMU
int qb_exchange_push_dev(const qb_plan* plan, int64_t batch, void* state, const void* const* peer_staging_ptrs, int32_t rank,
                         int32_t world, int32_t piece, int32_t pieces, int32_t phase, void* stream) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(world >= 2 && world <= 16 && (world & (world - 1)) == 0, "world must be a power of two in [2, 16]");
  QB_REQUIRE(rank >= 0 && rank < world, "bad rank");
  QB_REQUIRE((1 << (p.n_qubits - p.n_local)) == world, "plan was not built for this world size");
  QB_REQUIRE(pieces >= 1 && piece >= 0 && piece < pieces, "bad piece");
  for (const Exchange& ex : p.exchanges)
    for (int j = 0; j < ex.g; ++j)
      QB_REQUIRE(ex.pos[j] == p.n_local - ex.g + j, "the push exchange handles top-bit exchanges only (plan option exchange_any_bit = 0)");
  const size_t sz = p.dtype == QB_C64 ? 8 : 16;
  const uint64_t chunk_bytes = ((uint64_t(1) << p.n_local) / world) * sz;
  QB_REQUIRE((pieces & (pieces - 1)) == 0 && chunk_bytes % ((uint64_t)pieces * 16) == 0,
             "pieces must be a power of two that splits the chunk into 16-byte aligned pieces");
  const uint64_t chunk_vec = chunk_bytes / 16, piece_vec = chunk_vec / pieces;
  cudaStream_t st = (cudaStream_t)stream;
  if (phase == 0) {
    ex::PushArgs A{};
    for (int i = 0; i < world; ++i) {
      QB_REQUIRE(peer_staging_ptrs[i] != nullptr, "NULL staging pointer");
      A.staging.p[i] = const_cast<void*>(peer_staging_ptrs[i]);
    }
    A.local = reinterpret_cast<const int4*>(state);
    A.rank = rank, A.world = world, A.batch = batch;
    A.chunk_vec = chunk_vec, A.piece_vec = piece_vec, A.piece_off = (uint64_t)piece * piece_vec;
    if (hooks().exchange_tma) {
      static int attr_dev = -1;
      if (attr_dev != plan->device) {
        QB_CUDA(cudaFuncSetAttribute(ex::exchange_push_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ex::kTmaSmem));
        attr_dev = plan->device;
      }
      ex::exchange_push_tma_kernel<<<(unsigned)plan->num_sms, ex::kTmaThreads, ex::kTmaSmem, st>>>(A);
    } else {
      ex::exchange_push_ldst_kernel<<<(unsigned)plan->num_sms * 8, 256, 0, st>>>(A);
    }
  } else {
    QB_REQUIRE(phase == 1, "phase must be 0 (push) or 1 (unpack)");
    ex::exchange_unpack_kernel<<<(unsigned)plan->num_sms * 8, 256, 0, st>>>(reinterpret_cast<const int4*>(peer_staging_ptrs[rank]),
                                                                           reinterpret_cast<int4*>(state), rank, world, batch, chunk_vec,
                                                                           piece_vec, (uint64_t)piece * piece_vec);
  }
  QB_CUDA(cudaGetLastError());
  return 0;
}
#endif

int qb_run_host(const qb_plan* plan, int64_t batch, const void* shared_angles, int32_t n_shared, const void* batch_angles,
                int32_t n_batch_cols, const void* fixed_mats, int32_t n_mats, const void* init_state, int32_t measure,
                void* measure_out, void* final_state_out, const void* grad_out, void* grad_shared, void* grad_batch,
                void* grad_init_state) {
  if (int rc = check_plan(plan, batch)) return rc;
  const Plan& p = plan->p;
  QB_REQUIRE(p.n_local == p.n_qubits, "qb_run_host is for unsharded plans");
  QB_REQUIRE(n_shared >= p.n_shared_slots && n_mats >= p.n_fixed_mats, "too few angles / matrices for the program");
  const size_t szT = p.dtype == QB_C64 ? 4 : 8;
  const size_t N = size_t(1) << p.n_qubits;
  const size_t state_bytes = (size_t)batch * N * 2 * szT;
  const size_t out_elems = measure == QB_MEASURE_PROBS ? (size_t)batch * p.n_qubits : measure == QB_MEASURE_JOINT ? (size_t)batch * N : 0;
  cudaStream_t st;
  QB_CUDA(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  void *d_sa = nullptr, *d_ba = nullptr, *d_fm = nullptr, *d_state = nullptr, *d_lam = nullptr, *d_out = nullptr, *d_g = nullptr,
       *d_gs = nullptr, *d_gb = nullptr, *d_ws = nullptr, *d_work = nullptr;
  int rc = 0;
  auto cleanup = [&]() {
    cudaFree(d_sa); cudaFree(d_ba); cudaFree(d_fm); cudaFree(d_state); cudaFree(d_lam); cudaFree(d_out); cudaFree(d_g);
    cudaFree(d_gs); cudaFree(d_gb); cudaFree(d_ws); cudaFree(d_work);
    cudaStreamDestroy(st);
  };
#define QB_H(call)                                                     \
  do {                                                                 \
    cudaError_t _e = (call);                                           \
    if (_e != cudaSuccess) {                                           \
      cleanup();                                                       \
      return fail(std::string(#call) + ": " + cudaGetErrorString(_e)); \
    }                                                                  \
  } while (0)
#define QB_R(call)       \
  do {                   \
    rc = (call);         \
    if (rc) {            \
      cleanup();         \
      return rc;         \
    }                    \
  } while (0)
  if (n_shared > 0) {
    QB_H(cudaMalloc(&d_sa, n_shared * szT));
    QB_H(cudaMemcpyAsync(d_sa, shared_angles, n_shared * szT, cudaMemcpyHostToDevice, st));
  }
  if (n_batch_cols > 0) {
    QB_H(cudaMalloc(&d_ba, (size_t)batch * n_batch_cols * szT));
    QB_H(cudaMemcpyAsync(d_ba, batch_angles, (size_t)batch * n_batch_cols * szT, cudaMemcpyHostToDevice, st));
  }
  if (n_mats > 0) {
    QB_H(cudaMalloc(&d_fm, (size_t)n_mats * 8 * szT));
    QB_H(cudaMemcpyAsync(d_fm, fixed_mats, (size_t)n_mats * 8 * szT, cudaMemcpyHostToDevice, st));
  }
  QB_H(cudaMalloc(&d_state, state_bytes));
  if (init_state) QB_H(cudaMemcpyAsync(d_state, init_state, state_bytes, cudaMemcpyHostToDevice, st));
  QB_H(cudaMalloc(&d_ws, (size_t)qb_workspace_bytes(plan, batch)));
  if (out_elems) QB_H(cudaMalloc(&d_out, out_elems * szT));
  QB_R(qb_forward_dev(plan, batch, d_sa, d_ba, n_batch_cols, d_fm, init_state ? QB_INIT_STATE : QB_INIT_ZERO, d_state, measure,
                      d_out, d_ws, st));
  if (measure_out && out_elems) QB_H(cudaMemcpyAsync(measure_out, d_out, out_elems * szT, cudaMemcpyDeviceToHost, st));
  if (final_state_out) {
    if (measure != QB_MEASURE_STATE) QB_R(qb_convert_layout_dev(plan, batch, d_state, st));  // internal -> interleaved
    QB_H(cudaMemcpyAsync(final_state_out, d_state, state_bytes, cudaMemcpyDeviceToHost, st));
    if (measure != QB_MEASURE_STATE && grad_out) QB_R(qb_convert_layout_dev(plan, batch, d_state, st));  // and back
  }
  if (grad_out) {
    const size_t g_bytes = measure == QB_MEASURE_STATE ? state_bytes : out_elems * szT;
    QB_H(cudaMalloc(&d_g, g_bytes));
    QB_H(cudaMemcpyAsync(d_g, grad_out, g_bytes, cudaMemcpyHostToDevice, st));
    QB_H(cudaMalloc(&d_lam, state_bytes));
    QB_H(cudaMalloc(&d_work, state_bytes));
    if (n_shared > 0) QB_H(cudaMalloc(&d_gs, n_shared * szT));
    if (n_batch_cols > 0) QB_H(cudaMalloc(&d_gb, (size_t)batch * n_batch_cols * szT));
    // out of place, like the torch.library layer: the final state is read, the un-computed copy goes to d_work
    QB_R(qb_backward_from_dev(plan, batch, d_sa, d_ba, n_batch_cols, d_fm, d_state, d_work, d_lam, measure, d_g, d_gs, n_shared, d_gb, d_ws, st));
    if (grad_shared && n_shared > 0) QB_H(cudaMemcpyAsync(grad_shared, d_gs, n_shared * szT, cudaMemcpyDeviceToHost, st));
    if (grad_batch && n_batch_cols > 0)
      QB_H(cudaMemcpyAsync(grad_batch, d_gb, (size_t)batch * n_batch_cols * szT, cudaMemcpyDeviceToHost, st));
    if (grad_init_state) {
      // torch convention: gradient w.r.t. the complex initial state = 2 * dL/dpsi0*, interleaved complex
      QB_R(qb_convert_layout_dev(plan, batch, d_lam, st));
      const uint64_t tot = (uint64_t)batch * N * 2;
      if (p.dtype == QB_C64)
        scale_kernel<float><<<ew_grid(tot, plan->num_sms), 256, 0, st>>>((const float*)d_lam, (float*)d_lam, 2.0f, tot);
      else
        scale_kernel<double><<<ew_grid(tot, plan->num_sms), 256, 0, st>>>((const double*)d_lam, (double*)d_lam, 2.0, tot);
      QB_H(cudaGetLastError());
      QB_H(cudaMemcpyAsync(grad_init_state, d_lam, state_bytes, cudaMemcpyDeviceToHost, st));
    }
  }
  QB_H(cudaStreamSynchronize(st));
  cleanup();
  return 0;
}

}  // extern "C"

#ifdef QB_PHASE_TIMES
// variant build only (flat64.cuh: QB_PHASE_TIMES): read (and optionally reset) the per-phase cycle counters of the flat complex64 sweeps
extern "C" int qb_debug_phase_times(unsigned long long* out8, int reset) {  // 12 counters
  if (cudaDeviceSynchronize() != cudaSuccess) return 1;
  if (out8 && cudaMemcpyFromSymbol(out8, qb::fl::g_phase_cycles, 12 * sizeof(unsigned long long)) != cudaSuccess) return 2;
  if (reset) {
    const unsigned long long z[12] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
    if (cudaMemcpyToSymbol(qb::fl::g_phase_cycles, z, sizeof(z)) != cudaSuccess) return 3;
  }
  return 0;
}
#endif

