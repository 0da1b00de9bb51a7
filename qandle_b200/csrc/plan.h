// Host-side plan of the state-vector engine: gate program -> fused 1-qubit groups -> shared-memory sweeps.
//
// Reference behaviour being replaced: the python loop over built gates (reference
// src/qandle/qcircuit.py:163-174), each applying a dense 2^n x 2^n matrix (operators.py:289-298).
// Here a *sweep* is one pass over the state in HBM: every CTA stages a tile of 2^m amplitudes (m index
// bits chosen by the planner, the lowest `low_bits` always among them so HBM chunks stay contiguous),
// applies every gate scheduled into the sweep, and writes the tile back.  Gates whose only link to an
// un-staged index bit is a control or a diagonal phase are still applied (the condition / phase is uniform
// over the tile).  Runs of 1-qubit gates on one qubit are fused into a single 2x2 (the reference's
// "gate-matrix cache" _a/_b, operators.py:235-238, becomes these device-built 2x2 blocks).
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace qb {

constexpr int kMaxTileBits = 14;
constexpr int kMaxQubits = 40;

// kernel-op kinds (device side)
enum KKind : int32_t {
  K_U1 = 1,       // general 2x2 on local bit a
  K_D1 = 2,       // diagonal 2x2 on local bit a
  K_D1_EXT = 3,   // diagonal 2x2 on an un-staged bit ext_bit: whole tile scaled by d[bit]
  K_CX = 4,       // X on local bit a, control local bit c
  K_CX_EXT = 5,   // X on local bit a if (global_base & ext_mask) == ext_mask
  K_CZ = 6,       // -1 where local bits a and c are set
  K_CZ_EXT1 = 7,  // -1 where local bit a is set, if ext condition
  K_CZ_EXT2 = 8,  // -1 on the whole tile, if ext condition
  K_SWAP = 9,     // swap local bits a and c
};

struct KOp {           // 32 bytes, mirrored on the device
  int16_t kind;
  int8_t r;            // staged kernels: register-bit index of local bit a within the stage (-1: not a register bit)
  int8_t rc;           // staged kernels: register-bit index of local bit c (-1: not a register bit)
  int32_t a;           // local target bit
  int32_t c;           // local control / second bit
  int32_t mat;         // (group mat index << 1) | is_batch, or -1
  uint64_t ext_mask;   // bits of the GLOBAL index that must all be 1
  int32_t ext_bit;     // K_D1_EXT: global bit selecting d0/d1
  int32_t kslot;       // index of this op's gradient accumulator within the sweep, or -1
};
static_assert(sizeof(KOp) == 32, "KOp layout");

enum MemberKind : int32_t { M_RX = 1, M_RY = 2, M_RZ = 3, M_U = 4 };

struct Member {
  int32_t kind;   // MemberKind
  int32_t slot;   // angle column / fixed-matrix index
  int32_t batch;  // 1: angle comes from batch_angles
  int32_t pad;
};

struct Group {
  int32_t qubit;         // logical qubit
  int32_t member_begin;  // into Plan::members
  int32_t member_count;
  int32_t batch;         // 1: any member per-sample -> matrix is per-sample
  int32_t diag;          // 1: all members diagonal (RZ only)
  int32_t has_param;     // 1: has at least one rotation member (needs a gradient accumulator)
  int32_t mat_index;     // index among shared (batch=0) or per-sample (batch=1) matrices
  int32_t k_index;       // index among shared / per-sample gradient accumulators, or -1
};

struct KSlot {           // per sweep: where a local accumulator goes
  int32_t batch;
  int32_t k_index;
};

// A stage of a register-blocked sweep: every thread holds the 2^RB amplitudes that differ in the stage's
// register bits (RB = 4 for complex64, 3 for complex128), applies ops [op_begin, op_end) in registers, and
// writes them back: one shared-memory round trip per stage instead of per gate.  `low` stages use the
// RB lowest local bits (contiguous, 128-bit accesses); other stages use arbitrary local bits.
struct Stage {
  int32_t low;
  int32_t regbits[4];  // ascending local bits
  int32_t op_begin, op_end;
  // packed complex64 kernel: ops [op_begin, pre_end) and [suf_begin, op_end) are CNOTs absorbed into the stage's
  // shared-memory load / store addressing (they cost no data movement); [pre_end, suf_begin) run in registers
  int32_t pre_end, suf_begin;
  // flat complex64 kernel (flat64.cuh): the stage's ops are in the canonical (execution) order
  //   [op_begin, pre_end)   CNOTs absorbed into the load addressing
  //   [pre_end, la_end)     CNOTs involving the pack lane (local bit 0), applied in registers right after the load
  //   [la_end, d_end)       sign flips (CZ family) and diagonal phases on thread / out-of-tile bits (K_D1 with r < 0,
  //                         K_D1_EXT): n_sign + n_phase ops
  //   [d_end, suf_begin)    at most one 2x2 (K_U1, or K_D1 on a register bit) per register bit: u_op[r] or -1
  //   [suf_begin, op_end)   CNOTs absorbed into the store addressing
  // The backward sweep has its own stage list in the same form over the reversed op order (Sweep::ops_bwd).
  // flat == 0: the stage is not in this form (other kernels ignore the fields below)
  int32_t flat;
  int32_t la_end, d_end;
  int32_t u_op[4];
  int32_t shape;  // bit r set <=> u_op[r] >= 0
  int32_t n_sign, n_phase;
  // bit 0: some absorbed CNOT targets a thread bit: amplitudes move between threads (extra barrier after the loads).
  // bits 4-5 / 8-9 (plan.cpp: sync_cost): how narrow the barrier after the stage's stores / that extra barrier may be:
  // 0 = __syncthreads(), 1 = 128-thread, 2 = 64-thread named barrier, 3 = __syncwarp()
  int32_t xthread;
};
static_assert(sizeof(Stage) == 80, "Stage layout");

struct Sweep {
  std::vector<int32_t> tile_bits;     // physical bits staged (size m_eff): local bit k <-> tile_bits[k]; the lowest L are 0..L-1
  std::vector<int32_t> nontile_bits;  // sorted physical local bits not staged
  std::vector<KOp> ops;               // in execution order (stage by stage when staged)
  std::vector<KSlot> kslots;
  std::vector<Stage> stages;          // empty: the sweep runs on the generic (one smem pass per op) kernel
  // flat plans: the backward sweep's own linearisation (execution order of the adjoint sweep: every op is applied as
  // its adjoint) and its stages; empty otherwise (the backward kernels then walk `ops` / `stages` in reverse)
  std::vector<KOp> ops_bwd;
  std::vector<Stage> stages_bwd;
  int32_t has_ext_diag_param = 0;     // some K_D1_EXT op carries a gradient: needs the tile inner product
  // device copies (owned by the plan)
  KOp* d_ops = nullptr;
  KSlot* d_kslots = nullptr;
  Stage* d_stages = nullptr;
  KOp* d_ops_bwd = nullptr;
  Stage* d_stages_bwd = nullptr;
};

struct Step {
  int32_t type;   // QB_STEP_SWEEP / QB_STEP_EXCHANGE
  int32_t index;  // index into Plan::sweeps / Plan::exchanges
};

// One exchange step of an amplitude-sharded plan: rank bit j (physical bit n_local + j) is swapped with local bit pos[j].
struct Exchange {
  int32_t g = 0;
  int32_t pos[4] = {-1, -1, -1, -1};
};

struct Plan {
  int32_t n_qubits = 0;
  int32_t n_local = 0;
  int32_t dtype = 0;
  int32_t tile_bits = 0, low_bits = 0;
  int32_t host_only = 0;
  int32_t packed = 0;  // complex64 sweeps use the packed (FFMA2, planar smem) kernel
  int32_t flat = 0;    // ... with flat stages (straight-line stage bodies, flat64.cuh)
  int32_t n_shared_slots = 0, n_batch_slots = 0, n_fixed_mats = 0;
  std::vector<Member> members;
  std::vector<Group> groups;
  int32_t n_groups_shared = 0, n_groups_batch = 0;  // matrix counts
  int32_t n_k_shared = 0, n_k_batch = 0;            // gradient accumulator counts
  std::vector<Sweep> sweeps;
  std::vector<Step> steps;
  std::vector<Exchange> exchanges;
  std::vector<int32_t> final_pos;  // logical qubit -> physical bit at the end
  int32_t max_kslots = 0;          // max accumulators in one sweep
  int32_t max_ops = 0;
  // device copies
  Member* d_members = nullptr;
  Group* d_groups = nullptr;
  int32_t* d_final_pos = nullptr;
};

struct GateIn {
  int32_t kind, q0, q1, slot, batch;
};

struct PlanOptions {
  int32_t tile_bits = 0, low_bits = 0, fuse = 1, n_local = 0, host_only = 0, swap_relabel = 1, final_layout = 0,
          max_ops_per_sweep = 0, staged = 1, packed = 1, flat = 1, narrow_sync = 1;
  // amplitude sharding: 0 = every exchange swaps the rank bits with the TOP local bits (a plain all-to-all over contiguous chunks:
  // what NCCL / the push exchange can do); 1 = the planner picks, per exchange, the local bits whose qubits are not needed for the
  // longest time (fewer exchanges; needs the peer-memory exchange kernel, which handles any bit positions)
  int32_t exchange_any_bit = 0;
  // single-GPU plans: search over where each sweep ends (plan.cpp: sweep-size search); 0 = plain greedy fill
  int32_t trim_search = 1;
};

// Does an adjoint sweep with this many ops / gradient slots fit the streaming adjoint kernel (three CTAs' shared memory in one SM)?
// Defined next to the launcher (capi.cu), used by the planner to keep sweeps on the fast kernel.
bool flat_stream_fits(int m, int L, int n_ops, int n_kslots);

// Throws std::runtime_error on invalid programs.
void build_plan(const std::vector<GateIn>& gates, int n_qubits, int dtype, const PlanOptions& opt, Plan& plan);

// Serialisation for tests (qb_plan_dump): int64 words
//   [0] magic 0x5142504c414e ("QBPLAN") [1] n_qubits [2] n_local [3] dtype [4] n_groups [5] n_members [6] n_steps
//   [7] n_sweeps [8] n_groups_shared [9] n_groups_batch [10] n_k_shared [11] n_k_batch
//   then groups (8 words each), members (3 words: kind, slot, batch), steps (2 words each),
//   final_pos (n_qubits words), then per sweep: m, n_ops, n_kslots, has_ext_diag_param, tile_bits[m],
//   ops (8 words each: kind,a,c,mat,ext_mask,ext_bit,kslot,(r+1)|((rc+1)<<8)), kslots (2 words each),
//   n_stages, stages (20 words each: low, regbits[4], op_begin, op_end, pre_end, suf_begin, flat, la_end, d_end,
//   u_op[4], shape, n_sign, n_phase, xthread), n_ops_bwd, ops_bwd (8 words each), n_stages_bwd, stages_bwd (20 words
//   each); after the sweeps: n_exchanges, then per exchange: g, pos[g].
void dump_plan(const Plan& plan, std::vector<int64_t>& out);

}  // namespace qb
