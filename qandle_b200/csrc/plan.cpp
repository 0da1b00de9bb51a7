// Planner: gate program -> fused groups -> sweeps (see plan.h).  Pure host C++, no CUDA.
#include "plan.h"

#include <algorithm>
#include <cstdlib>
#include <stdexcept>

#include "../../include/qandle_b200.h"

namespace qb {

namespace {

enum FKind : int32_t { F_U1 = 1, F_CNOT = 2, F_CZ = 3, F_SWAP = 4, F_LAYOUT_SWAP = 5 };

struct FOp {
  int32_t kind;
  int32_t qa, qb;  // logical qubits (F_U1: qa; F_CNOT: control qa, target qb)
  int32_t group;   // F_U1
};

struct Accepted {
  int32_t kind;    // FKind
  int32_t pa, pb;  // physical bits at acceptance time (pa: target / first, pb: control / second)
  int32_t group;
};

inline uint64_t bit(int q) { return uint64_t(1) << q; }

void fuse_pass(const std::vector<GateIn>& gates, int n, bool fuse, Plan& plan, std::vector<FOp>& fops) {
  std::vector<int> pending(n, -1);  // open group per qubit
  std::vector<std::vector<Member>> gmembers;
  auto flush = [&](int q) {
    if (pending[q] >= 0) {
      fops.push_back({F_U1, q, -1, pending[q]});
      pending[q] = -1;
    }
  };
  auto all_diag = [&](int g) {
    for (auto& m : gmembers[g])
      if (m.kind != M_RZ) return false;
    return true;
  };
  for (const auto& g : gates) {
    switch (g.kind) {
      case QB_OP_RX:
      case QB_OP_RY:
      case QB_OP_RZ:
      case QB_OP_U: {
        int q = g.q0;
        if (q < 0 || q >= n) throw std::runtime_error("gate qubit out of range");
        if (g.slot < 0) throw std::runtime_error("negative slot");
        if (pending[q] < 0 || !fuse) {
          flush(q);
          pending[q] = (int)gmembers.size();
          gmembers.emplace_back();
          Group grp{};
          grp.qubit = q;
          plan.groups.push_back(grp);
        }
        Member m{};
        m.kind = g.kind == QB_OP_RX ? M_RX : g.kind == QB_OP_RY ? M_RY : g.kind == QB_OP_RZ ? M_RZ : M_U;
        m.slot = g.slot;
        m.batch = (g.kind != QB_OP_U && g.batch) ? 1 : 0;
        gmembers[pending[q]].push_back(m);
        if (m.kind == M_U)
          plan.n_fixed_mats = std::max(plan.n_fixed_mats, g.slot + 1);
        else if (m.batch)
          plan.n_batch_slots = std::max(plan.n_batch_slots, g.slot + 1);
        else
          plan.n_shared_slots = std::max(plan.n_shared_slots, g.slot + 1);
        break;
      }
      case QB_OP_CNOT:
      case QB_OP_CZ:
      case QB_OP_SWAP: {
        int a = g.q0, b = g.q1;
        if (a < 0 || a >= n || b < 0 || b >= n || a == b) throw std::runtime_error("two-qubit gate: bad qubits");
        if (g.kind == QB_OP_SWAP) {
          flush(a);
          flush(b);
          fops.push_back({F_SWAP, a, b, -1});
        } else if (g.kind == QB_OP_CNOT) {
          // a diagonal run on the control commutes with the CNOT: keep it open
          if (pending[a] >= 0 && !all_diag(pending[a])) flush(a);
          flush(b);
          fops.push_back({F_CNOT, a, b, -1});
        } else {
          if (pending[a] >= 0 && !all_diag(pending[a])) flush(a);
          if (pending[b] >= 0 && !all_diag(pending[b])) flush(b);
          fops.push_back({F_CZ, a, b, -1});
        }
        break;
      }
      default:
        throw std::runtime_error("unknown opcode in gate program");
    }
  }
  for (int q = 0; q < n; ++q) flush(q);
  // finalise groups
  for (size_t gi = 0; gi < plan.groups.size(); ++gi) {
    Group& grp = plan.groups[gi];
    grp.member_begin = (int)plan.members.size();
    grp.member_count = (int)gmembers[gi].size();
    grp.batch = 0;
    grp.diag = 1;
    grp.has_param = 0;
    for (auto& m : gmembers[gi]) {
      plan.members.push_back(m);
      if (m.batch) grp.batch = 1;
      if (m.kind != M_RZ) grp.diag = 0;
      if (m.kind != M_U) grp.has_param = 1;
    }
    grp.mat_index = grp.batch ? plan.n_groups_batch++ : plan.n_groups_shared++;
    grp.k_index = -1;
    if (grp.has_param) grp.k_index = grp.batch ? plan.n_k_batch++ : plan.n_k_shared++;
  }
}


// Partition the ops of one sweep into register-blocked stages (plan.h: Stage).  Ops may be reordered when
// they touch disjoint index bits (they commute); the final op order is the stage order.
void schedule_stages(Sweep& sw, int RB, bool packed) {
  const int m = (int)sw.tile_bits.size();
  sw.stages.clear();
  if (m < RB) return;
  for (const KOp& k : sw.ops)
    if (k.kind == K_SWAP) return;  // physical swaps (layout restore only) stay on the generic kernel
  const int LOWB = RB;
  auto touch = [&](const KOp& k) {
    uint64_t t = k.ext_mask;
    if (k.a >= 0) t |= bit(sw.tile_bits[k.a]);
    if (k.c >= 0) t |= bit(sw.tile_bits[k.c]);
    if (k.ext_bit >= 0) t |= bit(k.ext_bit);
    return t;
  };
  auto need = [&](const KOp& k) { return (k.kind == K_U1 || k.kind == K_CX || k.kind == K_CX_EXT) ? k.a : -1; };
  std::vector<KOp> remaining = sw.ops, ordered;
  ordered.reserve(sw.ops.size());
  while (!remaining.empty()) {
    int first_need = -1;
    for (const KOp& k : remaining)
      if (need(k) >= 0) {
        first_need = need(k);
        break;
      }
    Stage st{};
    st.low = (first_need < 0 || first_need < LOWB) ? 1 : 0;
    uint32_t regset = st.low ? ((1u << RB) - 1u) : 0u;
    int nreg = st.low ? RB : 0;
    if (packed) {  // local bit 0 is the pack lane of every stage; the other 3 register bits are free
      st.low = 0;
      regset = 1u;
      nreg = 1;
    }
    uint64_t blocked = 0;
    std::vector<KOp> acc, next;
    for (const KOp& k : remaining) {
      const uint64_t t = touch(k);
      bool ok = !(t & blocked);
      if (ok) {
        const int nd = need(k);
        if (nd >= 0 && !((regset >> nd) & 1u)) {
          if (!st.low && (packed || nd >= LOWB) && nreg < RB) {
            regset |= 1u << nd;
            ++nreg;
          } else {
            ok = false;
          }
        }
      }
      if (ok) {
        acc.push_back(k);
      } else {
        blocked |= t;
        next.push_back(k);
      }
    }
    // complete the register-bit set: unused high bits first, then low ones
    for (int b = (packed ? 5 : LOWB); b < m && nreg < RB; ++b)
      if (!((regset >> b) & 1u)) {
        regset |= 1u << b;
        ++nreg;
      }
    for (int b = 0; b < m && nreg < RB; ++b)
      if (!((regset >> b) & 1u)) {
        regset |= 1u << b;
        ++nreg;
      }
    int ri = 0;
    int reg_of[32];
    for (int b = 0; b < 32; ++b) reg_of[b] = -1;
    for (int b = 0; b < m; ++b)
      if ((regset >> b) & 1u) {
        st.regbits[ri] = b;
        reg_of[b] = ri++;
      }
    for (; ri < 4; ++ri) st.regbits[ri] = -1;
    st.op_begin = (int)ordered.size();
    for (KOp k : acc) {
      k.r = (int8_t)(k.a >= 0 ? reg_of[k.a] : -1);
      k.rc = (int8_t)(k.c >= 0 ? reg_of[k.c] : -1);
      ordered.push_back(k);
    }
    st.op_end = (int)ordered.size();
    st.pre_end = st.op_begin;
    st.suf_begin = st.op_end;
    if (packed) {
      // CNOTs whose target and control are not the pack lane (local bit 0) are absorbed into the addressing when
      // they sit at the start / end of the stage
      auto absorbable = [&](const KOp& k) {
        if (k.kind == K_CX) return k.a != 0 && k.c != 0;
        if (k.kind == K_CX_EXT) return k.a != 0;
        return false;
      };
      while (st.pre_end < st.op_end && absorbable(ordered[st.pre_end])) ++st.pre_end;
      while (st.suf_begin > st.pre_end && absorbable(ordered[st.suf_begin - 1])) --st.suf_begin;
    }
    sw.stages.push_back(st);
    remaining.swap(next);
  }
  // Runs of consecutive sign flips (CZ family) inside a stage body commute: the staged kernels merge each run into one
  // sign mask.  First sink every sign flip as late as its neighbours allow (it commutes with ops on other bits and
  // with diagonal ops), so that flips cluster into longer runs; then mark the runs: ext_bit (unused by these ops)
  // carries the run length at the first and last op of a run, 0 in between.
  for (const Stage& st : sw.stages) {
    auto is_sign = [&](int i) { return ordered[i].kind == K_CZ || ordered[i].kind == K_CZ_EXT1 || ordered[i].kind == K_CZ_EXT2; };
    auto is_diag = [&](int i) { return ordered[i].kind == K_D1 || ordered[i].kind == K_D1_EXT; };
    for (int i = st.suf_begin - 2; i >= st.pre_end; --i) {
      if (!is_sign(i)) continue;
      int j = i;
      while (j + 1 < st.suf_begin && !is_sign(j + 1) && (is_diag(j + 1) || !(touch(ordered[j]) & touch(ordered[j + 1])))) {
        std::swap(ordered[j], ordered[j + 1]);
        ++j;
      }
    }
  }
  for (const Stage& st : sw.stages) {
    int k = st.pre_end;
    while (k < st.suf_begin) {
      auto is_sign = [&](int i) { return ordered[i].kind == K_CZ || ordered[i].kind == K_CZ_EXT1 || ordered[i].kind == K_CZ_EXT2; };
      if (!is_sign(k)) {
        ++k;
        continue;
      }
      int e = k;
      while (e + 1 < st.suf_begin && is_sign(e + 1)) ++e;
      for (int i = k; i <= e; ++i) ordered[i].ext_bit = 0;
      ordered[k].ext_bit = e - k + 1;
      ordered[e].ext_bit = e - k + 1;
      k = e + 1;
    }
  }
  sw.ops.swap(ordered);
}

// Flat stages for the complex64 flat kernel (plan.h: Stage::flat, flat64.cuh).  A stage is straight-line code:
// CNOTs first (absorbed into the load addressing, or -- when they involve the pack lane -- applied in registers),
// then one sign mask + per-thread phase, then at most one 2x2 per register bit, then CNOTs absorbed into the store
// addressing.  `in` must be a valid execution order; ops are reordered only across ops on disjoint index bits or
// among diagonal ops.  The result (ops + stages) is another valid execution order of the same operator product.
//
// `hi_mask` (narrow barriers, see assign_sync): local bits whose use as a register bit or as the target of an absorbed
// CNOT forces a wide barrier around the stage.  With `prefer_private` a stage is first built with those bits banned and
// kept when it carries as many 2x2s as the unrestricted stage would.
void schedule_flat(const std::vector<int32_t>& tile_bits, const std::vector<KOp>& in, std::vector<KOp>& ordered,
                   std::vector<Stage>& stages, bool packed, uint32_t hi_mask = 0, bool prefer_private = false) {
  // complex64 (packed): local bit 0 is the pack lane and a register bit of every stage, 3 more register bits, CNOTs
  // touching the lane run in registers.  complex128: 3 free register bits, every CNOT is absorbed.
  const int RB = packed ? 4 : 3;
  const int m = (int)tile_bits.size();
  ordered.clear();
  stages.clear();
  ordered.reserve(in.size());
  auto lbit = [&](int a) { return bit(tile_bits[a]); };
  auto touch = [&](const KOp& k) {
    uint64_t t = k.ext_mask;
    if (k.a >= 0) t |= lbit(k.a);
    if (k.c >= 0) t |= lbit(k.c);
    if (k.kind == K_D1_EXT) t |= bit(k.ext_bit);
    return t;
  };
  enum { R_PRE = 0, R_LANE, R_D, R_U, R_SUF, R_COUNT };
  std::vector<KOp> remaining = in;
  struct Built {
    uint32_t regset = 0, locked = 0;
    int nreg = 0;
    std::vector<KOp> part[R_COUNT], next;
  };
  // one greedy stage over `remaining`; `banned`: local bits that may neither become register bits nor be the target of
  // an absorbed CNOT
  auto build_stage = [&](uint32_t banned, Built& B) {
    uint32_t& regset = B.regset;
    uint32_t& locked = B.locked;
    int& nreg = B.nreg;
    regset = packed ? 1u : 0u;
    locked = 0;
    nreg = packed ? 1 : 0;
    uint64_t b_lane = 0, b_d = 0, b_u = 0, b_suf = 0, blocked = 0;
    std::vector<KOp>(&part)[R_COUNT] = B.part;
    std::vector<KOp>& next = B.next;
    auto can_add_reg = [&](int a) { return nreg < RB && !((locked >> a) & 1u) && !((banned >> a) & 1u); };
    for (const KOp& k : remaining) {
      const uint64_t t = touch(k);
      int region = -1, add_reg = -1, lock_bit = -1;
      if (!(t & blocked)) {
        switch (k.kind) {
          case K_CX:
          case K_CX_EXT: {
            const bool ctl_lane = packed && k.kind == K_CX && k.c == 0;
            if ((!packed || k.a != 0) && !ctl_lane) {
              if ((banned >> k.a) & 1u) break;
              region = (t & (b_lane | b_d | b_u | b_suf)) ? R_SUF : R_PRE;
            } else if (!(t & (b_d | b_u | b_suf))) {
              if (ctl_lane && !((regset >> k.a) & 1u)) {  // lanes are exchanged between two of the thread's packs
                if (!can_add_reg(k.a)) break;
                add_reg = k.a;
              }
              region = R_LANE;
            }
            break;
          }
          case K_CZ:
          case K_CZ_EXT1:
          case K_CZ_EXT2:
          case K_D1_EXT:
            if (!(t & (b_u | b_suf))) region = R_D;
            break;
          case K_U1:
          case K_D1: {
            if (t & (b_u | b_suf)) break;
            const bool is_reg = (regset >> k.a) & 1u;
            // a diagonal group on a thread bit is a per-thread phase; keep >= RB bits available as register bits
            if (k.kind == K_D1 && !is_reg && (m - __builtin_popcount(locked) - 1 >= RB)) {
              region = R_D;
              lock_bit = k.a;
            } else {
              if (!is_reg) {
                if (!can_add_reg(k.a)) break;
                add_reg = k.a;
              }
              region = R_U;
            }
            break;
          }
          default:
            throw std::runtime_error("flat stage scheduler: unexpected op kind");
        }
      }
      if (region < 0) {
        blocked |= t;
        next.push_back(k);
        continue;
      }
      if (add_reg >= 0) {
        regset |= 1u << add_reg;
        ++nreg;
      }
      if (lock_bit >= 0) locked |= 1u << lock_bit;
      switch (region) {
        case R_LANE: b_lane |= t; break;
        case R_D: b_d |= t; break;
        case R_U: b_u |= t; break;
        case R_SUF: b_suf |= t; break;
        default: break;
      }
      part[region].push_back(k);
    }
    // complete the register-bit set (never a locked thread bit): targets of absorbed CNOTs first (a CNOT whose target
    // is a register bit only permutes a thread's own amplitudes), then unused non-low bits (the ones outside hi_mask
    // first: a needless high register bit would widen the stage's barriers), then low ones
    auto add_free = [&](int b) {
      if (nreg < RB && b >= 0 && b < m && !((regset >> b) & 1u) && !((locked >> b) & 1u)) {
        regset |= 1u << b;
        ++nreg;
      }
    };
    for (int region : {R_PRE, R_SUF})
      for (const KOp& k : part[region]) add_free(k.a);
    for (int b = packed ? 5 : 4; b < m; ++b)
      if (!((hi_mask >> b) & 1u)) add_free(b);
    for (int b = 0; b < m; ++b)
      if (!((hi_mask >> b) & 1u)) add_free(b);
    for (int b = 0; b < m; ++b) add_free(b);
    if (nreg < RB) throw std::runtime_error("flat stage scheduler: register-bit set incomplete (internal error)");
  };
  while (!remaining.empty()) {
    Built B;
    build_stage(0, B);
    if (prefer_private && hi_mask) {
      Built P;
      build_stage(hi_mask, P);
      if (!P.part[R_U].empty() && P.part[R_U].size() >= B.part[R_U].size() && P.next.size() <= B.next.size()) B = std::move(P);
    }
    const uint32_t regset = B.regset;
    std::vector<KOp>(&part)[R_COUNT] = B.part;
    std::vector<KOp>& next = B.next;
    Stage st{};
    int ri = 0;
    int reg_of[32];
    for (int b = 0; b < 32; ++b) reg_of[b] = -1;
    for (int b = 0; b < m; ++b)
      if ((regset >> b) & 1u) {
        st.regbits[ri] = b;
        reg_of[b] = ri++;
      }
    for (; ri < 4; ++ri) st.regbits[ri] = -1;
    std::sort(part[R_U].begin(), part[R_U].end(), [&](const KOp& x, const KOp& y) { return reg_of[x.a] < reg_of[y.a]; });
    st.flat = 1;
    st.op_begin = (int)ordered.size();
    for (int r = 0; r < 4; ++r) st.u_op[r] = -1;
    for (int region = 0; region < R_COUNT; ++region) {
      for (KOp k : part[region]) {
        k.r = (int8_t)(k.a >= 0 ? reg_of[k.a] : -1);
        k.rc = (int8_t)(k.c >= 0 ? reg_of[k.c] : -1);
        if (region == R_U) {
          st.u_op[k.r] = (int)ordered.size();
          st.shape |= 1 << k.r;
        }
        if ((region == R_PRE || region == R_SUF) && k.r < 0) st.xthread = 1;
        if (region == R_D) {
          if (k.kind == K_D1 || k.kind == K_D1_EXT)
            ++st.n_phase;
          else
            ++st.n_sign;
        }
        ordered.push_back(k);
      }
      const int e = (int)ordered.size();
      switch (region) {
        case R_PRE: st.pre_end = e; break;
        case R_LANE: st.la_end = e; break;
        case R_D: st.d_end = e; break;
        case R_U: st.suf_begin = e; break;
        default: st.op_end = e; break;
      }
    }
    stages.push_back(st);
    remaining.swap(next);
  }
}

// ---- narrow barriers -------------------------------------------------------------------------------------------------
// In a flat stage thread g holds the amplitudes whose non-register local bits spell g (ascending), so an aligned group of
// 2^k threads owns the sub-cube of the local bits below RB + k -- as long as every register bit of the stage lies below
// RB + k.  The absorbed CNOT maps only flip their target bits, so if those lie below RB + k as well, the group loads
// and stores exactly its own sub-cube.  Two consecutive stages that both satisfy this for the same k exchange data only
// inside the 2^k-thread groups: the barrier between them can be a warp / named sub-CTA barrier instead of
// __syncthreads(), and warps of different groups drift apart (their shared-memory phases overlap the others' FP phases).
// narrow code = 8 - k: 0 = CTA barrier, 1 = 128 threads, 2 = 64 threads, 3 = one warp.  Only for tiles that fill a
// 256-thread CTA exactly (m - RB == 8).
int stage_top_bit(const std::vector<KOp>& ops, const Stage& st, const int* newpos) {
  int top = 0;
  for (int r = 0; r < 4; ++r)
    if (st.regbits[r] >= 0) top = std::max(top, newpos[st.regbits[r]]);
  for (int i = st.op_begin; i < st.pre_end; ++i) top = std::max(top, newpos[ops[i].a]);
  for (int i = st.suf_begin; i < st.op_end; ++i) top = std::max(top, newpos[ops[i].a]);
  return top;
}

constexpr double kSyncWeight[4] = {1.0, 0.6, 0.35, 0.1};  // relative cost of a barrier by narrow code

// weighted barrier cost of a stage list under a relabelling of the local bits; with `assign` the codes are written
// into Stage::xthread (bits 4-5: barrier after the stage's stores, bits 8-9: the cross-thread barrier after its loads)
double sync_cost(const std::vector<KOp>& ops, std::vector<Stage>& stages, int RB, const int* newpos, bool assign) {
  const int n = (int)stages.size();
  std::vector<int> k(n);
  for (int i = 0; i < n; ++i) k[i] = std::min(8, std::max(5, stage_top_bit(ops, stages[i], newpos) - RB + 1));
  double cost = 0;
  for (int i = 0; i < n; ++i) {
    const int k_end = i + 1 < n ? std::max(k[i], k[i + 1]) : 8;  // the tile store after the last stage needs the CTA
    const int x = stages[i].xthread & 1;
    cost += kSyncWeight[8 - k_end] + (x ? kSyncWeight[8 - k[i]] : 0.0);
    if (assign) stages[i].xthread = x | ((8 - k_end) << 4) | ((8 - k[i]) << 8);
  }
  return cost;
}

void schedule_flat_stages(Sweep& sw, bool packed, int L, int narrow) {
  sw.stages.clear();
  sw.ops_bwd.clear();
  sw.stages_bwd.clear();
  // tiles of more than 256 threads x 16 (complex64) / 8 (complex128) amplitudes would need several passes per stage,
  // which the in-place cross-thread CNOT absorption does not allow: those sweeps use the interpreted kernels
  const int m = (int)sw.tile_bits.size();
  const int RB = packed ? 4 : 3;
  if (m < RB || m > (packed ? 12 : 11)) return;
  for (const KOp& k : sw.ops)
    if (k.kind == K_SWAP) return;  // physical swaps (layout restore only) stay on the generic kernel
  const std::vector<KOp> ops_in = sw.ops;
  auto run = [&](const std::vector<KOp>& in, uint32_t hi_mask, bool prefer_private) {
    std::vector<KOp> fwd;
    schedule_flat(sw.tile_bits, in, fwd, sw.stages, packed, hi_mask, prefer_private);
    sw.ops.swap(fwd);
    std::vector<KOp> rev(sw.ops.rbegin(), sw.ops.rend());
    schedule_flat(sw.tile_bits, rev, sw.ops_bwd, sw.stages_bwd, packed, hi_mask, prefer_private);
  };
  run(ops_in, 0, false);
  if (sw.stages.size() > 32 || sw.stages_bwd.size() > 32 || sw.ops.size() > 8000) {  // flat64.cuh: kMaxFlatStages, 16-bit fields
    sw.ops = ops_in;
    sw.stages.clear();
    sw.ops_bwd.clear();
    sw.stages_bwd.clear();
    return;
  }
  if (!narrow || m - RB != 8 || m - L < 3) return;  // codes stay 0: CTA barriers
  // which staged bits sit in the three highest local positions is free (positions >= L only select HBM chunks): take
  // the ordered triple with the cheapest barriers for the schedule at hand (adjoint sweep weighted by its share of time)
  int ident[32];
  for (int b = 0; b < 32; ++b) ident[b] = b;
  auto total = [&](const int* np, bool assign) {
    return sync_cost(sw.ops, sw.stages, RB, np, assign) + 2.5 * sync_cost(sw.ops_bwd, sw.stages_bwd, RB, np, assign);
  };
  auto make_pos = [&](int b0, int b1, int b2, int* np) {  // b0 -> m-3, b1 -> m-2, b2 -> m-1, the others keep their order
    for (int b = 0; b < L; ++b) np[b] = b;
    int nx = L;
    for (int b = L; b < m; ++b)
      if (b != b0 && b != b1 && b != b2) np[b] = nx++;
    np[b0] = m - 3;
    np[b1] = m - 2;
    np[b2] = m - 1;
  };
  double best = total(ident, false);
  int bt[3] = {m - 3, m - 2, m - 1};
  for (int b0 = L; b0 < m; ++b0)
    for (int b1 = L; b1 < m; ++b1)
      for (int b2 = L; b2 < m; ++b2) {
        if (b0 == b1 || b0 == b2 || b1 == b2) continue;
        int np[32];
        make_pos(b0, b1, b2, np);
        const double c = total(np, false);
        if (c < best - 1e-9) {
          best = c;
          bt[0] = b0, bt[1] = b1, bt[2] = b2;
        }
      }
  if (bt[0] != m - 3 || bt[1] != m - 2 || bt[2] != m - 1) {
    int np[32];
    make_pos(bt[0], bt[1], bt[2], np);
    std::vector<int32_t> tb(m);
    for (int b = 0; b < m; ++b) tb[np[b]] = sw.tile_bits[b];
    std::vector<KOp> in = ops_in;
    for (KOp& k : in) {
      if (k.a >= 0) k.a = np[k.a];
      if (k.c >= 0) k.c = np[k.c];
    }
    const std::vector<int32_t> tb_old = sw.tile_bits;
    const std::vector<KOp> o_f = sw.ops, o_b = sw.ops_bwd;
    const std::vector<Stage> s_f = sw.stages, s_b = sw.stages_bwd;
    sw.tile_bits = tb;
    run(in, 0, false);
    if (sw.stages.size() > s_f.size() || sw.stages_bwd.size() > s_b.size()) {  // never trade barriers for stages
      sw.tile_bits = tb_old;
      sw.ops = o_f, sw.ops_bwd = o_b, sw.stages = s_f, sw.stages_bwd = s_b;
    }
  }
  // second try: stages built with the three high bits banned whenever that loses no 2x2 ("private" stages)
  {
    const uint32_t hi = 7u << (m - 3);
    const std::vector<KOp> o_f = sw.ops, o_b = sw.ops_bwd;
    const std::vector<Stage> s_f = sw.stages, s_b = sw.stages_bwd;
    const double c0 = total(ident, false);
    // the unscheduled op list in the current labelling: sw.ops is a valid execution order of it
    std::vector<KOp> in = sw.ops;
    for (KOp& k : in) k.r = k.rc = -1;
    run(in, hi, true);
    const bool more_stages = sw.stages.size() > s_f.size() || sw.stages_bwd.size() > s_b.size();
    if (more_stages || sw.stages.size() > 32 || sw.stages_bwd.size() > 32 || total(ident, false) >= c0 - 1e-9)
      sw.ops = o_f, sw.ops_bwd = o_b, sw.stages = s_f, sw.stages_bwd = s_b;
  }
  total(ident, true);
  if (narrow == 2)  // QB_NARROW_SYNC=u: every inner barrier a __syncwarp() -- results are WRONG; measures what barriers cost
    for (auto* sl : {&sw.stages, &sw.stages_bwd})
      for (size_t i = 0; i < sl->size(); ++i) (*sl)[i].xthread = ((*sl)[i].xthread & 1) | (i + 1 < sl->size() ? 3 << 4 : 0) | (3 << 8);
}

// Exchange lookahead (build_plan).  A sweep costs max(HBM pass, its arithmetic): one that holds fewer non-diagonal 2x2s than this is
// bound by the pass over the state alone (measured: a 2x2 adds ~1/5 of an HBM pass on both sweep kernels), so running it before
// an exchange buys nothing -- its gates still fit the fuller sweeps after the exchange.
constexpr int kSparseSweepOps = 4;

// sweep-size search (build_plan): cuts tried per sweep, sweeps planned ahead per cut, and the program size up to which it runs (the search
// costs kTrimMax x kTrimHorizon sweep constructions per sweep and candidate plan: 0.2 - 0.7 s of planning for configs 2 / 3, once per
// circuit, state size and device)
constexpr int kTrimMax = 10, kTrimHorizon = 4, kTrimMaxOps = 6000;
constexpr double kLow4SweepExtraMs = 0.06;  // a sweep over 128-byte instead of 256-byte HBM chunks (forward + adjoint, 2 GiB state)

// What a sweep costs beyond its gates' arithmetic, in ms on a 2 GiB state (forward + adjoint): least-squares fit over the 110 sweeps of
// config 2, 20 qubits and config 3 under greedy and searched plans (profiles/r2_stage_model.md 5; rms 0.06 / 0.13 ms per sweep):
//   forward  0.43 + 0.146 stages + 0.087 stages with in-register fix-ups (lane CNOTs, sign masks, phases)   [+ 0.044 per 2x2]
//   adjoint  0.79 + 0.283 stages + 0.245 stages with fix-ups                                              [+ 0.143 per 2x2]
// Only the ratios matter to the search.
double sweep_overhead_ms(const Sweep& sw) {
  auto fixups = [](const std::vector<Stage>& st) {
    int n = 0;
    for (const Stage& s : st) n += s.d_end > s.pre_end ? 1 : 0;
    return n;
  };
  return 1.22 + 0.146 * (double)sw.stages.size() + 0.087 * fixups(sw.stages) + 0.283 * (double)sw.stages_bwd.size() + 0.245 * fixups(sw.stages_bwd);
}

}  // namespace

namespace {
// strategy: 0 = plain greedy fill, 1 / 2 = sweep-size search with the lexicographic / the overhead-per-work score (see the search below)
void build_plan_one(const std::vector<GateIn>& gates, int n, int dtype, const PlanOptions& opt, int strategy, Plan& plan) {
  if (n < 1 || n > kMaxQubits) throw std::runtime_error("n_qubits must be in [1, 40]");
  if (dtype != QB_C64 && dtype != QB_C128) throw std::runtime_error("dtype must be QB_C64 or QB_C128");
  plan = Plan();
  plan.n_qubits = n;
  plan.dtype = dtype;
  plan.host_only = opt.host_only;
  plan.packed = (dtype == QB_C64 && opt.packed && opt.staged) ? 1 : 0;
  // flat stages: complex64 on the packed kernel, complex128 on its own flat kernel (also needs low_bits <= 8: below)
  plan.flat = (opt.staged && opt.flat && (plan.packed || dtype == QB_C128)) ? 1 : 0;
  plan.n_local = opt.n_local > 0 ? opt.n_local : n;
  if (plan.n_local > n) throw std::runtime_error("n_local > n_qubits");
  const int g_bits = n - plan.n_local;  // rank bits
  const bool sharded = g_bits > 0;
  int m = opt.tile_bits > 0 ? opt.tile_bits : (dtype == QB_C64 ? 12 : 11);
  int L = opt.low_bits > 0 ? opt.low_bits : (dtype == QB_C64 ? 5 : 4);
  m = std::min({m, (int)kMaxTileBits, plan.n_local});
  L = std::min(L, m);
  const int min_L = dtype == QB_C64 ? 1 : 0;  // 16-byte vectors
  if (L < min_L) throw std::runtime_error("state too small for 16-byte vector access");
  if (sharded && plan.n_local < 2 * g_bits) throw std::runtime_error("n_local must be >= 2*log2(world)");
  if (sharded && plan.n_local - g_bits < L) L = std::max(min_L, plan.n_local - g_bits);
  plan.tile_bits = m;
  plan.low_bits = L;
  const bool relabel = opt.swap_relabel || sharded;
  const int max_ops = opt.max_ops_per_sweep > 0 ? opt.max_ops_per_sweep : 1 << 30;

  std::vector<FOp> fops;
  fuse_pass(gates, n, opt.fuse != 0, plan, fops);

  std::vector<int> pos(n);
  for (int q = 0; q < n; ++q) pos[q] = n - 1 - q;
  const uint64_t all_q = n == 64 ? ~uint64_t(0) : (bit(n) - 1);

  // One greedy sweep: walk the remaining ops in program order, accept what fits the tile (at most m index bits, never a rank bit)
  // and is not blocked by an earlier rejected op on the same qubit.  Pure function of (remaining, pos): the scheduler also calls
  // it speculatively (exchange lookahead below).
  const bool stream_cap = plan.flat && dtype == QB_C64 && m == 12 && !sharded;
  struct Fill {
    std::vector<Accepted> acc;
    std::vector<FOp> next;
    std::vector<int> pos;
    uint64_t T = 0;
    int cnt = 0;
    bool rank_blocked = false;  // some op was rejected only because it needs a rank bit
  };
  auto fill_sweep = [&](const std::vector<FOp>& remaining, std::vector<int> pos, int limit = 1 << 30) {
    Fill out;
    std::vector<Accepted>& acc = out.acc;
    std::vector<FOp>& next = out.next;
    bool& rank_blocked = out.rank_blocked;
    uint64_t T = L > 0 ? (bit(L) - 1) : 0;
    int cnt = L;
    uint64_t blocked = 0;
    int n_kslots = 0;
    next.reserve(remaining.size());
    size_t i = 0;
    for (; i < remaining.size(); ++i) {
      const FOp& f = remaining[i];
      uint64_t qs = bit(f.qa) | (f.qb >= 0 ? bit(f.qb) : 0);
      // (a complex64 sweep must also fit the streaming adjoint kernel: past that size the adjoint sweep would fall back to the two-CTA
      // kernel -- measured on a 57-op sweep of config 2: 12.8 ms where two sweeps of that work take 11.5)
      bool full = (int)acc.size() >= std::min(max_ops, limit);
      if (!full && stream_cap && !(qs & blocked)) {
        const bool param = f.kind == F_U1 && plan.groups[f.group].has_param;
        if (!flat_stream_fits(m, L, (int)acc.size() + 1, n_kslots + (param ? 1 : 0))) full = true;
      }
      if ((qs & blocked) || full) {
        blocked |= qs;
        next.push_back(f);
        if (blocked == all_q) {
          ++i;
          break;
        }
        continue;
      }
      uint64_t need = 0;
      bool pure_relabel = false;
      switch (f.kind) {
        case F_U1:
          if (!plan.groups[f.group].diag) need = bit(pos[f.qa]);
          break;
        case F_CNOT:
          need = bit(pos[f.qb]);
          break;
        case F_CZ:
          break;
        case F_SWAP:
          if (relabel)
            pure_relabel = true;
          else
            need = bit(pos[f.qa]) | bit(pos[f.qb]);
          break;
        case F_LAYOUT_SWAP:
          need = bit(pos[f.qa]) | bit(pos[f.qb]);
          break;
      }
      if (pure_relabel) {
        std::swap(pos[f.qa], pos[f.qb]);
        continue;
      }
      bool fits = true;
      if (need >> plan.n_local) fits = false, rank_blocked = true;  // needs a rank bit: only an exchange can help
      uint64_t Tn = T | need;
      int cn = __builtin_popcountll(Tn);
      if (cn > m) fits = false;
      if (!fits) {
        blocked |= qs;
        next.push_back(f);
        if (blocked == all_q) {
          ++i;
          break;
        }
        continue;
      }
      T = Tn;
      cnt = cn;
      if (f.kind == F_U1 && plan.groups[f.group].has_param) ++n_kslots;
      Accepted a{};
      a.kind = f.kind;
      a.group = f.group;
      if (f.kind == F_U1) {
        a.pa = pos[f.qa];
        a.pb = -1;
      } else if (f.kind == F_CNOT) {
        a.pa = pos[f.qb];  // target
        a.pb = pos[f.qa];  // control
      } else {
        a.pa = pos[f.qa];
        a.pb = pos[f.qb];
      }
      if (f.kind == F_LAYOUT_SWAP) std::swap(pos[f.qa], pos[f.qb]);
      acc.push_back(a);
    }
    for (; i < remaining.size(); ++i) next.push_back(remaining[i]);
    out.T = T;
    out.cnt = cnt;
    out.pos = std::move(pos);
    return out;
  };
  // Which local bits an exchange swaps with the rank bits.  Default: the top g (a plain all-to-all over contiguous chunks).  With
  // exchange_any_bit: the g local bits (not below min_xpos: transfers stay >= 4 KB contiguous) whose qubits are needed -- as the
  // target of a non-diagonal gate -- furthest in the future (Belady); diagonal gates and CNOT / CZ controls run on rank bits.
  const int min_xpos = std::min(std::max(L, 8), plan.n_local - g_bits);
  auto choose_exchange = [&](const std::vector<FOp>& remaining, const std::vector<int>& p) {
    Exchange ex;
    ex.g = g_bits;
    for (int j = 0; j < g_bits; ++j) ex.pos[j] = plan.n_local - g_bits + j;
    if (!opt.exchange_any_bit) return ex;
    // first use per PHYSICAL bit of the layout `p`: walk the ops ahead, following the relabelling SWAPs among them
    std::vector<int> phys_of = p;  // logical qubit -> physical bit under the relabels seen so far
    std::vector<int64_t> use_phys(n, int64_t(1) << 60);
    for (size_t i = 0; i < remaining.size(); ++i) {
      const FOp& f = remaining[i];
      int need_q = -1, need_q2 = -1;
      switch (f.kind) {
        case F_U1:
          if (!plan.groups[f.group].diag) need_q = f.qa;
          break;
        case F_CNOT:
          need_q = f.qb;
          break;
        case F_SWAP:
          if (relabel)
            std::swap(phys_of[f.qa], phys_of[f.qb]);
          else
            need_q = f.qa, need_q2 = f.qb;
          break;
        case F_LAYOUT_SWAP:
          need_q = f.qa, need_q2 = f.qb;
          break;
        default:
          break;
      }
      for (int q : {need_q, need_q2})
        if (q >= 0 && use_phys[phys_of[q]] > (int64_t)i) use_phys[phys_of[q]] = (int64_t)i;
    }
    std::vector<int> cand;
    for (int b = min_xpos; b < plan.n_local; ++b) cand.push_back(b);
    if ((int)cand.size() < g_bits) return ex;
    std::stable_sort(cand.begin(), cand.end(), [&](int a, int b) {
      if (use_phys[a] != use_phys[b]) return use_phys[a] > use_phys[b];  // needed later first
      return a > b;                                                      // ties: the higher bit (longer contiguous runs)
    });
    std::vector<int> pick(cand.begin(), cand.begin() + g_bits);
    std::sort(pick.begin(), pick.end());
    for (int j = 0; j < g_bits; ++j) ex.pos[j] = pick[j];
    return ex;
  };
  // rank bit j <-> local bit ex.pos[j]
  auto exchanged = [&](std::vector<int> p, const Exchange& ex) {
    for (int q = 0; q < n; ++q) {
      if (p[q] >= plan.n_local)
        p[q] = ex.pos[p[q] - plan.n_local];
      else
        for (int j = 0; j < ex.g; ++j)
          if (p[q] == ex.pos[j]) {
            p[q] = plan.n_local + j;
            break;
          }
    }
    return p;
  };
  // One sweep from the ops a fill accepted: tile bits, kernel ops, gradient slots, stage schedule.  Pure (reads plan.groups only), so the
  // sweep-size search below can build sweeps speculatively.
  auto build_sweep = [&](const std::vector<Accepted>& acc, uint64_t T, int cnt) {
    // fill the tile up to m bits with the lowest free local bits
    for (int b = 0; b < plan.n_local && cnt < m; ++b)
      if (!(T & bit(b))) {
        T |= bit(b);
        ++cnt;
      }
    Sweep sw;
    std::vector<int> local_of(64, -1);
    for (int b = 0; b < plan.n_local; ++b) {
      if (T & bit(b)) {
        local_of[b] = (int)sw.tile_bits.size();
        sw.tile_bits.push_back(b);
      } else {
        sw.nontile_bits.push_back(b);
      }
    }
    for (const Accepted& a : acc) {
      KOp k{};
      k.mat = -1;
      k.kslot = -1;
      k.ext_bit = -1;
      k.c = -1;
      k.r = -1;
      k.rc = -1;
      switch (a.kind) {
        case F_U1: {
          const Group& grp = plan.groups[a.group];
          k.mat = (grp.mat_index << 1) | grp.batch;
          if (grp.has_param) {
            k.kslot = (int)sw.kslots.size();
            sw.kslots.push_back({grp.batch, grp.k_index});
          }
          if (local_of[a.pa] >= 0) {
            k.kind = (int16_t)(grp.diag ? K_D1 : K_U1);
            k.a = local_of[a.pa];
          } else {
            k.kind = K_D1_EXT;  // only diagonal groups are accepted without their bit staged
            k.ext_bit = a.pa;
            k.a = -1;
            if (grp.has_param) sw.has_ext_diag_param = 1;
          }
          break;
        }
        case F_CNOT:
          k.a = local_of[a.pa];
          if (local_of[a.pb] >= 0) {
            k.kind = K_CX;
            k.c = local_of[a.pb];
          } else {
            k.kind = K_CX_EXT;
            k.ext_mask = bit(a.pb);
          }
          break;
        case F_CZ: {
          int la = local_of[a.pa], lb = local_of[a.pb];
          if (la >= 0 && lb >= 0) {
            k.kind = K_CZ;
            k.a = la;
            k.c = lb;
          } else if (la >= 0 || lb >= 0) {
            k.kind = K_CZ_EXT1;
            k.a = la >= 0 ? la : lb;
            k.ext_mask = bit(la >= 0 ? a.pb : a.pa);
          } else {
            k.kind = K_CZ_EXT2;
            k.a = -1;
            k.ext_mask = bit(a.pa) | bit(a.pb);
          }
          break;
        }
        case F_SWAP:
        case F_LAYOUT_SWAP:
          k.kind = K_SWAP;
          k.a = local_of[a.pa];
          k.c = local_of[a.pb];
          break;
      }
      sw.ops.push_back(k);
    }
    if (plan.flat && L <= 8) schedule_flat_stages(sw, dtype == QB_C64, L, opt.narrow_sync);
    if (opt.staged && sw.stages.empty())  // not flat (or the flat form does not apply to this sweep)
      schedule_stages(sw, dtype == QB_C64 ? 4 : 3, dtype == QB_C64 && opt.packed);
    return sw;
  };
  bool last_was_exchange = false;

  std::vector<FOp> remaining = fops;
  bool layout_appended = false;
  while (true) {
    if (remaining.empty()) {
      if (layout_appended || opt.final_layout == 1) break;
      layout_appended = true;
      // restore the identity layout with physical swaps (only needed after relabelled SWAPs / exchanges)
      std::vector<int> p = pos;
      for (int q = 0; q < n; ++q) {
        int t = n - 1 - q;
        if (p[q] == t) continue;
        int q2 = -1;
        for (int r = 0; r < n; ++r)
          if (p[r] == t) q2 = r;
        remaining.push_back({F_LAYOUT_SWAP, q, q2, -1});
        std::swap(p[q], p[q2]);
      }
      if (remaining.empty()) break;
      if (sharded) throw std::runtime_error("final_layout=0 is not supported for amplitude-sharded plans with a permuted layout");
    }
    // ---- greedy: fill one sweep -----------------------------------------------------------------------
    Fill fl = fill_sweep(remaining, pos);
    if (sharded && !last_was_exchange && fl.rank_blocked) {
      // When to exchange.  Round 1 drained every op that could still run first (119 sweeps + 40 exchanges for config 4 on 2 GPUs: every
      // run ended in sparse sweeps along the blocked gates' light cones); the first fix exchanged as soon as a sweep filled after the
      // exchange would take more ops (88 + 40).  With planner-chosen exchange bits (choose_exchange) the drained layout is the good
      // one -- the qubits sent out are the ones not needed for longest, so the cones are wide -- and exchanging early only multiplies
      // the exchanges: config 4 on 2 / 4 / 8 GPUs 72 sweeps + 14 / 16 / 20 exchanges -> 69 + 5 / 70 + 6 / 71 + 6, config 5 18 + 4 -> 17 + 1.
      // So: drain, but skip a sweep that would be sparse (kSparseSweepOps) when the exchange lets a fuller one run.
      const Exchange ex = choose_exchange(remaining, pos);
      Fill fx = fill_sweep(remaining, exchanged(pos, ex));
      auto n_2x2 = [](const Fill& f) {
        int k = 0;
        for (const Accepted& a : f.acc) k += a.kind == F_U1 ? 1 : 0;
        return k;
      };
      // (Exchanges of the TOP local bits -- the NCCL / push modes -- keep the early rule: there the drained layout is not chosen.)
      const bool early = opt.exchange_any_bit ? (n_2x2(fl) < kSparseSweepOps && n_2x2(fx) > n_2x2(fl)) : fx.acc.size() > fl.acc.size();
      if (fl.acc.empty() || early) {
        pos = exchanged(pos, ex);
        plan.steps.push_back({QB_STEP_EXCHANGE, (int)plan.exchanges.size()});
        plan.exchanges.push_back(ex);
        last_was_exchange = true;
        continue;
      }
    }
    last_was_exchange = false;
    // ---- sweep-size search (single-GPU plans) ---------------------------------------------------------------------------------------
    // The greedy fill takes every op that fits, but which ops END a sweep decides how the qubits regroup in the following ones: leaving
    // the last 1 ... kTrimMax accepted ops (a suffix in program order -- still a valid cut) to the next sweep often saves sweeps and stages
    // further on (20 qubits SEL x 10: 20 -> 18 sweeps, measured -5.9 % per step; config 3: 181 / 180 -> 167 / 165 stages, -2.2 %).  A sweep
    // costs about one HBM pass whatever it holds, so sweeps weigh most: every cut is scored by planning kTrimHorizon sweeps ahead greedily
    // (two scores, see `score`); build_plan keeps whichever complete plan -- greedy, or searched with either score -- has the smallest
    // modelled overhead (sweep_overhead_ms).
    if (strategy != 0 && fl.acc.size() > 1) {
      // score of a cut (larger is better): plan kTrimHorizon sweeps ahead greedily
      auto score = [&](int first_limit) {
        std::vector<FOp> rem = remaining;
        std::vector<int> pp = pos;
        double cost = 0, work = 0;
        long done = 0;
        int sweeps = 0;
        bool finished = false;
        for (int h = 0; h < kTrimHorizon; ++h) {
          Fill f = fill_sweep(rem, pp, h == 0 ? first_limit : (1 << 30));
          if (f.acc.empty()) {
            finished = f.next.empty();
            break;
          }
          const Sweep sw = build_sweep(f.acc, f.T, f.cnt);
          cost += sweep_overhead_ms(sw);
          for (const Accepted& a : f.acc) work += (a.kind == F_U1 && !plan.groups[a.group].diag) ? 1.0 : 0.25;
          done += (long)f.acc.size();
          ++sweeps;
          pp = f.pos;
          rem.swap(f.next);
          if (rem.empty()) {
            finished = true;
            break;
          }
        }
        if (strategy == 1)  // sweeps first: finished plans by their sweep count, unfinished ones by the ops done; then the modelled overhead
          return std::make_tuple(finished ? 1 : 0, finished ? -sweeps : 0, done, -cost);
        return std::make_tuple(0, 0, 0L, -cost / std::max(work, 1.0));  // overhead per unit of gate work over the horizon
      };
      const int full = (int)fl.acc.size();
      auto best = score(full);
      int best_k = 0;
      for (int k = 1; k <= kTrimMax && k < full; ++k) {
        const auto sc = score(full - k);
        if (sc > best) best = sc, best_k = k;
      }
      if (best_k > 0) fl = fill_sweep(remaining, pos, full - best_k);
    }
    std::vector<Accepted>& acc = fl.acc;
    uint64_t T = fl.T;
    int cnt = fl.cnt;
    pos = fl.pos;
    remaining.swap(fl.next);

    if (acc.empty()) {
      if (remaining.empty()) continue;  // only relabels were left
      throw std::runtime_error("planner made no progress (internal error)");
    }
    Sweep sw = build_sweep(acc, T, cnt);
    plan.max_kslots = std::max(plan.max_kslots, (int)sw.kslots.size());
    plan.max_ops = std::max(plan.max_ops, (int)sw.ops.size());
    plan.steps.push_back({QB_STEP_SWEEP, (int)plan.sweeps.size()});
    plan.sweeps.push_back(std::move(sw));
  }
  plan.final_pos = pos;
}
}  // namespace

void build_plan(const std::vector<GateIn>& gates, int n, int dtype, const PlanOptions& opt, Plan& plan) {
  build_plan_one(gates, n, dtype, opt, 0, plan);
  if (!opt.trim_search || plan.sweeps.size() < 2 || gates.size() > (size_t)kTrimMaxOps) return;
  // Candidates: the greedy plan and the two searched ones; for complex64 states of up to 2^24 amplitudes per sample also with 128-byte HBM
  // chunks (low_bits 4: one more free tile bit, fewer sweeps, each ~5 % dearer -- measured: config 2 -4.3 %, config 3 -1.4 %, 20 qubits
  // +0.5 %; on 2^30 ... 2^33-amplitude states the scattered 128-byte chunks cost more than the saved sweeps, so not there).  The plan with
  // the smallest modelled overhead wins.
  // (amplitude-sharded plans: an exchange step counts like three empty sweeps -- measured 2.9 ms against ~1 ms for an empty
  // forward + adjoint sweep pair of the same shard on 8 GPUs, and psi and lambda both cross in the backward)
  auto overhead = [](const Plan& p, double per_sweep_extra) {
    double c = 3.0 * 1.22 * (double)p.exchanges.size();
    for (const Sweep& sw : p.sweeps) c += sweep_overhead_ms(sw) + per_sweep_extra;
    return c;
  };
  double best = overhead(plan, 0.0);
  const bool try_low4 = dtype == QB_C64 && opt.low_bits <= 0 && plan.low_bits == 5 && plan.tile_bits == 12 && n <= 24;
  for (int low : {0, 4}) {
    if (low == 4 && !try_low4) continue;
    PlanOptions o = opt;
    if (low) o.low_bits = low;
    for (int strategy : {0, 1, 2}) {
      if (low == 0 && strategy == 0) continue;  // the plan we hold
      Plan cand;
      build_plan_one(gates, n, dtype, o, strategy, cand);
      const double c = overhead(cand, low == 4 ? kLow4SweepExtraMs : 0.0);
      if (c < best * (1.0 - 1e-6)) {
        best = c;
        plan = std::move(cand);
      }
    }
  }
}

void dump_plan(const Plan& plan, std::vector<int64_t>& out) {
  out.clear();
  out.push_back(0x5142504c414eLL);
  out.push_back(plan.n_qubits);
  out.push_back(plan.n_local);
  out.push_back(plan.dtype);
  out.push_back((int64_t)plan.groups.size());
  out.push_back((int64_t)plan.members.size());
  out.push_back((int64_t)plan.steps.size());
  out.push_back((int64_t)plan.sweeps.size());
  out.push_back(plan.n_groups_shared);
  out.push_back(plan.n_groups_batch);
  out.push_back(plan.n_k_shared);
  out.push_back(plan.n_k_batch);
  for (const Group& g : plan.groups) {
    out.push_back(g.qubit);
    out.push_back(g.member_begin);
    out.push_back(g.member_count);
    out.push_back(g.batch);
    out.push_back(g.diag);
    out.push_back(g.has_param);
    out.push_back(g.mat_index);
    out.push_back(g.k_index);
  }
  for (const Member& m : plan.members) {
    out.push_back(m.kind);
    out.push_back(m.slot);
    out.push_back(m.batch);
  }
  for (const Step& s : plan.steps) {
    out.push_back(s.type);
    out.push_back(s.index);
  }
  for (int p : plan.final_pos) out.push_back(p);
  for (const Sweep& sw : plan.sweeps) {
    out.push_back((int64_t)sw.tile_bits.size());
    out.push_back((int64_t)sw.ops.size());
    out.push_back((int64_t)sw.kslots.size());
    out.push_back(sw.has_ext_diag_param);
    for (int b : sw.tile_bits) out.push_back(b);
    auto put_ops = [&](const std::vector<KOp>& ops) {
      for (const KOp& k : ops) {
        out.push_back(k.kind);
        out.push_back(k.a);
        out.push_back(k.c);
        out.push_back(k.mat);
        out.push_back((int64_t)k.ext_mask);
        out.push_back(k.ext_bit);
        out.push_back(k.kslot);
        out.push_back((int64_t)(k.r + 1) | ((int64_t)(k.rc + 1) << 8));
      }
    };
    auto put_stages = [&](const std::vector<Stage>& stages) {
      out.push_back((int64_t)stages.size());
      for (const Stage& st : stages) {
        out.push_back(st.low);
        for (int i = 0; i < 4; ++i) out.push_back(st.regbits[i]);
        out.push_back(st.op_begin);
        out.push_back(st.op_end);
        out.push_back(st.pre_end);
        out.push_back(st.suf_begin);
        out.push_back(st.flat);
        out.push_back(st.la_end);
        out.push_back(st.d_end);
        for (int i = 0; i < 4; ++i) out.push_back(st.u_op[i]);
        out.push_back(st.shape);
        out.push_back(st.n_sign);
        out.push_back(st.n_phase);
        out.push_back(st.xthread);
      }
    };
    put_ops(sw.ops);
    for (const KSlot& s : sw.kslots) {
      out.push_back(s.batch);
      out.push_back(s.k_index);
    }
    put_stages(sw.stages);
    out.push_back((int64_t)sw.ops_bwd.size());
    put_ops(sw.ops_bwd);
    put_stages(sw.stages_bwd);
  }
  out.push_back((int64_t)plan.exchanges.size());
  for (const Exchange& ex : plan.exchanges) {
    out.push_back(ex.g);
    for (int j = 0; j < ex.g; ++j) out.push_back(ex.pos[j]);
  }
}

}  // namespace qb
