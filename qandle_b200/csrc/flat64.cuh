// Flat complex64 sweep kernel (sm_100a): the packed kernel's data path (packed64.cuh: pack-planar 16-byte units,
// swizzled shared memory, FFMA2 2x2s, CNOTs folded into the stage addressing) with STRAIGHT-LINE stage bodies.
//
// Why: ncu's source view of the interpreted stage bodies (a loop over ops with a switch over handlers) showed ~36
// register-to-register MOVs per op at the loop head -- every handler leaves the 64 live amplitude registers renamed, and
// the loop-carried merge copies them back -- plus handler decode and per-op warp reductions: 3-4 non-FP instructions
// per FP instruction.  Here the planner (plan.cpp: schedule_flat) puts every stage into the canonical form
//     [CNOTs absorbed into the load address] [CNOTs on the pack lane: in-place XOR swaps]
//     [one sign mask + one per-thread phase] [at most one 2x2 per register bit] [CNOTs absorbed into the store address]
// so a stage is: load 8 (16) units, a few in-place fix-ups, ONE switch over the 16 "shapes" (which register bits carry a
// 2x2) whose cases are fully unrolled FFMA2 code ending in the shared-memory store -- the amplitude registers die inside
// the case, so there is no merge and no copy.  The adjoint sweep runs its own stage list (Sweep::ops_bwd), scheduled in
// its own execution order, through the same code on (psi, lambda) with adjoint matrices and the Pauli-vector
// accumulation in front of every 2x2 (4-value transposed warp reduction: 6 shuffles instead of 15).
#pragma once
#include "packed64.cuh"

namespace qb {
namespace fl {

using pk::NP;
#ifdef QB_PHASE_TIMES
// Variant build (build.py --variant phase -DQB_PHASE_TIMES): thread 0 of every CTA of the flat complex64 sweeps adds the clock64()
// cycles it spends per phase of the tile loop to g_phase_cycles (tools/sweep_times.py --phases); the default build has none of this.
//   [0] work-item setup  [1] tile: prefetch issue + per-tile tables  [2] wait for the tile (cp.async + barrier)  [3] stages
//   [4] tile epilogue (fused measurement) + shared -> HBM  [5] barrier at the end of the tile  [6] tiles  [7] CTA lifetime
//   [8] of [1]: tile offsets + cp.async issue (the rest of [1] = per-tile XOR tables)
__device__ unsigned long long g_phase_cycles[12];
__device__ __forceinline__ long long* ph_smem() {
  __shared__ long long a[14];
  return a;
}
#define QB_PH_DECL if (threadIdx.x == 0) { long long* ph = ph_smem(); for (int k = 0; k < 12; ++k) ph[k] = 0; ph[12] = ph[13] = clock64(); }
#define QB_PH(k) if (threadIdx.x == 0) { long long* ph = ph_smem(); const long long ph_n = clock64(); ph[k] += ph_n - ph[12]; ph[12] = ph_n; }
#define QB_PH_COUNT(k) if (threadIdx.x == 0) { ph_smem()[k] += 1; }
#define QB_PH_FLUSH if (threadIdx.x == 0) { long long* ph = ph_smem(); ph[7] = clock64() - ph[13]; for (int k = 0; k < 12; ++k) atomicAdd(&g_phase_cycles[k], (unsigned long long)ph[k]); }
#else
#define QB_PH_DECL
#define QB_PH(k)
#define QB_PH_COUNT(k)
#define QB_PH_FLUSH
#endif
constexpr int kMatF = 8;  // per op: the raw 2x2 (ar, ai, br, bi, cr, ci, dr, di); adjoint in the backward sweep

// 2x2 on pack-index bit RBIT: out0 = a x + b y, out1 = c x + d y (complex) on both lanes of the packs.  The matrix
// entries are SCALAR registers: FFMA2 / FMUL2 broadcast a 32-bit operand to both lanes (SASS: `FFMA2 R, -R9.F32, ...`),
// so a 2x2 needs 8 constant registers and two LDS.128 -- not 12 pre-broadcast pairs.
template <int RBIT>
__device__ __forceinline__ void u1_pack_s(float2 (&R)[NP], float2 (&I)[NP], const float* M) {
  const float4 m0 = reinterpret_cast<const float4*>(M)[0], m1 = reinterpret_cast<const float4*>(M)[1];
  const float2 ar = {m0.x, m0.x}, ai = {m0.y, m0.y}, br = {m0.z, m0.z}, bi = {m0.w, m0.w};
  const float2 cr = {m1.x, m1.x}, ci = {m1.y, m1.y}, dr = {m1.z, m1.z}, di = {m1.w, m1.w};
  const float2 nai = {-m0.y, -m0.y}, nbi = {-m0.w, -m0.w}, nci = {-m1.y, -m1.y}, ndi = {-m1.w, -m1.w};
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (j & (1 << RBIT)) continue;
    const int k = j | (1 << RBIT);
    const float2 xr = R[j], xi = I[j], yr = R[k], yi = I[k];
    R[j] = pk::f2fma(nbi, yi, pk::f2fma(br, yr, pk::f2fma(nai, xi, pk::f2mul(ar, xr))));
    I[j] = pk::f2fma(bi, yr, pk::f2fma(br, yi, pk::f2fma(ai, xr, pk::f2mul(ar, xi))));
    R[k] = pk::f2fma(ndi, yi, pk::f2fma(dr, yr, pk::f2fma(nci, xi, pk::f2mul(cr, xr))));
    I[k] = pk::f2fma(di, yr, pk::f2fma(dr, yi, pk::f2fma(ci, xr, pk::f2mul(cr, xi))));
  }
}

// bits (lane + 2 * pack) of the thread's 16 amplitudes whose register bit r is set
__device__ __forceinline__ uint32_t reg_pattern(int r) { return (uint32_t)(0xFF00F0F0CCCCAAAAull >> (16 * r)) & 0xFFFFu; }

// Pauli sums of one 2x2 -> per-warp accumulators.  Transposed butterfly: after the xor-16 and xor-8 steps every lane
// carries one of the 4 values (sx, sy, sz, pad), 3 more steps finish the sum; lanes 0 / 8 / 16 own sx / sy / sz.
__device__ __forceinline__ void warp_reduce3_accumulate(float s0, float s1, float s2, float* wacc_slot, bool write) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  const bool up16 = lane & 16, up8 = lane & 8;
  const float a0 = (up16 ? s2 : s0) + __shfl_xor_sync(full, up16 ? s0 : s2, 16);
  const float a1 = (up16 ? 0.f : s1) + __shfl_xor_sync(full, up16 ? s1 : 0.f, 16);
  float b = (up8 ? a1 : a0) + __shfl_xor_sync(full, up8 ? a0 : a1, 8);
  b += __shfl_xor_sync(full, b, 4);
  b += __shfl_xor_sync(full, b, 2);
  b += __shfl_xor_sync(full, b, 1);
  if (write && (lane & 7) == 0 && lane < 24) wacc_slot[lane >> 3] += b;
}

// XOR the sign bit of the amplitudes selected by the 16-bit mask M (bit = lane + 2 * pack)
__device__ __forceinline__ void apply_sign_mask(float2 (&R)[NP], float2 (&I)[NP], uint32_t M) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const uint32_t sx = (M << (31 - 2 * j)) & 0x80000000u, sy = (M << (30 - 2 * j)) & 0x80000000u;
    R[j].x = __uint_as_float(__float_as_uint(R[j].x) ^ sx);
    R[j].y = __uint_as_float(__float_as_uint(R[j].y) ^ sy);
    I[j].x = __uint_as_float(__float_as_uint(I[j].x) ^ sx);
    I[j].y = __uint_as_float(__float_as_uint(I[j].y) ^ sy);
  }
}

// 2x2 on the pack lane (local bit 0): x = lane .x, y = lane .y of every pack.  Packed form: the DATA is the broadcast
// scalar operand and the matrix columns are the pairs P1 = (ar, cr), P2 = (ai, ci), P3 = (br, dr), P4 = (bi, di):
//   R' = xr P1 - xi P2 + yr P3 - yi P4,   I' = xi P1 + xr P2 + yi P3 + yr P4      (8 FFMA2 per pack, results born as pairs)
// M: the op's constants in the order (ar, cr, ai, ci, br, dr, bi, di) (lane ops are stored transposed at CTA setup).
__device__ __forceinline__ void u1_lane_s(float2 (&R)[NP], float2 (&I)[NP], const float* M) {
  const float4 m0 = reinterpret_cast<const float4*>(M)[0], m1 = reinterpret_cast<const float4*>(M)[1];
  const float2 P1 = {m0.x, m0.y}, P2 = {m0.z, m0.w}, P3 = {m1.x, m1.y}, P4 = {m1.z, m1.w};
  const float2 N2 = {-m0.z, -m0.w}, N4 = {-m1.z, -m1.w};
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const float2 xr = {R[j].x, R[j].x}, xi = {I[j].x, I[j].x}, yr = {R[j].y, R[j].y}, yi = {I[j].y, I[j].y};
    R[j] = pk::f2fma(yi, N4, pk::f2fma(yr, P3, pk::f2fma(xi, N2, pk::f2mul(xr, P1))));
    I[j] = pk::f2fma(yr, P4, pk::f2fma(yi, P3, pk::f2fma(xr, P2, pk::f2mul(xi, P1))));
  }
}

// Per-stage descriptor built once per CTA (32 bytes: two LDS.128 per stage and thread)
struct alignas(16) SDesc {
  uint16_t la_begin, la_end, d_end;  // lane-CNOT ops [la_begin, la_end), sign / phase ops [la_end, d_end)
  uint8_t shape;                     // bit r: a 2x2 on register bit r
  uint8_t flags;                     // kXThread | kHasPhase | kNeedIb
  uint16_t u_mat[4];                 // float offset of the 2x2 of register bit r in smats
  int16_t u_kslot[4];                // backward: its gradient accumulator, or -1
  uint8_t regbits[4];
  // Absorbed CNOTs controlled by an OUT-OF-TILE bit contribute a per-tile XOR to the stage's load / store slots (extc).  Walking the
  // absorbed maps from x = 0, nothing happens before the first such op, so only [start, end) needs walking per tile:
  //   load side (maps applied in reverse): ops pre_end - 1 - ext_pre_skip down, ext_pre_n of them (0: no such op)
  //   store side: ops d_end + ext_suf_off up, ext_suf_n of them
  uint8_t ext_pre_skip, ext_pre_n, ext_suf_off, ext_suf_n;
};
static_assert(sizeof(SDesc) == 32, "SDesc layout");
constexpr int kXThread = 1, kHasPhase = 2, kNeedIb = 4;
constexpr int kExtWide = 128;  // flags: the ext ranges do not fit the u8 fields -> walk the whole ranges from the global stage table

// the ext-range fields of a stage descriptor (shared by the complex64 and complex128 flat kernels)
__device__ __forceinline__ void fill_ext_ranges(SDesc& d, const Stage& st, const KOp* ops) {
  int last_pre = -1, first_suf = -1;
  for (int i = st.op_begin; i < st.pre_end; ++i)
    if ((ops[i].kind & 0xFFFF) != K_CX) last_pre = i;
  for (int i = st.op_end - 1; i >= st.suf_begin; --i)
    if ((ops[i].kind & 0xFFFF) != K_CX) first_suf = i;
  const int pre_skip = last_pre >= 0 ? st.pre_end - 1 - last_pre : 0, pre_n = last_pre >= 0 ? last_pre - st.op_begin + 1 : 0;
  const int suf_off = first_suf >= 0 ? first_suf - st.d_end : 0, suf_n = first_suf >= 0 ? st.op_end - first_suf : 0;
  if (pre_skip > 255 || pre_n > 255 || suf_off > 255 || suf_n > 255) d.flags |= kExtWide;
  d.ext_pre_skip = (uint8_t)pre_skip, d.ext_pre_n = (uint8_t)pre_n, d.ext_suf_off = (uint8_t)suf_off, d.ext_suf_n = (uint8_t)suf_n;
}

// per-tile XOR constant of stage side `i & 1` of stage `i >> 1` (unit index, before the slot swizzle)
__device__ __forceinline__ uint32_t tile_ext_xor(const SDesc* sdesc, const Stage* stages, const KOp* sops, int i, uint64_t gbase) {
  const SDesc& d = sdesc[i >> 1];
  if (d.flags & kExtWide) {
    const Stage& st = stages[i >> 1];
    return (i & 1) ? pk::absorb_maps<false>(0u, sops, st.suf_begin, st.op_end, false, gbase, true)
                   : pk::absorb_maps<false>(0u, sops, st.op_begin, st.pre_end, true, gbase, true);
  }
  if (i & 1) {
    const int o0 = (int)d.d_end + d.ext_suf_off;
    return d.ext_suf_n ? pk::absorb_maps<false>(0u, sops, o0, o0 + d.ext_suf_n, false, gbase, true) : 0u;
  }
  const int o1 = (int)d.la_begin - d.ext_pre_skip;  // la_begin = pre_end
  return d.ext_pre_n ? pk::absorb_maps<false>(0u, sops, o1 - d.ext_pre_n, o1, true, gbase, true) : 0u;
}

// The per-tile XOR of a stage side in closed form.  The absorbed maps are GF(2)-linear in x and every out-of-tile-controlled CNOT adds
// its target bit iff its control bit is set in the tile's base, so  extc = XOR over those ops e of [bit_e(gbase)] * (the image of e's
// target bit under the maps that follow e)  -- and the slot swizzle is linear too.  With at most two such ops per side (the rule; `wide`
// otherwise: walk the maps, tile_ext_xor) a tile costs one 8-byte load and two bit tests instead of a chain of dependent shared-memory loads
// (thread 0's clock: 2.0 k -> 0.3 k cycles per tile under the other CTAs' shared-memory traffic).
struct alignas(8) ExtD {
  uint16_t val[2];  // slot offset >> 4 contributed when the control bit is set
  uint8_t bit[2];   // the control's bit of the global index
  uint8_t n, wide;
};
static_assert(sizeof(ExtD) == 8, "ExtD layout");

template <typename SlotFn>
__device__ __forceinline__ ExtD make_extd(const Stage& st, const KOp* ops, int side, SlotFn slot) {
  ExtD d{};
  const int o0 = side ? st.suf_begin : st.op_begin, o1 = side ? st.op_end : st.pre_end;
  for (int q = 0; q < o1 - o0; ++q) {
    const int i = side ? o0 + q : o1 - 1 - q;  // traversal order: store side forward, load side in reverse
    const KOp& op = ops[i];
    if ((op.kind & 0xFFFF) == K_CX) continue;
    if (d.n == 2 || __popcll(op.ext_mask) != 1) {
      d.wide = 1;
      break;
    }
    const uint32_t v = side ? pk::absorb_maps<false>(1u << op.a, ops, i + 1, o1, false, 0, false)
                            : pk::absorb_maps<false>(1u << op.a, ops, o0, i, true, 0, false);
    d.val[d.n] = (uint16_t)(slot(v) >> 4);
    d.bit[d.n] = (uint8_t)(__ffsll((long long)op.ext_mask) - 1);
    ++d.n;
  }
  return d;
}

// extc of stage side i for the tile at gbase, as a slot offset (before the forward kernels' buffer select)
template <typename SlotFn>
__device__ __forceinline__ uint32_t tile_extc(const ExtD* extd, const SDesc* sdesc, const Stage* stages, const KOp* sops, int i, uint64_t gbase,
                                              SlotFn slot) {
  const ExtD d = extd[i];
  if (d.wide) return slot(tile_ext_xor(sdesc, stages, sops, i, gbase));
  uint32_t x = 0;
  if (d.n > 0 && ((gbase >> d.bit[0]) & 1ull)) x ^= d.val[0];
  if (d.n > 1 && ((gbase >> d.bit[1]) & 1ull)) x ^= d.val[1];
  return x << 4;
}
constexpr int kEndNarrowShift = 3, kXNarrowShift = 5;  // flags bits 3-4 / 5-6: Stage::xthread bits 4-5 / 8-9 (narrow barriers)

// Barrier over the aligned group of (256 >> narrow) threads of a 256-thread CTA (plan.cpp: sync_cost).  Consecutive stages
// that keep their register bits and absorbed-CNOT targets below the group's top bit exchange amplitudes only inside the
// group, so the other warps of the CTA need not wait: 0 = __syncthreads(), 1 / 2 = named barriers of 128 / 64 threads
// (ids 1-2 / 3-6), 3 = __syncwarp().  `narrow` is CTA-uniform.
__device__ __forceinline__ void group_barrier(int narrow) {
  if (narrow == 0)
    __syncthreads();
  else if (narrow == 3)
    __syncwarp();
  else if (narrow == 1)
    asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x >> 7)) : "memory");
  else
    asm volatile("bar.sync %0, 64;" ::"r"(3 + (int)(threadIdx.x >> 6)) : "memory");
}
// A stage's load / store address table: 8 byte offsets (< 32 KB).  Forward kernels keep them as u16 pairs = ONE LDS.128 per use and an
// extra SHF per odd entry (two uses per stage: forward sweeps -2.5 ... -3.4 %); the adjoint kernels use a table five times per stage and
// keep plain u32 (two LDS.128 per use: with the packed form their extra ALU work cost +0.5 % -- both measured).
template <bool PACKED>
__device__ __forceinline__ void load_tab(const uint32_t* tab, uint32_t (&tw)[8]) {
  const uint4 t = *reinterpret_cast<const uint4*>(tab);
  if constexpr (PACKED) {
    tw[0] = t.x & 0xFFFFu, tw[1] = t.x >> 16, tw[2] = t.y & 0xFFFFu, tw[3] = t.y >> 16;
    tw[4] = t.z & 0xFFFFu, tw[5] = t.z >> 16, tw[6] = t.w & 0xFFFFu, tw[7] = t.w >> 16;
  } else {
    const uint4 u = reinterpret_cast<const uint4*>(tab)[1];
    tw[0] = t.x, tw[1] = t.y, tw[2] = t.z, tw[3] = t.w, tw[4] = u.x, tw[5] = u.y, tw[6] = u.z, tw[7] = u.w;
  }
}
// store entry j of table slot `slot` (= 2 * stage + side) in the layout load_tab<PACKED> reads
template <bool PACKED>
__device__ __forceinline__ void store_tab(uint32_t* stab, int slot, int j, uint32_t v) {
  if constexpr (PACKED)
    reinterpret_cast<uint16_t*>(stab)[slot * 16 + j] = (uint16_t)v;
  else
    stab[slot * 8 + j] = v;
}

// Fixed shared-memory offsets of the per-stage tables for a capacity of NS stages per sweep (immediate operands).
template <int NS>
struct FlatLay {
  static constexpr uint32_t kOffDesc = 0;                                  // SDesc [NS]
  static constexpr uint32_t kOffStab = kOffDesc + NS * 32;                 // [NS][2 (load, store)] x 32 bytes: NP byte offsets, u32 (adjoint) or
                                                                           // u16 in the first 16 bytes (forward): load_tab
  static constexpr uint32_t kOffExtc = kOffStab + NS * 2 * NP * 4;         // u32 [NS][2] per-tile XOR of out-of-tile controls
  static constexpr uint32_t kOffExtd = kOffExtc + NS * 2 * 4;              // ExtD [NS][2]: what extc is made of (per CTA)
  static constexpr uint32_t kOffTtab = kOffExtd + NS * 2 * 8;              // u16 [NS][2][32] thread-group nibble tables
  static constexpr uint32_t kOffHik = kOffTtab + NS * 2 * 32 * 2;          // u64 [32]: HBM byte offset of a tile's k-th 256-vector slab
  static constexpr uint32_t kOffBase = kOffHik + 32 * 8;                   // u64 [4]: ring of the CTA's next tile offsets
  static constexpr uint32_t kOffBuf = (kOffBase + 32 + 255) & ~255u;       // tile buffers
};
constexpr int kMaxFlatStages = 32;  // per sweep (plan.cpp falls back to the interpreted kernels beyond it)
constexpr int kStreamStages = 16;   // table capacity of the streaming adjoint kernel (3 CTAs / SM need the smaller tables)
constexpr uint32_t kOffDesc = FlatLay<kMaxFlatStages>::kOffDesc;
constexpr uint32_t kOffStab = FlatLay<kMaxFlatStages>::kOffStab;
constexpr uint32_t kOffExtc = FlatLay<kMaxFlatStages>::kOffExtc;
constexpr uint32_t kOffExtd = FlatLay<kMaxFlatStages>::kOffExtd;
constexpr uint32_t kOffTtab = FlatLay<kMaxFlatStages>::kOffTtab;
constexpr uint32_t kOffHik = FlatLay<kMaxFlatStages>::kOffHik;
constexpr uint32_t kOffBase = FlatLay<kMaxFlatStages>::kOffBase;
constexpr uint32_t kOffBuf = FlatLay<kMaxFlatStages>::kOffBuf;
constexpr uint32_t kFullBufBytes = 8u << 12;                            // one buffer of a full (2^12 amplitudes) tile

// Transposed butterfly reduction of P (4, 8 or 16) per-lane values over the warp: log2(P) exchange steps halve the number
// of values a lane carries, the remaining steps finish the sums.  Returns the total of value index (lane >> (5 - log2 P))
// (every lane of that group holds it): P - 1 + (5 - log2 P) shuffles instead of 5 P.
template <int P, typename T = float>
__device__ __forceinline__ T warp_transpose_reduce(T (&v)[P]) {
  const unsigned full = 0xffffffffu;
  const int lane = threadIdx.x & 31;
  int o = 16;
#pragma unroll
  for (int c = P; c > 1; c >>= 1, o >>= 1) {
    const bool up = lane & o;
#pragma unroll
    for (int i = 0; i < c / 2; ++i) {
      const T keep = up ? v[i + c / 2] : v[i];
      const T send = up ? v[i] : v[i + c / 2];
      v[i] = keep + __shfl_xor_sync(full, send, o);
    }
  }
  T t = v[0];
#pragma unroll
  for (; o > 0; o >>= 1) t += __shfl_xor_sync(full, t, o);
  return t;
}

// Pauli sums (sx, sy, sz) of the thread's amplitudes for a 2x2 on register bit RR, from the states after the group
template <int RR>
__device__ __forceinline__ void pauli_sums(const float2 (&R)[NP], const float2 (&I)[NP], const float2 (&LR)[NP],
                                           const float2 (&LI)[NP], float& sx, float& sy, float& sz) {
  if constexpr (RR == 0) {
    sx = sy = sz = 0.f;
    pk::pauli_lane(R, I, LR, LI, sx, sy, sz);
  } else {
    float2 px = {0, 0}, nx = {0, 0}, py = {0, 0}, ny = {0, 0}, pz = {0, 0}, nz = {0, 0};
    pk::pauli_pack<(RR > 0 ? RR - 1 : 0)>(R, I, LR, LI, px, nx, py, ny, pz, nz);
    sx = (px.x - nx.x) + (px.y - nx.y);
    sy = (py.x - ny.x) + (py.y - ny.y);
    sz = (pz.x - nz.x) + (pz.y - nz.y);
  }
}

template <int RR>
__device__ __forceinline__ void u_apply(float2 (&R)[NP], float2 (&I)[NP], const float* Mf) {
  if constexpr (RR == 0)
    u1_lane_s(R, I, Mf);
  else
    u1_pack_s<(RR > 0 ? RR - 1 : 0)>(R, I, Mf);
}

__host__ __device__ constexpr int popc4(int x) { return (x & 1) + ((x >> 1) & 1) + ((x >> 2) & 1) + ((x >> 3) & 1); }

// The 2x2s of one stage shape + the shared-memory store.  Everything is unrolled: no loop-carried amplitude registers.
// BWD: the 2x2s of a stage act on different bits and commute, so each of them may be taken as the last one applied: ALL
// Pauli sums come from the loaded (psi, lambda), in one batched warp reduction, and the adjoint 2x2s on psi and on
// lambda are then two independent instruction streams.
// FULL (2^12 tile on 256 threads): the tile buffers sit at CONSTANT shared-memory offsets (psi at kOffBuf -- the forward
// prefetch parity is folded into the per-tile XOR constants --, lambda 32 KB above), so every LDS / STS address is
// "slot register + immediate": no per-access IADD (ncu: 16 / 32 of them per forward / adjoint stage).
template <bool BWD, int SHAPE, bool FULL, int NS>
__device__ __forceinline__ void shape_body(float2 (&R)[NP], float2 (&I)[NP], float2 (&LR)[NP], float2 (&LI)[NP], const uint4 dw1,
                                           const float* smats, float* wacc, bool active, unsigned char* pbuf,
                                           unsigned char* lbuf, uint32_t sb, const uint32_t* tab_st) {
  constexpr uint32_t kOffBuf = FlatLay<NS>::kOffBuf;  // (shadows the 32-stage layout's constant)
  // dw1 = {u_mat[0..1], u_mat[2..3], u_kslot[0..1], u_kslot[2..3]}
  const float* M0 = smats + (dw1.x & 0xFFFFu);
  const float* M1 = smats + (dw1.x >> 16);
  const float* M2 = smats + (dw1.y & 0xFFFFu);
  const float* M3 = smats + (dw1.y >> 16);
  constexpr bool SUMS = BWD && SHAPE != 0;
  constexpr int NU = popc4(SHAPE);
  constexpr int P = NU <= 1 ? 4 : (NU == 2 ? 8 : 16);
  float v[P];
  float total = 0.f;
  int ks[4] = {-1, -1, -1, -1};  // kslot of the u-th 2x2 of the stage (ascending register bit)
  if constexpr (SUMS) {
    int u = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) v[i] = 0.f;
    if constexpr (SHAPE & 1) {
      pauli_sums<0>(R, I, LR, LI, v[4 * u], v[4 * u + 1], v[4 * u + 2]);
      ks[u++] = (int)(int16_t)(dw1.z & 0xFFFFu);
    }
    if constexpr (SHAPE & 2) {
      pauli_sums<1>(R, I, LR, LI, v[4 * u], v[4 * u + 1], v[4 * u + 2]);
      ks[u++] = (int)(int16_t)(dw1.z >> 16);
    }
    if constexpr (SHAPE & 4) {
      pauli_sums<2>(R, I, LR, LI, v[4 * u], v[4 * u + 1], v[4 * u + 2]);
      ks[u++] = (int)(int16_t)(dw1.w & 0xFFFFu);
    }
    if constexpr (SHAPE & 8) {
      pauli_sums<3>(R, I, LR, LI, v[4 * u], v[4 * u + 1], v[4 * u + 2]);
      ks[u++] = (int)(int16_t)(dw1.w >> 16);
    }
    if (!FULL && !active) {
#pragma unroll
      for (int i = 0; i < P; ++i) v[i] = 0.f;
    }
    total = warp_transpose_reduce<P>(v);
  }
  if constexpr (SUMS) {
    constexpr int SH = P == 4 ? 3 : (P == 8 ? 2 : 1);  // lanes per value = 1 << SH
    const int lane = threadIdx.x & 31, vi = lane >> SH, uu = vi >> 2, comp = vi & 3;
    const int kslot = uu == 0 ? ks[0] : (uu == 1 ? ks[1] : (uu == 2 ? ks[2] : ks[3]));
    if ((lane & ((1 << SH) - 1)) == 0 && comp < 3 && kslot >= 0) wacc[kslot * kAcc + comp] += total;
  }
  if constexpr (SHAPE & 1) u_apply<0>(R, I, M0);
  if constexpr (SHAPE & 2) u_apply<1>(R, I, M1);
  if constexpr (SHAPE & 4) u_apply<2>(R, I, M2);
  if constexpr (SHAPE & 8) u_apply<3>(R, I, M3);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t tw[NP];
  load_tab<!BWD>(tab_st, tw);
  // adjoint: psi is stored BEFORE lambda's 2x2s, so its 32 registers are free while lambda is processed (the stores
  // overlap that arithmetic, and the allocator has room to form the STS.128 quads without copies)
  if (FULL || active) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const uint32_t o = sb ^ tw[j];
      *reinterpret_cast<float4*>(FULL ? smem_raw + kOffBuf + o : pbuf + o) = float4{R[j].x, R[j].y, I[j].x, I[j].y};
    }
  }
  if constexpr (BWD) {
    if constexpr (SHAPE & 1) u_apply<0>(LR, LI, M0);
    if constexpr (SHAPE & 2) u_apply<1>(LR, LI, M1);
    if constexpr (SHAPE & 4) u_apply<2>(LR, LI, M2);
    if constexpr (SHAPE & 8) u_apply<3>(LR, LI, M3);
    if (FULL || active) {
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const uint32_t o = sb ^ tw[j];
        *reinterpret_cast<float4*>(FULL ? smem_raw + kOffBuf + kFullBufBytes + o : lbuf + o) = float4{LR[j].x, LR[j].y, LI[j].x, LI[j].y};
      }
    }
  }
}

// One in-register CNOT involving the pack lane (in-place XOR swaps; see pk::cx_static)
template <bool BWD>
__device__ __forceinline__ void lane_cx(float2 (&R)[NP], float2 (&I)[NP], float2 (&LR)[NP], float2 (&LI)[NP], const KOp& op,
                                        uint32_t ib, uint64_t gbase) {
#define QB_LCX(RT, RC)                      \
  {                                         \
    pk::cx_static<RT, RC>(R, I);            \
    if (BWD) pk::cx_static<RT, RC>(LR, LI); \
  }
  if (op.kind == K_CX && op.c == 0) {  // control = lane: the .y lanes of two packs are exchanged
    switch (op.r) {
      case 1: QB_LCX(1, 0) break;
      case 2: QB_LCX(2, 0) break;
      default: QB_LCX(3, 0) break;
    }
  } else {  // target = lane
    if (op.kind == K_CX && op.rc >= 1) {
      switch (op.rc) {
        case 1: QB_LCX(0, 1) break;
        case 2: QB_LCX(0, 2) break;
        default: QB_LCX(0, 3) break;
      }
    } else {
      const bool p = op.kind == K_CX ? ((ib >> op.c) & 1u) : ((gbase & op.ext_mask) == op.ext_mask);
      if (p) QB_LCX(0, 4)
    }
  }
#undef QB_LCX
}

// All stages of one tile (execution order; the adjoint sweep has its own list).  Deliberately NOT inlined: the tile loop's
// state (HBM addresses, prefetch bookkeeping) stays out of the stage loop's register budget.
template <bool BWD, bool FULL, int NS>
__device__ __noinline__ void run_stages(unsigned char* pbuf, unsigned char* lbuf, const int n_stages, const uint32_t n_groups,
                                        const uint64_t gbase, const float tdot, const float* smats, float* wacc, const KOp* sops) {
  using Lay = FlatLay<NS>;
  constexpr uint32_t kOffDesc = Lay::kOffDesc, kOffStab = Lay::kOffStab, kOffExtc = Lay::kOffExtc, kOffTtab = Lay::kOffTtab, kOffBuf = Lay::kOffBuf;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const uint16_t* ttab = reinterpret_cast<const uint16_t*>(smem_raw + kOffTtab);
  const bool warp_busy = FULL || (uint32_t)(tid & ~31) < n_groups;  // whole warps idle when the tile is small
  const bool active = FULL || (uint32_t)tid < n_groups;             // idle lanes of a partial warp shadow the last group
  const uint32_t my_g = active ? (uint32_t)tid : n_groups - 1;
  const uint32_t* tt_lo = reinterpret_cast<const uint32_t*>(ttab) + (my_g & 15);  // {load, store} base units as one 32-bit load
  const uint32_t* tt_hi = reinterpret_cast<const uint32_t*>(ttab) + 16 + (my_g >> 4);
  for (int si = 0; si < n_stages; ++si) {
    const unsigned char* sp = smem_raw + si * 32;  // every per-stage table is at a fixed offset + a multiple of si * 32
    const uint4 dw0 = *reinterpret_cast<const uint4*>(sp + kOffDesc);  // la_begin|la_end, d_end|shape|flags, u_mat[0..3]
    const uint2 ks = *reinterpret_cast<const uint2*>(sp + kOffDesc + 16);
    const uint4 dw1 = {dw0.z, dw0.w, ks.x, ks.y};
    const int shape = (dw0.y >> 16) & 0xFF, flags = dw0.y >> 24;
    const uint32_t* tab_ld = reinterpret_cast<const uint32_t*>(sp + si * 32 + kOffStab);
    const uint32_t* tab_st = tab_ld + NP;
    const uint2 ex = *reinterpret_cast<const uint2*>(smem_raw + kOffExtc + si * 8);
    const uint32_t ttx = tt_lo[si * 32] ^ tt_hi[si * 32];  // base unit of this thread's group: load side | store side << 16
    float2 R[NP], I[NP], LR[NP], LI[NP];
    // ---- load through the inverse of the absorbed prefix CNOTs --------------------------------------------------------
    if (warp_busy) {
      const uint32_t sbl = ((ttx & 0xFFFFu) << 4) ^ ex.x;
      uint32_t tw[NP];
      load_tab<!BWD>(tab_ld, tw);
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const uint32_t o = sbl ^ tw[j];
        const float4 pu = *reinterpret_cast<const float4*>(FULL ? smem_raw + kOffBuf + o : pbuf + o);
        R[j] = float2{pu.x, pu.y};
        I[j] = float2{pu.z, pu.w};
        if (BWD) {
          const float4 lu = *reinterpret_cast<const float4*>(FULL ? smem_raw + kOffBuf + kFullBufBytes + o : lbuf + o);
          LR[j] = float2{lu.x, lu.y};
          LI[j] = float2{lu.z, lu.w};
        }
      }
    }
    // absorbed CNOTs whose target is a thread bit move amplitudes between threads: every load of the stage must be done
    // before the first store
    if (flags & kXThread) group_barrier((flags >> kXNarrowShift) & 3);
    if (warp_busy) {
      // ---- rare in-place fix-ups: lane CNOTs, sign mask, per-thread phase (+ its gradients) ------------------------------
      if (flags & kNeedIb) {
        const int la_begin = dw0.x & 0xFFFF, la_end = dw0.x >> 16, d_end = dw0.y & 0xFFFF;
        const uint32_t rbw = *reinterpret_cast<const uint32_t*>(sp + kOffDesc + 24);  // regbits[0..3]
        uint32_t ib = my_g << 1;
        ib = ins0(ib, (rbw >> 8) & 0xFF);
        ib = ins0(ib, (rbw >> 16) & 0xFF);
        ib = ins0(ib, rbw >> 24);
        for (int i = la_begin; i < la_end; ++i) lane_cx<BWD>(R, I, LR, LI, sops[i], ib, gbase);
        uint32_t M = 0;
        float2 ph = {1.f, 0.f};
        float gsum = 0.f;
        if (BWD && (flags & kHasPhase)) {
          // sum over the thread's amplitudes of Im(conj(lam) psi): invariant under everything else in the stage
          gsum = pk::diag_grad_static<4>(R, I, LR, LI);
          if (!FULL && !active) gsum = 0.f;
        }
        for (int i = la_end; i < d_end; ++i) {
          const KOp& o = sops[i];
          const int kind = o.kind;
          if (kind == K_D1 || kind == K_D1_EXT) {
            const float* Mf = smats + (size_t)i * kMatF;
            const bool one = kind == K_D1 ? ((ib >> o.a) & 1u) : ((gbase >> o.ext_bit) & 1ull);
            const float2 d = one ? float2{Mf[6], Mf[7]} : float2{Mf[0], Mf[1]};
            ph = cmul(ph, d);
            if (BWD && o.kslot >= 0) {
              if (kind == K_D1)
                warp_accumulate1<float>(one ? -gsum : gsum, wacc + o.kslot * kAcc);
              else if (tid == 0)
                wacc[o.kslot * kAcc + 2] += one ? -tdot : tdot;
            }
          } else {
            uint32_t ok = 1u, ma = 0xFFFFu, mc = 0xFFFFu;
            if (kind != K_CZ) ok = ((gbase & o.ext_mask) == o.ext_mask) ? 1u : 0u;
            if (kind != K_CZ_EXT2) {
              if (o.r >= 0)
                ma = reg_pattern(o.r);
              else
                ok &= (ib >> o.a) & 1u;
            }
            if (kind == K_CZ) {
              if (o.rc >= 0)
                mc = reg_pattern(o.rc);
              else
                ok &= (ib >> o.c) & 1u;
            }
            M ^= ok ? (ma & mc) : 0u;
          }
        }
        if (M) {
          apply_sign_mask(R, I, M);
          if (BWD) apply_sign_mask(LR, LI, M);
        }
        if (flags & kHasPhase) {
          pk::diag_all(R, I, ph);
          if (BWD) pk::diag_all(LR, LI, ph);
        }
      }
      // ---- the stage's 2x2s + store through the absorbed suffix CNOTs: one fully unrolled case per shape ----------------
      const uint32_t sbs = ((ttx >> 16) << 4) ^ ex.y;
#define QB_SHAPE(S) \
case S: shape_body<BWD, S, FULL, NS>(R, I, LR, LI, dw1, smats, wacc, active, pbuf, lbuf, sbs, tab_st); break;
      switch (shape) {
        QB_SHAPE(0) QB_SHAPE(1) QB_SHAPE(2) QB_SHAPE(3) QB_SHAPE(4) QB_SHAPE(5) QB_SHAPE(6) QB_SHAPE(7)
        QB_SHAPE(8) QB_SHAPE(9) QB_SHAPE(10) QB_SHAPE(11) QB_SHAPE(12) QB_SHAPE(13) QB_SHAPE(14)
        default: shape_body<BWD, 15, FULL, NS>(R, I, LR, LI, dw1, smats, wacc, active, pbuf, lbuf, sbs, tab_st); break;
      }
#undef QB_SHAPE
    }
    group_barrier((flags >> kEndNarrowShift) & 3);
  }
}

// ---- streaming adjoint: the default adjoint kernel for full tiles -------------------------------------------------------
// The generic adjoint stage (run_stages<true>) holds psi AND lambda in registers (64) next to the Pauli accumulators: 128
// registers, 2 CTAs (16 warps) per SM.  Here lambda is never resident together with the accumulators: the Pauli sums are taken with psi in
// registers and lambda STREAMED from shared memory one 16-byte unit at a time; psi's adjoint 2x2s run and psi is stored;
// only then is lambda loaded in full, transformed and stored.  Cost: lambda is read twice (+8 LDS.128 per thread and
// stage), stages whose absorbed CNOTs move amplitudes between threads need a second barrier (all lambda re-loads before
// the first lambda store), stages with in-place fix-ups rewrite the fixed-up lambda first.  Gain: the stage fits 80
// registers -> 3 CTAs (24 warps) per SM and three independent barrier domains instead of two.
// Restrictions (the launcher falls back to the default kernel): full 2^12 tiles, at most kStreamStages stages, no
// parametrised diagonal in the sweep (their gradient needs psi and lambda together in the fix-up section).

// lane CNOTs, sign mask and fixed phase of a stage on ONE state (the default kernel does psi and lambda in one pass)
__device__ __forceinline__ void stage_fixups(float2 (&R)[NP], float2 (&I)[NP], const uint4 dw0, const uint32_t rbw, const uint32_t my_g,
                                             const uint64_t gbase, const float* smats, const KOp* sops, const int flags) {
  const int la_begin = dw0.x & 0xFFFF, la_end = dw0.x >> 16, d_end = dw0.y & 0xFFFF;
  uint32_t ib = my_g << 1;
  ib = ins0(ib, (rbw >> 8) & 0xFF);
  ib = ins0(ib, (rbw >> 16) & 0xFF);
  ib = ins0(ib, rbw >> 24);
  for (int i = la_begin; i < la_end; ++i) lane_cx<false>(R, I, R, I, sops[i], ib, gbase);
  uint32_t M = 0;
  float2 ph = {1.f, 0.f};
  for (int i = la_end; i < d_end; ++i) {
    const KOp& o = sops[i];
    const int kind = o.kind;
    if (kind == K_D1 || kind == K_D1_EXT) {
      const float* Mf = smats + (size_t)i * kMatF;
      const bool one = kind == K_D1 ? ((ib >> o.a) & 1u) : ((gbase >> o.ext_bit) & 1ull);
      ph = cmul(ph, one ? float2{Mf[6], Mf[7]} : float2{Mf[0], Mf[1]});
    } else {
      uint32_t ok = 1u, ma = 0xFFFFu, mc = 0xFFFFu;
      if (kind != K_CZ) ok = ((gbase & o.ext_mask) == o.ext_mask) ? 1u : 0u;
      if (kind != K_CZ_EXT2) {
        if (o.r >= 0)
          ma = reg_pattern(o.r);
        else
          ok &= (ib >> o.a) & 1u;
      }
      if (kind == K_CZ) {
        if (o.rc >= 0)
          mc = reg_pattern(o.rc);
        else
          ok &= (ib >> o.c) & 1u;
      }
      M ^= ok ? (ma & mc) : 0u;
    }
  }
  if (M) apply_sign_mask(R, I, M);
  if (flags & kHasPhase) pk::diag_all(R, I, ph);
}

// Pauli sums of the stage's 2x2s with psi in registers and lambda streamed from shared memory.  For the unit j of lambda
// and a 2x2 on pack bit b (partner k = j ^ (1 << b), s = +1 if bit b of j is clear, -1 if set), with c = conj(lam_j) psi_k:
//   sx += Im c,   sy -= s Re c,   sz += s Im(conj(lam_j) psi_j)          (pk::pauli_pack, one unit of lambda at a time)
template <int SHAPE, int P>
__device__ __forceinline__ void stream_pauli_sums(const float2 (&R)[NP], const float2 (&I)[NP], const unsigned char* lam_tile,
                                                  const uint32_t sbl, const uint32_t (&twl)[NP], float (&v)[P]) {
  float2 ax[3], ay[3], az[3];
#pragma unroll
  for (int b = 0; b < 3; ++b) ax[b] = ay[b] = az[b] = float2{0.f, 0.f};
  float lx[2] = {0.f, 0.f}, ly[2] = {0.f, 0.f}, lz[2] = {0.f, 0.f};  // lane 2x2: two accumulator sets (dependent-FMA latency)
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const float4 lu = *reinterpret_cast<const float4*>(lam_tile + (sbl ^ twl[j]));
    const float2 lr = {lu.x, lu.y}, li = {lu.z, lu.w};
    const float2 nlr = {-lu.x, -lu.y}, nli = {-lu.z, -lu.w};
#pragma unroll
    for (int b = 0; b < 3; ++b) {
      if (!(SHAPE & (2 << b))) continue;
      const int k = j ^ (1 << b);
      const bool hi = (j >> b) & 1;
      ax[b] = pk::f2fma(nli, R[k], pk::f2fma(lr, I[k], ax[b]));
      if (hi) {
        ay[b] = pk::f2fma(li, I[k], pk::f2fma(lr, R[k], ay[b]));
        az[b] = pk::f2fma(li, R[j], pk::f2fma(nlr, I[j], az[b]));
      } else {
        ay[b] = pk::f2fma(nli, I[k], pk::f2fma(nlr, R[k], ay[b]));
        az[b] = pk::f2fma(nli, R[j], pk::f2fma(lr, I[j], az[b]));
      }
    }
    if constexpr (SHAPE & 1)
      pauli_acc(lx[j & 1], ly[j & 1], lz[j & 1], float2{R[j].x, I[j].x}, float2{R[j].y, I[j].y}, float2{lr.x, li.x}, float2{lr.y, li.y});
  }
  int u = 0;
#pragma unroll
  for (int i = 0; i < P; ++i) v[i] = 0.f;
  if constexpr (SHAPE & 1) {
    v[0] = lx[0] + lx[1], v[1] = ly[0] + ly[1], v[2] = lz[0] + lz[1];
    u = 1;
  }
#pragma unroll
  for (int b = 0; b < 3; ++b) {
    if (!(SHAPE & (2 << b))) continue;
    v[4 * u] = ax[b].x + ax[b].y, v[4 * u + 1] = ay[b].x + ay[b].y, v[4 * u + 2] = az[b].x + az[b].y;
    ++u;
  }
}

template <int SHAPE, int NS>
__device__ __forceinline__ void shape_body_stream(float2 (&R)[NP], float2 (&I)[NP], const uint4 dw1, const float* smats, float* wacc,
                                                  const uint32_t sbl, const uint32_t* tab_ld, const uint32_t sbs, const uint32_t* tab_st,
                                                  const int flags) {
  using Lay = FlatLay<NS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  unsigned char* const psi_tile = smem_raw + Lay::kOffBuf;
  unsigned char* const lam_tile = smem_raw + Lay::kOffBuf + kFullBufBytes;
  const float* M0 = smats + (dw1.x & 0xFFFFu);
  const float* M1 = smats + (dw1.x >> 16);
  const float* M2 = smats + (dw1.y & 0xFFFFu);
  const float* M3 = smats + (dw1.y >> 16);
  constexpr int NU = popc4(SHAPE);
  constexpr int P = NU <= 1 ? 4 : (NU == 2 ? 8 : 16);
  constexpr bool SUMS = SHAPE != 0;
  float v[P];
  float total = 0.f;
  if constexpr (SUMS) {
    uint32_t twl[NP];
    load_tab<false>(tab_ld, twl);
    stream_pauli_sums<SHAPE, P>(R, I, lam_tile, sbl, twl, v);
  }
  // psi and lambda go through ONE copy of the stage's 2x2 code, executed twice (the registers are reused anyway) instead of two
  // unrolled copies: the kernel's code shrinks by about the size of the shape bodies' arithmetic (three CTAs at different places of
  // a 200 KB kernel thrash the instruction cache: ncu `no_instruction` 0.36 -> 1.49 stall cycles per instruction; measured +1 %).
  if constexpr (SUMS) {
    total = warp_transpose_reduce<P>(v);
    int ks[4] = {-1, -1, -1, -1};
    int u = 0;
    if constexpr (SHAPE & 1) ks[u++] = (int)(int16_t)(dw1.z & 0xFFFFu);
    if constexpr (SHAPE & 2) ks[u++] = (int)(int16_t)(dw1.z >> 16);
    if constexpr (SHAPE & 4) ks[u++] = (int)(int16_t)(dw1.w & 0xFFFFu);
    if constexpr (SHAPE & 8) ks[u++] = (int)(int16_t)(dw1.w >> 16);
    constexpr int SH = P == 4 ? 3 : (P == 8 ? 2 : 1);
    const int lane = threadIdx.x & 31, vi = lane >> SH, uu = vi >> 2, comp = vi & 3;
    const int kslot = uu == 0 ? ks[0] : (uu == 1 ? ks[1] : (uu == 2 ? ks[2] : ks[3]));
    if ((lane & ((1 << SH) - 1)) == 0 && comp < 3 && kslot >= 0) wacc[kslot * kAcc + comp] += total;
  }
#pragma unroll 1
  for (int pass = 0; pass < 2; ++pass) {
    unsigned char* const tile = pass ? lam_tile : psi_tile;
    if (pass) {  // lambda in full, into the registers psi has left
      uint32_t twl[NP];
      load_tab<false>(tab_ld, twl);
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float4 lu = *reinterpret_cast<const float4*>(lam_tile + (sbl ^ twl[j]));
        R[j] = float2{lu.x, lu.y};
        I[j] = float2{lu.z, lu.w};
      }
      if (flags & kXThread) group_barrier((flags >> kXNarrowShift) & 3);  // every lambda re-load before the first lambda store
    }
    if constexpr (SHAPE & 1) u_apply<0>(R, I, M0);
    if constexpr (SHAPE & 2) u_apply<1>(R, I, M1);
    if constexpr (SHAPE & 4) u_apply<2>(R, I, M2);
    if constexpr (SHAPE & 8) u_apply<3>(R, I, M3);
    uint32_t tws[NP];
    load_tab<false>(tab_st, tws);
#pragma unroll
    for (int j = 0; j < NP; ++j) *reinterpret_cast<float4*>(tile + (sbs ^ tws[j])) = float4{R[j].x, R[j].y, I[j].x, I[j].y};
  }
}

// Stage loop of the streaming adjoint sweep (full tiles: every thread owns one group; see run_stages for the shared parts)
template <int NS>
__device__ __noinline__ void run_stages_stream(const int n_stages, const uint64_t gbase, const float* smats, float* wacc, const KOp* sops) {
  using Lay = FlatLay<NS>;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t my_g = threadIdx.x;
  const uint16_t* ttab = reinterpret_cast<const uint16_t*>(smem_raw + Lay::kOffTtab);
  const uint32_t* tt_lo = reinterpret_cast<const uint32_t*>(ttab) + (my_g & 15);  // {load, store} base units as one 32-bit load
  const uint32_t* tt_hi = reinterpret_cast<const uint32_t*>(ttab) + 16 + (my_g >> 4);
  unsigned char* const psi_tile = smem_raw + Lay::kOffBuf;
  unsigned char* const lam_tile = smem_raw + Lay::kOffBuf + kFullBufBytes;
  for (int si = 0; si < n_stages; ++si) {
    const unsigned char* sp = smem_raw + si * 32;
    const uint4 dw0 = *reinterpret_cast<const uint4*>(sp + Lay::kOffDesc);
    const uint2 ks = *reinterpret_cast<const uint2*>(sp + Lay::kOffDesc + 16);
    const uint4 dw1 = {dw0.z, dw0.w, ks.x, ks.y};
    const int shape = (dw0.y >> 16) & 0xFF, flags = dw0.y >> 24;
    const uint32_t* tab_ld = reinterpret_cast<const uint32_t*>(sp + si * 32 + Lay::kOffStab);
    const uint32_t* tab_st = tab_ld + NP;
    const uint2 ex = *reinterpret_cast<const uint2*>(smem_raw + Lay::kOffExtc + si * 8);
    const uint32_t ttx = tt_lo[si * 32] ^ tt_hi[si * 32];  // base unit of this thread's group: load side | store side << 16
    const uint32_t sbl = ((ttx & 0xFFFFu) << 4) ^ ex.x;
    const uint32_t rbw = *reinterpret_cast<const uint32_t*>(sp + Lay::kOffDesc + 24);  // regbits[0..3]
    float2 R[NP], I[NP];
    {
      uint32_t tw[NP];
      load_tab<false>(tab_ld, tw);
      if (flags & kNeedIb) {
        // lambda's in-place fix-ups first, written back to where they were loaded (slots private to this thread)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float4 lu = *reinterpret_cast<const float4*>(lam_tile + (sbl ^ tw[j]));
          R[j] = float2{lu.x, lu.y};
          I[j] = float2{lu.z, lu.w};
        }
        stage_fixups(R, I, dw0, rbw, my_g, gbase, smats, sops, flags);
#pragma unroll
        for (int j = 0; j < NP; ++j) *reinterpret_cast<float4*>(lam_tile + (sbl ^ tw[j])) = float4{R[j].x, R[j].y, I[j].x, I[j].y};
      }
#pragma unroll
      for (int j = 0; j < NP; ++j) {
        const float4 pu = *reinterpret_cast<const float4*>(psi_tile + (sbl ^ tw[j]));
        R[j] = float2{pu.x, pu.y};
        I[j] = float2{pu.z, pu.w};
      }
    }
    // every psi load of the stage before the first psi store (absorbed CNOTs with a thread-bit target)
    if (flags & kXThread) group_barrier((flags >> kXNarrowShift) & 3);
    if (flags & kNeedIb) stage_fixups(R, I, dw0, rbw, my_g, gbase, smats, sops, flags);
    const uint32_t sbs = ((ttx >> 16) << 4) ^ ex.y;
#define QB_SSHAPE(S) \
case S: shape_body_stream<S, NS>(R, I, dw1, smats, wacc, sbl, tab_ld, sbs, tab_st, flags); break;
    switch (shape) {
      QB_SSHAPE(0) QB_SSHAPE(1) QB_SSHAPE(2) QB_SSHAPE(3) QB_SSHAPE(4) QB_SSHAPE(5) QB_SSHAPE(6) QB_SSHAPE(7)
      QB_SSHAPE(8) QB_SSHAPE(9) QB_SSHAPE(10) QB_SSHAPE(11) QB_SSHAPE(12) QB_SSHAPE(13) QB_SSHAPE(14)
      default: shape_body_stream<15, NS>(R, I, dw1, smats, wacc, sbl, tab_ld, sbs, tab_st, flags); break;
    }
#undef QB_SSHAPE
    group_barrier((flags >> kEndNarrowShift) & 3);
  }
}

// shared-memory layout (dynamic):
//   [fixed-offset stage tables: kOffDesc .. kOffBuf][tile buffer 0][tile buffer 1][smats: n_ops x 8 f32]
//   [bwd: wacc (warps x kslots x kAcc) + wred (warps)][hi_off: 2^(m-L) u32][ops: n_ops KOp]
__host__ __device__ inline size_t flat_smem_bytes(int m, int L, int n_ops, int n_kslots, int n_stages, bool backward,
                                                  bool stream_tables = false) {
  auto al = [](size_t x) { return (x + 15) & ~size_t(15); };
  // forward: two psi buffers (double-buffered prefetch); backward: psi + lambda
  size_t b = (stream_tables ? FlatLay<kStreamStages>::kOffBuf : kOffBuf) + (size_t(1) << m) * 8 * 2;
  b += size_t(n_ops) * kMatF * 4;
  if (backward) b += size_t(kMaxWarps) * n_kslots * kAcc * 4 + size_t(kMaxWarps) * 4;
  b = al(b);
  b = al(b + (size_t(1) << (m - L)) * 4);
  b = al(b + size_t(n_ops) * sizeof(KOp));
  return b;
}

// threads per CTA: one per 16 amplitudes of the tile, at most 256 (flat tiles are at most 2^12 amplitudes); at least 64 and
// at least one HBM chunk of vectors, which the tile <-> HBM address split (see the kernel) relies on
__host__ __device__ inline int flat_threads(int m, int L) {
  int t = 1 << (m > 4 ? m - 4 : 0);
  if (t > kSweepThreads) t = kSweepThreads;
  if (t < 64) t = 64;
  if (t < (1 << (L > 0 ? L - 1 : 0))) t = 1 << (L - 1);
  return t;
}

// static shared memory of the fused probability reduction, behind functions so that only the kernels that use it carry it
__device__ __forceinline__ double* pr_acc_smem() {
  __shared__ double a[49];
  return a;
}
__device__ __forceinline__ float* pr_wred_smem() {
  __shared__ float a[8 * 13];
  return a;
}

// The flat complex64 sweep kernel.
// BWD:    adjoint sweep (psi <- G^+ psi, lambda <- G^+ lambda, Pauli-vector gradient sums); otherwise the forward sweep with a
//         double-buffered tile prefetch.
// FULL:   the tile has 2^12 amplitudes and the CTA 256 threads (every thread owns one group, constant buffer offsets).
// STREAM: (adjoint, full tiles) run_stages_stream -- 80 registers, 3 CTAs / SM, stage tables for kStreamStages stages.
// DYN:    persistent CTAs.  The grid is one CTA per resident slot; the sample-independent setup (stage descriptors, address tables)
//         is done once per CTA and the B * cps work items -- (sample, tile-subset) pairs, the CTAs of a static launch -- are claimed
//         through an atomic counter: no partial last wave (config 2: 4096 CTAs on 444 slots = 9.2 waves) and the table setup is paid
//         444 times instead of 4096 (measured +3.5 % config 2, +4.7 % 20 qubits against the static launch).
// FUSE:   the measurement next to the circuit runs inside the sweep -- forward: MeasureProbability's reduction on every finished tile
//         (PA.probs_part; replaces probs_partial_kernel's pass over the state); adjoint: the seed lambda = w (.) psi built from the psi
//         tile in shared memory (PA.seed_grad; replaces seed_probs_kernel's read + write and this sweep's read of lambda).  The
//         full-tile kernels compile these paths only when FUSE is set (as run-time branches they cost the hot kernels registers
//         and stack: -2.7 % on every adjoint sweep measured); the generic kernels always carry them and test the pointers.
// Forward sweeps of a circuit that starts from |0...0> may be told to build their tiles in shared memory (PA.zero_init) instead of
// reading a state that a separate pass has just written.
template <bool BWD, bool FULL = false, bool STREAM = false, bool DYN = false, bool FUSE = false>
__global__ void __launch_bounds__(kSweepThreads, BWD ? (STREAM ? 3 : 2) : 3) sweep_flat_kernel(const __grid_constant__ pk::PackedArgs PA) {
  static_assert(!STREAM || (BWD && FULL), "the streaming adjoint kernel handles full tiles only");
  static_assert(!DYN || FULL, "persistent CTAs are built for the full-tile kernels");
  // full-tile kernels that run 3 CTAs per SM (streaming adjoint, forward) use the 16-stage tables: with the 32-stage ones a forward sweep of
  // more than ~30 ops no longer fits three times into an SM's shared memory (measured: config 2's 128-byte-chunk plan, forward +6 %)
  constexpr int NS = (STREAM || (FULL && !BWD)) ? kStreamStages : kMaxFlatStages;
  constexpr bool CAN_FUSE = FUSE || !FULL;
  using Lay = FlatLay<NS>;
  const SweepArgs& A = PA.s;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = A.m, L = A.L;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n_stages = PA.n_stages;
  const uint32_t buf_bytes = 8u << m;                    // one tile of 16-byte units
  unsigned char* buf0 = smem_raw + Lay::kOffBuf;              // FWD: psi buffer 0 / BWD: psi
  unsigned char* buf1 = smem_raw + Lay::kOffBuf + buf_bytes;  // FWD: psi buffer 1 / BWD: lambda
  float* smats = reinterpret_cast<float*>(smem_raw + Lay::kOffBuf + size_t(buf_bytes) * 2);
  float* wacc_all = smats + size_t(A.n_ops) * kMatF;
  float* wred = wacc_all + (BWD ? size_t(kMaxWarps) * A.n_kslots * kAcc : 0);
  auto al = [](size_t x) { return (x + 15) & ~size_t(15); };
  size_t off = Lay::kOffBuf + size_t(buf_bytes) * 2 + size_t(A.n_ops) * kMatF * 4;
  if (BWD) off += (size_t(kMaxWarps) * A.n_kslots * kAcc + size_t(kMaxWarps)) * 4;
  off = al(off);
  uint32_t* hi_off = reinterpret_cast<uint32_t*>(smem_raw + off);
  off = al(off + (size_t(1) << (m - L)) * 4);
  KOp* sops = reinterpret_cast<KOp*>(smem_raw + off);
  SDesc* sdesc = reinterpret_cast<SDesc*>(smem_raw + Lay::kOffDesc);
  uint32_t* stab = reinterpret_cast<uint32_t*>(smem_raw + Lay::kOffStab);
  uint32_t* extc = reinterpret_cast<uint32_t*>(smem_raw + Lay::kOffExtc);
  ExtD* extd = reinterpret_cast<ExtD*>(smem_raw + Lay::kOffExtd);
  uint16_t* ttab = reinterpret_cast<uint16_t*>(smem_raw + Lay::kOffTtab);  // base unit of thread group g = T[g & 15] ^ T[16 + (g >> 4)]
  uint64_t* hik = reinterpret_cast<uint64_t*>(smem_raw + Lay::kOffHik);
  uint64_t* sbase = reinterpret_cast<uint64_t*>(smem_raw + Lay::kOffBase);

  QB_PH_DECL
  int vb = blockIdx.x;  // work item: the CTA index of a static launch (DYN: claimed from the queue after the first)
  int b = vb / A.cps;
  int c = vb % A.cps;
  const uint32_t n_groups = 1u << (m - 4);  // <= blockDim (the planner emits flat stages only for m <= 12)

  // per-sample 2x2s of the work item's sample -> shared memory (adjoint sweep: conjugate transposes); `all`: also the shared ones
  auto load_mats = [&](bool all) {
    for (int i = tid; i < A.n_ops; i += nthr) {
      const KOp kop = all ? A.ops[i] : sops[i];
      if (all) sops[i] = kop;
      const int mat = kop.mat;
      if (!all && (mat < 0 || !(mat & 1))) continue;
      float M[8] = {1, 0, 0, 0, 0, 0, 1, 0};
      if (mat >= 0) {
        const float* src = (mat & 1) ? reinterpret_cast<const float*>(A.mats_batch) + ((size_t)b * A.n_groups_batch + (mat >> 1)) * 8
                                     : reinterpret_cast<const float*>(A.mats_shared) + (size_t)(mat >> 1) * 8;
#pragma unroll
        for (int k = 0; k < 8; ++k) M[k] = src[k];
      }
      float* o = smats + (size_t)i * kMatF;
      float ar, ai, br, bi, cr, ci, dr, di;
      if (BWD) {  // adjoint
        ar = M[0], ai = -M[1], br = M[4], bi = -M[5], cr = M[2], ci = -M[3], dr = M[6], di = -M[7];
      } else {
        ar = M[0], ai = M[1], br = M[2], bi = M[3], cr = M[4], ci = M[5], dr = M[6], di = M[7];
      }
      if (kop.r == 0 && (kop.kind == K_U1 || kop.kind == K_D1)) {  // 2x2 on the pack lane: column pairs (u1_lane_s)
        o[0] = ar, o[1] = cr, o[2] = ai, o[3] = ci, o[4] = br, o[5] = dr, o[6] = bi, o[7] = di;
      } else {
        o[0] = ar, o[1] = ai, o[2] = br, o[3] = bi, o[4] = cr, o[5] = ci, o[6] = dr, o[7] = di;
      }
    }
  };

  // ---- per-CTA setup --------------------------------------------------------------------------------------
  load_mats(true);
  for (int i = tid; i < n_stages; i += nthr) {
    const Stage& st = PA.stages[i];
    SDesc d;
    d.la_begin = (uint16_t)st.pre_end;
    d.la_end = (uint16_t)st.la_end;
    d.d_end = (uint16_t)st.d_end;
    d.shape = (uint8_t)st.shape;
    d.flags = (uint8_t)(((st.xthread & 1) ? kXThread : 0) | (st.n_phase > 0 ? kHasPhase : 0) | (st.d_end > st.pre_end ? kNeedIb : 0) |
                         (((st.xthread >> 4) & 3) << kEndNarrowShift) | (((st.xthread >> 8) & 3) << kXNarrowShift));
    for (int r = 0; r < 4; ++r) {
      d.u_mat[r] = (uint16_t)(st.u_op[r] >= 0 ? st.u_op[r] * kMatF : 0);
      d.u_kslot[r] = (int16_t)(st.u_op[r] >= 0 ? A.ops[st.u_op[r]].kslot : -1);
      d.regbits[r] = (uint8_t)st.regbits[r];
    }
    fill_ext_ranges(d, st, A.ops);
    sdesc[i] = d;
  }
  for (int i = tid; i < n_stages * 2; i += nthr) extd[i] = make_extd(PA.stages[i >> 1], A.ops, i & 1, [](uint32_t x) { return pk::slot_off(x); });
  {
    const int nh = 1 << (m - L);
    for (int h = tid; h < nh; h += nthr) {
      uint64_t o = 0;
      for (int k = 0; k < m - L; ++k) o |= (uint64_t)((h >> k) & 1) << A.tile_bits[L + k];
      hi_off[h] = (uint32_t)(o >> L);
    }
  }
  if (BWD)
    for (int i = tid; i < kMaxWarps * A.n_kslots * kAcc; i += nthr) wacc_all[i] = 0;
  // address tables: linear part of the absorbed CNOT maps applied to the register-bit patterns ...
  for (int i = tid; i < n_stages * 2 * NP; i += nthr) {
    const Stage& st = PA.stages[i / (2 * NP)];
    const int side = (i / NP) & 1, j = i % NP;
    uint32_t x = 0;
    for (int k = 0; k < 3; ++k)
      if ((j >> k) & 1) x |= 1u << st.regbits[k + 1];
    if (side == 0)
      x = pk::absorb_maps<false>(x, A.ops, st.op_begin, st.pre_end, true, 0, false);
    else
      x = pk::absorb_maps<false>(x, A.ops, st.suf_begin, st.op_end, false, 0, false);
    store_tab<!BWD>(stab, i / NP, j, pk::slot_off(x));
  }
  // ... and to the thread-group index g (a thread's 16 amplitudes share the index with the register bits cleared), split
  // into nibbles
  for (int i = tid; i < n_stages * 2 * 32; i += nthr) {
    const int si = i >> 6, side = (i >> 5) & 1, e = i & 31;
    const Stage& st = PA.stages[si];
    const uint32_t g = e < 16 ? (uint32_t)e : (uint32_t)(e - 16) << 4;
    uint32_t x = g << 1;
    x = ins0(x, st.regbits[1]);
    x = ins0(x, st.regbits[2]);
    x = ins0(x, st.regbits[3]);
    if (side == 0)
      x = pk::absorb_maps<false>(x, A.ops, st.op_begin, st.pre_end, true, 0, false);
    else
      x = pk::absorb_maps<false>(x, A.ops, st.suf_begin, st.op_end, false, 0, false);
    ttab[((si * 32 + e) << 1) + side] = (uint16_t)(pk::slot_off(x) >> 4);  // u32 entry e of stage si: load side low half, store side high half
  }
  __syncthreads();
  float* wacc = wacc_all + (BWD ? size_t(tid >> 5) * A.n_kslots * kAcc : 0);

  for (;;) {  // one pass per work item (a single pass unless DYN)
  // (adjoint kernels that may run first in a backward read psi from PA.psi_src when set: the forward's final state stays intact and
  // the un-computed copy goes to A.psi -- no clone pass)
  const void* psi_in = A.psi;
  if constexpr (BWD && CAN_FUSE) psi_in = PA.psi_src ? PA.psi_src : A.psi;
  const float2* gpsi = reinterpret_cast<const float2*>(psi_in) + ((uint64_t)b << A.n_local);
  float2* gpsi_w = reinterpret_cast<float2*>(A.psi) + ((uint64_t)b << A.n_local);
  float2* glam_w = BWD ? reinterpret_cast<float2*>(A.lam) + ((uint64_t)b << A.n_local) : nullptr;
  const uint32_t n_tiles = 1u << (A.n_local - m);
  const int n_vec = (1 << m) >> 1;  // 16-byte vectors (2 amplitudes) per tile
  const int vpc_log = L - 1;        // vectors per contiguous HBM chunk (<= 2^8: low_bits <= 9)
  // tile <-> HBM: thread t moves vectors v = t + 256 k.  With blockDim = 256 >= vectors per chunk, the in-chunk part
  // and the chunk index split as  chunk(v) = chunk(t) | chunk(256 k)  (bit deposits are OR-separable), and the
  // swizzled slot of v is slot(t) + 4096 k: one 64-bit add per vector instead of re-deriving the address.
  const int n_slab = (n_vec + nthr - 1) / nthr;
  if (tid < n_slab) hik[tid] = ((uint64_t)hi_off[(tid * nthr) >> vpc_log] << L) * sizeof(float2);
  // full tiles (2048 vectors on 256 threads): slab k's offset = the bits of k on tile bits 9, 10, 11, in bytes
  // (read from the kernel parameters where they are used: a constant-bank load, not a register that lives across the stage loop)
  auto slab_stride_of = [&](int j) { return (uint64_t)sizeof(float2) << A.tile_bits[FULL ? 9 + j : 0]; };
  const uint64_t my_goff = n_vec > tid ? (((uint64_t)hi_off[tid >> vpc_log] << L) + (uint64_t)((tid & ((1 << vpc_log) - 1)) << 1)) : 0u;
  const uint32_t my_slot = pk::slot_off((uint32_t)tid << 1);
  const bool mover = tid < n_vec;
  if (tid < 2) sbase[tid] = (uint32_t)c + tid * A.cps < n_tiles ? tile_base(A, c + tid * A.cps) : 0;  // CTA-uniform: derived once
  double* pr_acc = nullptr;  // [0] sum |amp|^2, [1 + p] the part with layout bit p set (this work item's tiles)
  float* pr_wred = nullptr;  // per warp: tile total, then S1 of the 12 tile-index bits
  bool fuse_probs = false, fuse_seed = false;
  if constexpr (!BWD && CAN_FUSE) {
    fuse_probs = PA.probs_part != nullptr;
    if (fuse_probs) {
      pr_acc = pr_acc_smem();
      pr_wred = pr_wred_smem();
      if (tid < 49) pr_acc[tid] = 0;
    }
  }
  if constexpr (BWD && CAN_FUSE) fuse_seed = PA.seed_grad != nullptr;
  __syncthreads();

  auto prefetch_tile = [&](unsigned char* dst, const float2* gsrc, uint64_t base_) {
    // |0...0> built in shared memory (plain stores: the barrier after cp_async_wait orders them like the copies)
    if (!BWD && PA.zero_init) {
      if (mover) {
        unsigned char* d = dst + my_slot;
        for (int k = 0; k < n_slab; ++k, d += nthr * 16) *reinterpret_cast<float4*>(d) = make_float4(0.f, 0.f, 0.f, 0.f);
        if (tid == 0 && (base_ | A.rank_bits) == 0) *reinterpret_cast<float*>(dst + my_slot) = 1.f;  // re(amplitude 0) of unit 0
      }
      return;
    }
    if (mover) {
      const char* g0p = reinterpret_cast<const char*>(gsrc + base_ + my_goff);
      uint32_t d = (uint32_t)__cvta_generic_to_shared(dst) + my_slot;
      if constexpr (FULL) {
        // full tiles: slab k = 0..7 of a thread sits at the HBM offset spelled by k's bits on the three highest tile bits (slab_stride_of), so
        // no table read stands between the thread and its copies (under the other CTAs' shared-memory traffic every extra LSU
        // instruction of this phase costs ~100 cycles of queueing)
#pragma unroll 1
        for (int kh = 0; kh < 2; ++kh) {
          const char* gh = g0p + (kh ? slab_stride_of(2) : 0);
#pragma unroll
          for (int kl = 0; kl < 4; ++kl, d += 256 * 16)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gh + ((kl & 1) ? slab_stride_of(0) : 0) + ((kl & 2) ? slab_stride_of(1) : 0)));
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < n_slab; ++k, d += nthr * 16)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g0p + hik[k]));
      }
    }
  };
  if (!BWD && (uint32_t)c < n_tiles) {
    prefetch_tile(buf0, gpsi, sbase[0]);
    pk::cp_async_commit();
  }
  int it = 0;
  QB_PH(0)
  for (uint32_t tau = c; tau < n_tiles; tau += A.cps, ++it) {
    const bool has_next = tau + A.cps < n_tiles;
    // ring of tile offsets: slot it & 3 is this tile, thread 0 derives the one after next (barriers in between)
    const uint64_t base = sbase[it & 3];
    const uint64_t gbase = base | A.rank_bits;
    const uint64_t base_next = sbase[(it + 1) & 3];
    if (tid == 0 && tau + 2 * A.cps < n_tiles) sbase[(it + 2) & 3] = tile_base(A, tau + 2 * A.cps);
    unsigned char* pbuf;  // psi tile
    unsigned char* lbuf;  // lambda tile (BWD)
    if (BWD) {
      pbuf = buf0;
      lbuf = buf1;
      prefetch_tile(pbuf, gpsi, base);
      if (!fuse_seed) prefetch_tile(lbuf, glam_w, base);
      pk::cp_async_commit();
    } else {
      pbuf = (it & 1) ? buf1 : buf0;
      lbuf = nullptr;
      if (has_next) {  // next tile of this CTA into the other buffer while this one is processed
        prefetch_tile((it & 1) ? buf0 : buf1, gpsi, base_next);
        pk::cp_async_commit();
      }
    }
    QB_PH(8)
    // per-tile XOR constants of the CNOTs controlled by out-of-tile bits (uniform over the tile); FULL: + which of the
    // two forward buffers holds this tile (slots are < 32 KB, so XOR with the buffer size adds it)
    const uint32_t bufsel = (FULL && !BWD && (it & 1)) ? kFullBufBytes : 0u;
    for (int i = tid; i < n_stages * 2; i += nthr)
      extc[i] = tile_extc(extd, sdesc, PA.stages, sops, i, gbase, [](uint32_t x) { return pk::slot_off(x); }) ^ bufsel;
    if constexpr (DYN && !BWD) {
      // forward sweeps claim the NEXT work item while the first tile of this one is on its way (thread 0 waits for the atomic where the
      // CTA waits for HBM anyway); it is read after the barrier that ends the item.  (Measured: forward sweeps -3 ... -4 %; the
      // adjoint kernel loses 3 % with the same change -- one more spilled register in its tile loop -- and keeps the claim at the end.)
      if (it == 0 && tid == 0) reinterpret_cast<volatile int*>(sbase)[9] = (int)gridDim.x + atomicAdd(PA.dyn_counter, 1);
    }
    QB_PH(1)
    if (!BWD && has_next)
      pk::cp_async_wait<1>();
    else
      pk::cp_async_wait<0>();
    __syncthreads();
    QB_PH(2)
    if constexpr (BWD && CAN_FUSE) {
      // fused adjoint seed: lambda_i = (sum_q g_q [bit final_pos(q) of i == 0]) psi_i on the tile in shared memory (kernels.cuh:
      // seed_probs_kernel's weights).  A mover thread owns the same slots in both buffers; tile-index bit j is the layout bit
      // tile_bits[j].
      if (fuse_seed) {  // CTA-uniform
        if (mover) {
          const float* g = PA.seed_grad + (size_t)b * PA.seed_n_qubits;
          uint64_t in_tile = 0;
          for (int j = 0; j < m; ++j) in_tile |= uint64_t(1) << A.tile_bits[j];
          float w_base = 0;  // qubits on out-of-tile bits: uniform over the tile
          for (int q = 0; q < PA.seed_n_qubits; ++q)
            if (!(((gbase | in_tile) >> PA.seed_final_pos[q]) & 1)) w_base += g[q];
          float G[12];
#pragma unroll
          for (int j = 0; j < 12; ++j) G[j] = j < m ? g[PA.seed_tile_q[j]] : 0.f;
          const unsigned char* ps = pbuf + my_slot;
          unsigned char* ls = lbuf + my_slot;
          for (int k = 0; k < n_slab; ++k) {
            const uint32_t t = (uint32_t)(tid + k * nthr) << 1;  // tile index of the unit's first amplitude (bit 0 clear)
            float w1 = w_base;
#pragma unroll
            for (int j = 1; j < 12; ++j)
              if (!((t >> j) & 1)) w1 += G[j];  // bits >= m of t are 0 and add G[j] = 0
            const float w0 = w1 + G[0];
            const float4 v = *reinterpret_cast<const float4*>(ps + k * (nthr * 16));
            *reinterpret_cast<float4*>(ls + k * (nthr * 16)) = make_float4(v.x * w0, v.y * w1, v.z * w0, v.w * w1);
          }
        }
        __syncthreads();
      }
    }
    float tdot = 0;
    if (BWD && A.need_tile_dot) {  // Im <lam|psi> over the tile (same slots in both buffers)
      float s = 0;
      for (uint32_t q = tid; q < (1u << (m - 1)); q += nthr) {
        const float4 pu = *reinterpret_cast<const float4*>(pbuf + q * 16), lu = *reinterpret_cast<const float4*>(lbuf + q * 16);
        s += (lu.x * pu.z - lu.z * pu.x) + (lu.y * pu.w - lu.w * pu.y);
      }
      s = warp_sum(s);
      if ((tid & 31) == 0) wred[tid >> 5] = s;
      __syncthreads();
      for (int w = 0; w < (nthr >> 5); ++w) tdot += wred[w];
    }
    if constexpr (STREAM)
      run_stages_stream<NS>(n_stages, gbase, smats, wacc, sops);
    else
      run_stages<BWD, FULL, NS>(pbuf, lbuf, n_stages, n_groups, gbase, tdot, smats, wacc, sops);
    QB_PH(3)
    if constexpr (!BWD && CAN_FUSE) {
      // fused MeasureProbability reduction on the finished tile (measurements.py:113-123; the separate pass is
      // probs_partial_kernel).  A mover thread squares the units it is about to store; S1 of tile-index bit j belongs to layout
      // bit tile_bits[j], the out-of-tile bits take the tile total where the tile's base has them set.
      if (fuse_probs) {  // CTA-uniform
        float sv[13];
#pragma unroll
        for (int j = 0; j < 13; ++j) sv[j] = 0.f;
        if (mover) {
          const unsigned char* ps = pbuf + my_slot;
          for (int k = 0; k < n_slab; ++k) {
            const float4 v = *reinterpret_cast<const float4*>(ps + k * (nthr * 16));
            const float p0 = v.x * v.x + v.z * v.z, p1 = v.y * v.y + v.w * v.w, pp = p0 + p1;
            const uint32_t t = (uint32_t)(tid + k * nthr) << 1;
            sv[0] += pp;
            sv[1] += p1;
#pragma unroll
            for (int j = 1; j < 12; ++j)
              if ((t >> j) & 1) sv[1 + j] += pp;
          }
        }
#pragma unroll
        for (int j = 0; j < 13; ++j) sv[j] = warp_sum(sv[j]);
        if ((tid & 31) == 0) {
#pragma unroll
          for (int j = 0; j < 13; ++j) pr_wred[(tid >> 5) * 13 + j] = sv[j];
        }
        __syncthreads();
        if (tid < 49) {
          float tot = 0;
          for (int w = 0; w < (nthr >> 5); ++w) tot += pr_wred[w * 13];
          if (tid == 0) {
            pr_acc[0] += (double)tot;
          } else {
            const int p = tid - 1;  // layout bit
            int j = -1;
            for (int i = 0; i < m; ++i)
              if (A.tile_bits[i] == p) j = i;
            if (j >= 0) {
              float s1 = 0;
              for (int w = 0; w < (nthr >> 5); ++w) s1 += pr_wred[w * 13 + 1 + j];
              pr_acc[tid] += (double)s1;
            } else if ((gbase >> p) & 1) {
              pr_acc[tid] += (double)tot;
            }
          }
        }
        // (pr_wred is rewritten only after the barrier that ends this tile's store)
      }
    }
    // ---- shared -> HBM (units are already in the HBM layout) ----------------------------------------------------------
    if (mover) {
      char* p0 = reinterpret_cast<char*>(gpsi_w + base + my_goff);
      char* l0 = BWD ? reinterpret_cast<char*>(glam_w + base + my_goff) : nullptr;
      const unsigned char* ps = pbuf + my_slot;
      const unsigned char* ls = BWD ? lbuf + my_slot : nullptr;
      if constexpr (FULL) {
#pragma unroll 1
        for (int kh = 0; kh < 2; ++kh) {
          const uint64_t oh = kh ? slab_stride_of(2) : 0;
#pragma unroll
          for (int kl = 0; kl < 4; ++kl) {
            const uint64_t go = oh + ((kl & 1) ? slab_stride_of(0) : 0) + ((kl & 2) ? slab_stride_of(1) : 0);
            const int k = kh * 4 + kl;
            __stcs(reinterpret_cast<float4*>(p0 + go), *reinterpret_cast<const float4*>(ps + k * (256 * 16)));
            if (BWD) __stcs(reinterpret_cast<float4*>(l0 + go), *reinterpret_cast<const float4*>(ls + k * (256 * 16)));
          }
        }
      } else {
#pragma unroll 4
        for (int k = 0; k < n_slab; ++k) {
          const uint64_t go = hik[k];
          __stcs(reinterpret_cast<float4*>(p0 + go), *reinterpret_cast<const float4*>(ps + k * (nthr * 16)));
          if (BWD) __stcs(reinterpret_cast<float4*>(l0 + go), *reinterpret_cast<const float4*>(ls + k * (nthr * 16)));
        }
      }
    }
    QB_PH(4)
    __syncthreads();
    QB_PH(5)
    QB_PH_COUNT(6)
  }
  if constexpr (!BWD && CAN_FUSE) {
    if (fuse_probs) {  // the tile loop ended with a barrier: pr_acc is complete
      double* row = PA.probs_part + (size_t)vb * kProbPartStride;
      if (tid < kProbPartStride) {
        double v = 0;
        if (tid == 0)
          v = pr_acc[0];
        else if (tid <= kProbSegBits)
          v = pr_acc[0] - 2.0 * pr_acc[tid];  // W_p = S0 - S1, layout bits 0..9
        else if (tid - 1 < 48)
          v = pr_acc[tid];  // S1 of layout bit tid - 1 >= 10 sits at row[11 + (p - 10)] = row[tid]
        row[tid] = v;
      }
    }
  }
  if (BWD) {
    float* out = reinterpret_cast<float*>(A.partials) + (size_t)vb * A.n_kslots * kAcc;
    for (int i = tid; i < A.n_kslots * kAcc; i += nthr) {
      float s = 0;
      for (int w = 0; w < (nthr >> 5); ++w) s += wacc_all[(size_t)w * A.n_kslots * kAcc + i];
      out[i] = s;
    }
  }
  if constexpr (!DYN) {
    QB_PH(0)
    break;
  } else {
    // next work item; its sample's matrices replace the current ones, the gradient accumulators restart from zero
    __syncthreads();  // the partial sums above have been read, nobody still uses smats
    if (BWD || n_tiles <= (uint32_t)c) {  // (forward: an item without tiles claimed nothing -- cannot happen, cps <= n_tiles)
      if (tid == 0) reinterpret_cast<volatile int*>(sbase)[9] = (int)gridDim.x + atomicAdd(PA.dyn_counter, 1);  // (sbase ring: ints 0-7)
      __syncthreads();
    }
    vb = reinterpret_cast<volatile int*>(sbase)[9];
    if (vb >= PA.dyn_items) {
      QB_PH(0)
      break;
    }
    const int b_new = vb / A.cps;
    c = vb % A.cps;
    if (b_new != b) {
      b = b_new;
      load_mats(false);  // only the per-sample 2x2s change
    }
    if (BWD)
      for (int i = tid; i < kMaxWarps * A.n_kslots * kAcc; i += nthr) wacc_all[i] = 0;
    __syncthreads();
  }
  }  // work items
  QB_PH_FLUSH
}

}  // namespace fl
}  // namespace qb
