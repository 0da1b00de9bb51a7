// Packed complex64 sweep kernel (sm_100a): the workhorse for complex64 states.
//
// Same stage model as kernels.cuh (plan.h: Stage), but built around Blackwell's packed fp32 pipe:
//  * complex64 states are kept PACK-PLANAR in HBM and shared memory: 16-byte units (re0, re1, im0, im1) of the amplitude
//    pair (i, i|1).  A *pack* is one float2 half of a unit: local bit 0 is the pack lane and a register bit of every
//    stage.  A thread owns 8 units = 16 amplitudes (register bits {0, b1, b2, b3}); one LDS.128 / STS.128 per unit, and
//    tiles move HBM <-> shared with cp.async (forward: double-buffered prefetch of the CTA's next tile).
//  * a 2x2 on b1..b3 is 16 FFMA2/FMUL2 per pack pair (8 per amplitude pair instead of 16 scalar FFMA); a 2x2 on bit 0
//    mixes the two lanes and costs the scalar count.
//  * units are stored at slot q ^ f((q >> 3) & 7) (q = i >> 1; f: see slot_off): 128-bit accesses of consecutive threads on consecutive
//    units (tile load/store, stages on high bits) and of threads 16 amplitudes apart (stages on low bits) are both
//    bank-conflict free.  The swizzle is GF(2)-linear, so the address of register pack j is
//    slot(i_base') XOR table[stage][j] with a tile-independent table.
//  * CNOTs at the start / end of a stage never move data: as index maps i -> i ^ (bit_c(i) << t) they are linear, so
//    they are folded into the load / store address (table for register-bit controls, a per-thread XOR for thread-bit
//    and out-of-tile controls).
#pragma once
#include "kernels.cuh"

namespace qb {
namespace pk {

constexpr int NP = 8;         // packs per thread and plane
constexpr int kMatFloats = 32;  // per op: 12 broadcast pairs + the raw 2x2

// byte offset in a tile buffer of the 16-byte unit (re0, re1, im0, im1) of the amplitude pair (i, i|1).  Units are
// stored at q ^ f((q >> 3) & 7): 128-bit accesses of 8 consecutive threads hit 8 different 16-byte bank groups both for
// consecutive units and for threads 8 units apart.  GF(2)-linear.
__device__ __forceinline__ uint32_t slot_off(uint32_t i) {
  uint32_t q = i >> 1;
#ifndef QB_SWIZZLE_IDENTITY
  // Unit bits 3, 4, 5 fold into the bank-group bits as 111b, 110b, 101b instead of 001b, 010b, 100b.  A quarter-warp's eight
  // threads differ in the three lowest NON-register unit bits; with the identity fold every triple that contains both u_i
  // and u_(i+3) lands in 4 bank groups (12 of the 20 triples of u0..u5: 2-way conflicts in ~20 % of the stage accesses of
  // the bench plans, ncu: 19-24 % of the adjoint sweeps' wavefronts); with this fold only 4 triples do
  // (tools/bank_conflicts.py: 1.21-1.25 -> 1.02-1.05 wavefronts per access; measured +3.3 % evals/s on config 2, +2.4 % at 20
  // qubits, +1.7 % on config 3, profiles/r1_final_swizzle_ab.md).  Still GF(2)-linear, still inside unit bits 0-7.
  // -DQB_SWIZZLE_IDENTITY builds the old fold q ^ ((q >> 3) & 7) for comparison.
  q ^= (0x8D53B8u >> (3u * ((q >> 3) & 7u))) & 7u;  // f(h) = 7 h0 ^ 6 h1 ^ 5 h2 for h = 0..7: 0, 7, 6, 1, 5, 2, 3, 4
#else
  q ^= (q >> 3) & 7u;
#endif
  return q << 4;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const uint32_t d = (uint32_t)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;"); }
template <int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N)); }

__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }

// ---- 2x2 on pack-index bit RBIT: out0 = a x + b y, out1 = c x + d y (complex), constants pre-broadcast -------------
template <int RBIT>
__device__ __forceinline__ void u1_pack(float2 (&R)[NP], float2 (&I)[NP], const float2* C) {
  // C: 0 (ar,ar) 1 (-ai,-ai) 2 (br,br) 3 (-bi,-bi) 4 (ai,ai) 5 (bi,bi) 6 (cr,cr) 7 (-ci,-ci) 8 (dr,dr) 9 (-di,-di)
  //    10 (ci,ci) 11 (di,di)
  const float4* C4 = reinterpret_cast<const float4*>(C);
  const float4 c0 = C4[0], c1 = C4[1], c2 = C4[2], c3 = C4[3], c4 = C4[4], c5 = C4[5];
  const float2 ar = {c0.x, c0.y}, nai = {c0.z, c0.w}, br = {c1.x, c1.y}, nbi = {c1.z, c1.w}, ai = {c2.x, c2.y}, bi = {c2.z, c2.w};
  const float2 cr = {c3.x, c3.y}, nci = {c3.z, c3.w}, dr = {c4.x, c4.y}, ndi = {c4.z, c4.w}, ci = {c5.x, c5.y}, di = {c5.z, c5.w};
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (j & (1 << RBIT)) continue;
    const int k = j | (1 << RBIT);
    const float2 xr = R[j], xi = I[j], yr = R[k], yi = I[k];
    R[j] = f2fma(nbi, yi, f2fma(br, yr, f2fma(nai, xi, f2mul(ar, xr))));
    I[j] = f2fma(bi, yr, f2fma(br, yi, f2fma(ai, xr, f2mul(ar, xi))));
    R[k] = f2fma(ndi, yi, f2fma(dr, yr, f2fma(nci, xi, f2mul(cr, xr))));
    I[k] = f2fma(di, yr, f2fma(dr, yi, f2fma(ci, xr, f2mul(cr, xi))));
  }
}

// ---- 2x2 on the pack lane (local bit 0): x = lane .x, y = lane .y ---------------------------------------------------
__device__ __forceinline__ void u1_lane(float2 (&R)[NP], float2 (&I)[NP], const float* M) {
  const float ar = M[0], ai = M[1], br = M[2], bi = M[3], cr = M[4], ci = M[5], dr = M[6], di = M[7];
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const float xr = R[j].x, xi = I[j].x, yr = R[j].y, yi = I[j].y;
    R[j].x = fmaf(-bi, yi, fmaf(br, yr, fmaf(-ai, xi, ar * xr)));
    I[j].x = fmaf(bi, yr, fmaf(br, yi, fmaf(ai, xr, ar * xi)));
    R[j].y = fmaf(-di, yi, fmaf(dr, yr, fmaf(-ci, xi, cr * xr)));
    I[j].y = fmaf(di, yr, fmaf(dr, yi, fmaf(ci, xr, cr * xi)));
  }
}

// Pauli vector accumulation on pack-index bit RBIT (both lanes at once); P* collect the positive terms, N* the negative
template <int RBIT>
__device__ __forceinline__ void pauli_pack(const float2 (&R)[NP], const float2 (&I)[NP], const float2 (&LR)[NP],
                                           const float2 (&LI)[NP], float2& px, float2& nx, float2& py, float2& ny,
                                           float2& pz, float2& nz) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    if (j & (1 << RBIT)) continue;
    const int k = j | (1 << RBIT);
    const float2 xr = R[j], xi = I[j], yr = R[k], yi = I[k], lxr = LR[j], lxi = LI[j], lyr = LR[k], lyi = LI[k];
    px = f2fma(lxr, yi, f2fma(lyr, xi, px));
    nx = f2fma(lxi, yr, f2fma(lyi, xr, nx));
    py = f2fma(lyr, xr, f2fma(lyi, xi, py));
    ny = f2fma(lxr, yr, f2fma(lxi, yi, ny));
    pz = f2fma(lxr, xi, f2fma(lyi, yr, pz));
    nz = f2fma(lxi, xr, f2fma(lyr, yi, nz));
  }
}

__device__ __forceinline__ void pauli_lane(const float2 (&R)[NP], const float2 (&I)[NP], const float2 (&LR)[NP],
                                           const float2 (&LI)[NP], float& sx, float& sy, float& sz) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    pauli_acc(sx, sy, sz, float2{R[j].x, I[j].x}, float2{R[j].y, I[j].y}, float2{LR[j].x, LI[j].x},
              float2{LR[j].y, LI[j].y});
  }
}

// ---- fully static register permutations / sign flips (every index is a compile-time constant) --------------------
// XOR swap on the bit patterns: a real in-place exchange.  (A plain swap is turned into register renaming, which
// makes ptxas copy the whole amplitude array at the head of the op loop -- 30 MOVs per op.)
__device__ __forceinline__ void swapf(float& a, float& b) {
  unsigned x = __float_as_uint(a), y = __float_as_uint(b);
  asm volatile("xor.b32 %0, %0, %1;\n\txor.b32 %1, %1, %0;\n\txor.b32 %0, %0, %1;" : "+r"(x), "+r"(y));
  a = __uint_as_float(x);
  b = __uint_as_float(y);
}
// X on target RT (0: the pack lane, 1..3: pack-index bit RT-1) for the amplitudes whose control RC is set
// (RC: 0 lane, 1..3 pack-index bit, 4: unconditional)
template <int RT, int RC>
__device__ __forceinline__ void cx_static(float2 (&R)[NP], float2 (&I)[NP]) {
  if constexpr (RT == 0) {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (RC != 4 && !((j >> (RC - 1)) & 1)) continue;
      swapf(R[j].x, R[j].y);
      swapf(I[j].x, I[j].y);
    }
  } else {
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      if (j & (1 << (RT - 1))) continue;
      const int k = j | (1 << (RT - 1));
      if constexpr (RC == 0) {  // control = lane: only the .y lanes move
        swapf(R[j].y, R[k].y);
        swapf(I[j].y, I[k].y);
      } else {
        if (RC != 4 && !((j >> (RC - 1)) & 1)) continue;
        swapf(R[j].x, R[k].x);
        swapf(R[j].y, R[k].y);
        swapf(I[j].x, I[k].x);
        swapf(I[j].y, I[k].y);
      }
    }
  }
}

// value of local bit (reg index r) for pack j, lane l:  r == 0 -> l;  r in 1..3 -> bit (r-1) of j;  r < 0 -> thread bit
__device__ __forceinline__ bool bit_of(int r, bool thread_bit, int j, int l) {
  return r == 0 ? (l != 0) : (r > 0 ? ((j >> (r - 1)) & 1) : thread_bit);
}

// negate amplitudes where  ok && bit(a) [&& bit(c)]   (CZ family; diagonal, so it rides along in any stage)
__device__ __forceinline__ void negate_where(float2 (&R)[NP], float2 (&I)[NP], bool ok, int ra, bool ta, bool use_a, int rc,
                                             bool tc, bool use_c) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const bool p0 = ok && (!use_a || bit_of(ra, ta, j, 0)) && (!use_c || bit_of(rc, tc, j, 0));
    const bool p1 = ok && (!use_a || bit_of(ra, ta, j, 1)) && (!use_c || bit_of(rc, tc, j, 1));
    R[j] = float2{p0 ? -R[j].x : R[j].x, p1 ? -R[j].y : R[j].y};
    I[j] = float2{p0 ? -I[j].x : I[j].x, p1 ? -I[j].y : I[j].y};
  }
}

// amplitude *= d for every amplitude of the thread (diagonal on a thread bit / out-of-tile bit)
__device__ __forceinline__ void diag_all(float2 (&R)[NP], float2 (&I)[NP], float2 d) {
  const float2 dr = {d.x, d.x}, di = {d.y, d.y}, ndi = {-d.y, -d.y};
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const float2 r0 = R[j], i0 = I[j];
    R[j] = f2fma(ndi, i0, f2mul(dr, r0));
    I[j] = f2fma(di, r0, f2mul(dr, i0));
  }
}
// diagonal on register bit RD (0: lane, 1..3: pack-index bit)
template <int RD>
__device__ __forceinline__ void diag_static(float2 (&R)[NP], float2 (&I)[NP], float2 d0, float2 d1) {
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    float2 dr, di;
    if constexpr (RD == 0) {
      dr = float2{d0.x, d1.x};
      di = float2{d0.y, d1.y};
    } else {
      const bool one = (j >> (RD - 1)) & 1;
      dr = one ? float2{d1.x, d1.x} : float2{d0.x, d0.x};
      di = one ? float2{d1.y, d1.y} : float2{d0.y, d0.y};
    }
    const float2 ndi = {-di.x, -di.y};
    const float2 r0 = R[j], i0 = I[j];
    R[j] = f2fma(ndi, i0, f2mul(dr, r0));
    I[j] = f2fma(di, r0, f2mul(dr, i0));
  }
}
// sum over the thread's amplitudes of  (+/-) Im(conj(l) psi), sign - where register bit RD is set
template <int RD>
__device__ __forceinline__ float diag_grad_static(const float2 (&R)[NP], const float2 (&I)[NP], const float2 (&LR)[NP],
                                                  const float2 (&LI)[NP]) {
  float sz = 0;
#pragma unroll
  for (int j = 0; j < NP; ++j) {
    const float im0 = LR[j].x * I[j].x - LI[j].x * R[j].x, im1 = LR[j].y * I[j].y - LI[j].y * R[j].y;
    if constexpr (RD == 0) {
      sz += im0 - im1;
    } else if constexpr (RD == 4) {  // no register bit: caller applies the sign
      sz += im0 + im1;
    } else {
      sz += ((j >> (RD - 1)) & 1) ? -(im0 + im1) : (im0 + im1);
    }
  }
  return sz;
}

// handler ids (pre-decoded once per CTA from KOp::kind / r / rc)
enum Handler : int {
  H_NOP = 0,
  H_U1 = 1,        // + r (0 lane, 1..3 pack bit)                     -> 1..4
  H_D1 = 5,        // + class (0 lane, 1..3 pack bit, 4 thread, 5 ext)  -> 5..10
  H_CZ = 11,       // generic sign flip
  H_CX = 12,       // + rt * 6 + cc  (cc: 0 lane, 1..3 pack bit, 4 thread bit, 5 out-of-tile)  -> 12..35
  H_COUNT = 36
};

__device__ __forceinline__ int handler_of(const KOp& op) {
  switch (op.kind) {
    case K_U1: return H_U1 + op.r;
    case K_D1: return H_D1 + (op.r >= 0 ? op.r : 4);
    case K_D1_EXT: return H_D1 + 5;
    case K_CX: return H_CX + op.r * 6 + (op.rc >= 0 ? op.rc : 4);
    case K_CX_EXT: return H_CX + op.r * 6 + 5;
    case K_CZ:
    case K_CZ_EXT1:
    case K_CZ_EXT2: return H_CZ;
    default: return H_NOP;
  }
}

struct PackedArgs {
  SweepArgs s;
  const Stage* stages;
  int32_t n_stages;
  // persistent mode of the flat complex64 full-tile kernels (flat64.cuh, DYN): the grid is one CTA per resident slot and the
  // `dyn_items` = B * cps (sample, tile-subset) work items are handed out through an atomic counter
  int32_t dyn_items = 0;
  int32_t* dyn_counter = nullptr;
  // flat64.cuh FUSE (adjoint): the first adjoint sweep after MeasureProbability derives lambda = w (.) psi from the psi tile it has
  // just loaded (kernels.cuh: seed_probs_kernel's weights) instead of reading a lambda that a separate pass wrote
  const float* seed_grad = nullptr;         // [B][n_qubits] dL/dprobs, or null: lambda comes from HBM
  const int32_t* seed_final_pos = nullptr;  // qubit -> bit position in the final layout
  int32_t seed_n_qubits = 0;
  int8_t seed_tile_q[16] = {};              // qubit measured on tile bit j
  // flat64.cuh FUSE (forward): the last forward sweep before MeasureProbability also reduces |amp|^2 per index bit; one row of
  // probs_partial_kernel's layout (kernels.cuh: kProbPartStride doubles) per work item, or null
  double* probs_part = nullptr;
  int32_t zero_init = 0;  // flat64.cuh: the forward sweep starts from |0...0> and does not read the state
  const void* psi_src = nullptr;  // flat64.cuh (adjoint, FUSE / generic kernels): read psi from here, write it to s.psi
};

__host__ __device__ inline size_t packed_smem_bytes(int m, int L, int n_ops, int n_kslots, int n_stages, bool backward) {
  size_t b = (size_t(1) << m) * 8 * 2;  // forward: two psi buffers (double-buffered prefetch); backward: psi + lambda
  b += size_t(n_ops) * kMatFloats * 4;
  if (backward) b += size_t(kMaxWarps) * n_kslots * kAcc * 4 + size_t(kMaxWarps) * 4;
  b = (b + 15) & ~size_t(15);
  b += (size_t(1) << (m - L)) * 4;
  b = (b + 15) & ~size_t(15);
  b += size_t(n_ops) * sizeof(KOp);
  b = (b + 15) & ~size_t(15);
  b += size_t(n_stages) * sizeof(Stage);
  b = (b + 15) & ~size_t(15);
  b += size_t(n_stages) * 2 * NP * sizeof(uint32_t);
  return b;
}

// apply the index maps of the absorbed CNOTs [o0, o1) to x (forward order, or reverse).  DECODED: the ops are the
// CTA's pre-decoded copies (original kind in the top byte of word 0)
template <bool DECODED>
__device__ __forceinline__ uint32_t absorb_maps(uint32_t x, const KOp* ops, int o0, int o1, bool reverse, uint64_t gbase,
                                                bool with_ext) {
  const int n = o1 - o0;
  for (int q = 0; q < n; ++q) {
    const KOp& op = ops[reverse ? (o1 - 1 - q) : (o0 + q)];
    const int w0 = reinterpret_cast<const int*>(&op)[0];
    const int kind = DECODED ? (w0 >> 24) : (w0 & 0xFFFF);
    uint32_t ctl;
    if (kind == K_CX)
      ctl = (x >> op.c) & 1u;
    else
      ctl = (with_ext && (gbase & op.ext_mask) == op.ext_mask) ? 1u : 0u;
    x ^= ctl << op.a;
  }
  return x;
}

template <bool BWD>
__global__ void __launch_bounds__(kSweepThreads, BWD ? 2 : 3) sweep_packed_kernel(const __grid_constant__ PackedArgs PA) {
  const SweepArgs& A = PA.s;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = A.m, L = A.L;
  const uint32_t buf_bytes = 8u << m;                 // one tile of 16-byte units
  unsigned char* buf0 = smem_raw;                     // FWD: psi buffer 0 / BWD: psi
  unsigned char* buf1 = smem_raw + buf_bytes;         // FWD: psi buffer 1 / BWD: lambda
  float* smats = reinterpret_cast<float*>(smem_raw + size_t(buf_bytes) * 2);
  float* wacc_all = smats + size_t(A.n_ops) * kMatFloats;
  float* wred = wacc_all + (BWD ? size_t(kMaxWarps) * A.n_kslots * kAcc : 0);
  size_t off = size_t(buf_bytes) * 2 + size_t(A.n_ops) * kMatFloats * 4;
  if (BWD) off += (size_t(kMaxWarps) * A.n_kslots * kAcc + size_t(kMaxWarps)) * 4;
  off = (off + 15) & ~size_t(15);
  uint32_t* hi_off = reinterpret_cast<uint32_t*>(smem_raw + off);
  off = (off + (size_t(1) << (m - L)) * 4 + 15) & ~size_t(15);
  KOp* sops = reinterpret_cast<KOp*>(smem_raw + off);
  off = (off + size_t(A.n_ops) * sizeof(KOp) + 15) & ~size_t(15);
  Stage* sst = reinterpret_cast<Stage*>(smem_raw + off);
  off = (off + size_t(PA.n_stages) * sizeof(Stage) + 15) & ~size_t(15);
  uint32_t* stab = reinterpret_cast<uint32_t*>(smem_raw + off);  // [n_stages][2 (in, out)][NP] byte offsets

  const int b = blockIdx.x / A.cps;
  const int c = blockIdx.x % A.cps;
  const int tid = threadIdx.x, nthr = blockDim.x;

  // ---- per-CTA setup --------------------------------------------------------------------------------------
  for (int i = tid; i < A.n_ops; i += nthr) {
    KOp op = A.ops[i];
    // pre-decode: low 16 bits of word 0 = handler id, top byte = original kind (r / rc keep their bytes ... they are
    // only read through the struct for the generic CZ handler, so store them separately below)
    const int hid = handler_of(op);
    const int kind = op.kind;
    sops[i] = op;
    reinterpret_cast<int*>(sops + i)[0] = hid | (kind << 24);
    // r / rc were in bytes 2..3 of word 0: the CZ handler needs them -> keep a copy in the (unused there) mat field
    if (hid == H_CZ) sops[i].mat = (int)(uint8_t)op.r | ((int)(uint8_t)op.rc << 8);
  }
  for (int i = tid; i < PA.n_stages; i += nthr) sst[i] = PA.stages[i];
  for (int i = tid; i < A.n_ops; i += nthr) {
    const int mat = A.ops[i].mat;
    float M[8] = {1, 0, 0, 0, 0, 0, 1, 0};
    if (mat >= 0) {
      const float* src = (mat & 1) ? reinterpret_cast<const float*>(A.mats_batch) + ((size_t)b * A.n_groups_batch + (mat >> 1)) * 8
                                   : reinterpret_cast<const float*>(A.mats_shared) + (size_t)(mat >> 1) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) M[k] = src[k];
    }
    float ar, ai, br, bi, cr, ci, dr, di;
    if (BWD) {  // adjoint
      ar = M[0], ai = -M[1], br = M[4], bi = -M[5], cr = M[2], ci = -M[3], dr = M[6], di = -M[7];
    } else {
      ar = M[0], ai = M[1], br = M[2], bi = M[3], cr = M[4], ci = M[5], dr = M[6], di = M[7];
    }
    float* o = smats + (size_t)i * kMatFloats;
    const float v[12] = {ar, -ai, br, -bi, ai, bi, cr, -ci, dr, -di, ci, di};
#pragma unroll
    for (int k = 0; k < 12; ++k) {
      o[2 * k] = v[k];
      o[2 * k + 1] = v[k];
    }
    o[24] = ar, o[25] = ai, o[26] = br, o[27] = bi, o[28] = cr, o[29] = ci, o[30] = dr, o[31] = di;
  }
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
  // address tables: linear part of the absorbed maps applied to the register-bit patterns
  for (int i = tid; i < PA.n_stages * 2 * NP; i += nthr) {
    const Stage& st = PA.stages[i / (2 * NP)];
    const int side = (i / NP) & 1, j = i % NP;
    uint32_t x = 0;
    for (int k = 0; k < 3; ++k)
      if ((j >> k) & 1) x |= 1u << st.regbits[k + 1];
    if (side == 0)
      x = absorb_maps<false>(x, A.ops, st.op_begin, st.pre_end, true, 0, false);
    else
      x = absorb_maps<false>(x, A.ops, st.suf_begin, st.op_end, false, 0, false);
    stab[i] = slot_off(x);
  }
  __syncthreads();
  float* wacc = wacc_all + (BWD ? size_t(tid >> 5) * A.n_kslots * kAcc : 0);

  const float2* gpsi = reinterpret_cast<const float2*>(A.psi) + ((uint64_t)b << A.n_local);
  float2* gpsi_w = reinterpret_cast<float2*>(A.psi) + ((uint64_t)b << A.n_local);
  float2* glam_w = BWD ? reinterpret_cast<float2*>(A.lam) + ((uint64_t)b << A.n_local) : nullptr;
  const uint32_t n_tiles = 1u << (A.n_local - m);
  const uint32_t n_groups = m >= 4 ? (1u << (m - 4)) : 1u;
  const int n_vec = (1 << m) >> 1;  // 16-byte vectors (2 amplitudes) per tile
  const int vpc_log = L - 1;

  // asynchronous HBM -> shared copy of one tile (cp.async, 16-byte units land directly in their swizzled slots)
  auto prefetch_tile = [&](unsigned char* dst, const float2* gsrc, uint32_t tau_) {
    const uint64_t base_ = tile_base(A, tau_);
    for (int v = tid; v < n_vec; v += nthr) {
      const int h = v >> vpc_log, w = v & ((1 << vpc_log) - 1);
      cp_async16(dst + slot_off((uint32_t)v << 1), gsrc + base_ + ((uint64_t)hi_off[h] << L) + ((uint64_t)w << 1));
    }
  };
  if (!BWD && (uint32_t)c < n_tiles) {
    prefetch_tile(buf0, gpsi, c);
    cp_async_commit();
  }
  int it = 0;
  for (uint32_t tau = c; tau < n_tiles; tau += A.cps, ++it) {
    const uint64_t base = tile_base(A, tau);
    const uint64_t gbase = base | A.rank_bits;
    unsigned char* pbuf;  // psi tile
    unsigned char* lbuf;  // lambda tile (BWD)
    if (BWD) {
      pbuf = buf0;
      lbuf = buf1;
      prefetch_tile(pbuf, gpsi, tau);
      prefetch_tile(lbuf, glam_w, tau);
      cp_async_commit();
      cp_async_wait<0>();
    } else {
      pbuf = (it & 1) ? buf1 : buf0;
      lbuf = nullptr;
      if (tau + A.cps < n_tiles) {  // next tile of this CTA into the other buffer while this one is processed
        prefetch_tile((it & 1) ? buf0 : buf1, gpsi, tau + A.cps);
        cp_async_commit();
        cp_async_wait<1>();
      } else {
        cp_async_wait<0>();
      }
    }
    __syncthreads();
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
    // ---- stages ---------------------------------------------------------------------------------------------
    for (int sq = 0; sq < PA.n_stages; ++sq) {
      const int si = BWD ? (PA.n_stages - 1 - sq) : sq;
      const Stage st = sst[si];
      const uint32_t* tab_ld = stab + (si * 2 + (BWD ? 1 : 0)) * NP;
      const uint32_t* tab_st = stab + (si * 2 + (BWD ? 0 : 1)) * NP;
      for (uint32_t g0 = 0; g0 < n_groups; g0 += nthr) {
        const uint32_t g = g0 + tid;
        const bool active = g < n_groups;
        float2 R[NP], I[NP], LR[NP], LI[NP];
        uint32_t ib = 0;
        if (active) {
          ib = g << 1;
          ib = ins0(ib, st.regbits[1]);
          ib = ins0(ib, st.regbits[2]);
          ib = ins0(ib, st.regbits[3]);
          // FWD loads through the inverse of the prefix maps, BWD through the suffix maps
          const uint32_t x = BWD ? absorb_maps<true>(ib, sops, st.suf_begin, st.op_end, false, gbase, true)
                                 : absorb_maps<true>(ib, sops, st.op_begin, st.pre_end, true, gbase, true);
          const uint32_t sb = slot_off(x);
          const uint4 ta = reinterpret_cast<const uint4*>(tab_ld)[0], tb4 = reinterpret_cast<const uint4*>(tab_ld)[1];
          const uint32_t tw[NP] = {ta.x, ta.y, ta.z, ta.w, tb4.x, tb4.y, tb4.z, tb4.w};
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const uint32_t o = sb ^ tw[j];
            const float4 pu = *reinterpret_cast<const float4*>(pbuf + o);
            R[j] = float2{pu.x, pu.y};
            I[j] = float2{pu.z, pu.w};
            if (BWD) {
              const float4 lu = *reinterpret_cast<const float4*>(lbuf + o);
              LR[j] = float2{lu.x, lu.y};
              LI[j] = float2{lu.z, lu.w};
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < NP; ++j) R[j] = I[j] = LR[j] = LI[j] = float2{0.f, 0.f};
        }
        // ---- body ops in registers: flat jump table over pre-decoded handlers -----------------------------------------
        const int nb = st.suf_begin - st.pre_end;
        for (int q = 0; q < nb; ++q) {
          const int oi = BWD ? (st.suf_begin - 1 - q) : (st.pre_end + q);
          const int4 ow = *reinterpret_cast<const int4*>(sops + oi);  // {hid | r | rc, a, c, mat}
          const int hid = ow.x & 0xFFFF;
          const float* Mf = smats + (size_t)oi * kMatFloats;
          const float2* C = reinterpret_cast<const float2*>(Mf);
#define QB_U1_CASE(RR, CALLF)                                                                   \
  case H_U1 + RR: {                                                                             \
    if (BWD) {                                                                                  \
      const int kslot = sops[oi].kslot;                                                         \
      if (kslot >= 0) {                                                                         \
        float sx = 0, sy = 0, sz = 0;                                                           \
        if (RR == 0) {                                                                          \
          pauli_lane(R, I, LR, LI, sx, sy, sz);                                                 \
        } else {                                                                                \
          float2 px = {0, 0}, nx = {0, 0}, py = {0, 0}, ny = {0, 0}, pz = {0, 0}, nz = {0, 0};  \
          pauli_pack<(RR > 0 ? RR - 1 : 0)>(R, I, LR, LI, px, nx, py, ny, pz, nz);              \
          sx = (px.x - nx.x) + (px.y - nx.y);                                                   \
          sy = (py.x - ny.x) + (py.y - ny.y);                                                   \
          sz = (pz.x - nz.x) + (pz.y - nz.y);                                                   \
        }                                                                                       \
        warp_accumulate3<float>(sx, sy, sz, wacc + kslot * kAcc);                               \
      }                                                                                         \
    }                                                                                           \
    CALLF(R, I);                                                                                \
    if (BWD) CALLF(LR, LI);                                                                     \
  } break;
#define QB_U1_LANE(RA, IA) u1_lane(RA, IA, Mf + 24)
#define QB_U1_P0(RA, IA) u1_pack<0>(RA, IA, C)
#define QB_U1_P1(RA, IA) u1_pack<1>(RA, IA, C)
#define QB_U1_P2(RA, IA) u1_pack<2>(RA, IA, C)
#define QB_D1_CASE(CL, GRADEXPR, APPLY)                                   \
  case H_D1 + CL: {                                                       \
    const float2 d0 = {Mf[24], Mf[25]}, d1 = {Mf[30], Mf[31]};            \
    if (BWD) {                                                            \
      const int kslot = sops[oi].kslot;                                   \
      if (kslot >= 0) { GRADEXPR; }                                       \
    }                                                                     \
    APPLY(R, I);                                                          \
    if (BWD) APPLY(LR, LI);                                               \
  } break;
#define QB_CX_CASE(RT, CC)                                                         \
  case H_CX + RT * 6 + CC: {                                                       \
    if (CC <= 3) {                                                                 \
      cx_static<RT, (CC <= 3 ? CC : 4)>(R, I);                                     \
      if (BWD) cx_static<RT, (CC <= 3 ? CC : 4)>(LR, LI);                          \
    } else {                                                                       \
      const bool p = CC == 4 ? ((ib >> ow.z) & 1u) : ((gbase & sops[oi].ext_mask) == sops[oi].ext_mask); \
      if (p) {                                                                     \
        cx_static<RT, 4>(R, I);                                                    \
        if (BWD) cx_static<RT, 4>(LR, LI);                                         \
      }                                                                            \
    }                                                                              \
  } break;
          switch (hid) {
            QB_U1_CASE(0, QB_U1_LANE)
            QB_U1_CASE(1, QB_U1_P0)
            QB_U1_CASE(2, QB_U1_P1)
            QB_U1_CASE(3, QB_U1_P2)
#define QB_D1_APPLY0(RA, IA) diag_static<0>(RA, IA, d0, d1)
#define QB_D1_APPLY1(RA, IA) diag_static<1>(RA, IA, d0, d1)
#define QB_D1_APPLY2(RA, IA) diag_static<2>(RA, IA, d0, d1)
#define QB_D1_APPLY3(RA, IA) diag_static<3>(RA, IA, d0, d1)
            QB_D1_CASE(0, warp_accumulate1<float>(diag_grad_static<0>(R, I, LR, LI), wacc + kslot * kAcc), QB_D1_APPLY0)
            QB_D1_CASE(1, warp_accumulate1<float>(diag_grad_static<1>(R, I, LR, LI), wacc + kslot * kAcc), QB_D1_APPLY1)
            QB_D1_CASE(2, warp_accumulate1<float>(diag_grad_static<2>(R, I, LR, LI), wacc + kslot * kAcc), QB_D1_APPLY2)
            QB_D1_CASE(3, warp_accumulate1<float>(diag_grad_static<3>(R, I, LR, LI), wacc + kslot * kAcc), QB_D1_APPLY3)
            case H_D1 + 4: {  // diagonal on a thread bit
              const bool one = (ib >> ow.y) & 1u;
              const float2 d = one ? float2{Mf[30], Mf[31]} : float2{Mf[24], Mf[25]};
              if (BWD) {
                const int kslot = sops[oi].kslot;
                if (kslot >= 0) {
                  const float g = diag_grad_static<4>(R, I, LR, LI);
                  warp_accumulate1<float>(one ? -g : g, wacc + kslot * kAcc);
                }
              }
              diag_all(R, I, d);
              if (BWD) diag_all(LR, LI, d);
            } break;
            case H_D1 + 5: {  // diagonal on an out-of-tile bit: uniform over the tile
              const bool one = (gbase >> sops[oi].ext_bit) & 1ull;
              const float2 d = one ? float2{Mf[30], Mf[31]} : float2{Mf[24], Mf[25]};
              if (BWD) {
                const int kslot = sops[oi].kslot;
                if (kslot >= 0 && tid == 0 && g0 == 0) wacc[kslot * kAcc + 2] += one ? -tdot : tdot;
              }
              diag_all(R, I, d);
              if (BWD) diag_all(LR, LI, d);
            } break;
            case H_CZ: {
              // merged run of sign flips: one 16-bit mask over (lane + 2 * pack) for the whole run
              const int run = sops[oi].ext_bit;
              uint32_t M = 0;
              for (int u = 0; u < run; ++u) {
                const KOp& o2 = sops[BWD ? (oi - u) : (oi + u)];
                const int w0 = reinterpret_cast<const int*>(&o2)[0];
                const int kind2 = w0 >> 24;
                const int r2 = (int)(int8_t)(o2.mat & 0xFF), rc2 = (int)(int8_t)((o2.mat >> 8) & 0xFF);
                uint32_t ok = 1u, ma = 0xFFFFu, mc = 0xFFFFu;
                if (kind2 != K_CZ) ok = ((gbase & o2.ext_mask) == o2.ext_mask) ? 1u : 0u;
                if (kind2 != K_CZ_EXT2) {
                  if (r2 >= 0)
                    ma = 0xFF00F0F0CCCCAAAAull >> (16 * r2) & 0xFFFFu;
                  else
                    ok &= (ib >> o2.a) & 1u;
                }
                if (kind2 == K_CZ) {
                  if (rc2 >= 0)
                    mc = 0xFF00F0F0CCCCAAAAull >> (16 * rc2) & 0xFFFFu;
                  else
                    ok &= (ib >> o2.c) & 1u;
                }
                M ^= ok ? (ma & mc) : 0u;
              }
              q += run - 1;
#pragma unroll
              for (int j = 0; j < NP; ++j) {
                const uint32_t sx = ((M >> (2 * j)) & 1u) << 31, sy = ((M >> (2 * j + 1)) & 1u) << 31;
                R[j] = float2{__uint_as_float(__float_as_uint(R[j].x) ^ sx), __uint_as_float(__float_as_uint(R[j].y) ^ sy)};
                I[j] = float2{__uint_as_float(__float_as_uint(I[j].x) ^ sx), __uint_as_float(__float_as_uint(I[j].y) ^ sy)};
                if (BWD) {
                  LR[j] = float2{__uint_as_float(__float_as_uint(LR[j].x) ^ sx), __uint_as_float(__float_as_uint(LR[j].y) ^ sy)};
                  LI[j] = float2{__uint_as_float(__float_as_uint(LI[j].x) ^ sx), __uint_as_float(__float_as_uint(LI[j].y) ^ sy)};
                }
              }
            } break;
            QB_CX_CASE(0, 1) QB_CX_CASE(0, 2) QB_CX_CASE(0, 3) QB_CX_CASE(0, 4) QB_CX_CASE(0, 5)
            QB_CX_CASE(1, 0) QB_CX_CASE(1, 2) QB_CX_CASE(1, 3) QB_CX_CASE(1, 4) QB_CX_CASE(1, 5)
            QB_CX_CASE(2, 0) QB_CX_CASE(2, 1) QB_CX_CASE(2, 3) QB_CX_CASE(2, 4) QB_CX_CASE(2, 5)
            QB_CX_CASE(3, 0) QB_CX_CASE(3, 1) QB_CX_CASE(3, 2) QB_CX_CASE(3, 4) QB_CX_CASE(3, 5)
            default:
              break;
          }
        }
        if (!active) continue;
        {
          const uint32_t x = BWD ? absorb_maps<true>(ib, sops, st.op_begin, st.pre_end, true, gbase, true)
                                 : absorb_maps<true>(ib, sops, st.suf_begin, st.op_end, false, gbase, true);
          const uint32_t sb = slot_off(x);
          const uint4 ta = reinterpret_cast<const uint4*>(tab_st)[0], tb4 = reinterpret_cast<const uint4*>(tab_st)[1];
          const uint32_t tw[NP] = {ta.x, ta.y, ta.z, ta.w, tb4.x, tb4.y, tb4.z, tb4.w};
#pragma unroll
          for (int j = 0; j < NP; ++j) {
            const uint32_t o = sb ^ tw[j];
            *reinterpret_cast<float4*>(pbuf + o) = float4{R[j].x, R[j].y, I[j].x, I[j].y};
            if (BWD) *reinterpret_cast<float4*>(lbuf + o) = float4{LR[j].x, LR[j].y, LI[j].x, LI[j].y};
          }
        }
      }
      __syncthreads();
    }
    // ---- shared -> HBM (units are already in the HBM layout) ----------------------------------------------------------
    for (int v = tid; v < n_vec; v += nthr) {
      const int h = v >> vpc_log, w = v & ((1 << vpc_log) - 1);
      const uint64_t e = base + ((uint64_t)hi_off[h] << L) + ((uint64_t)w << 1);
      const uint32_t so = slot_off((uint32_t)v << 1);
      __stcs(reinterpret_cast<float4*>(gpsi_w + e), *reinterpret_cast<const float4*>(pbuf + so));
      if (BWD) __stcs(reinterpret_cast<float4*>(glam_w + e), *reinterpret_cast<const float4*>(lbuf + so));
    }
    __syncthreads();
  }
  if (BWD) {
    float* out = reinterpret_cast<float*>(A.partials) + (size_t)blockIdx.x * A.n_kslots * kAcc;
    for (int i = tid; i < A.n_kslots * kAcc; i += nthr) {
      float s = 0;
      for (int w = 0; w < (nthr >> 5); ++w) s += wacc_all[(size_t)w * A.n_kslots * kAcc + i];
      out[i] = s;
    }
  }
}

}  // namespace pk
}  // namespace qb
