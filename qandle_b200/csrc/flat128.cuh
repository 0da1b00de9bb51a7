// Flat complex128 sweep kernel (sm_100a): the flat-stage design of flat64.cuh for interleaved double2 amplitudes.
//
// A 16-byte shared-memory unit is ONE amplitude (re, im), so there is no pack lane: every CNOT is absorbed into the
// stage addressing, a thread owns 8 amplitudes (3 register bits, 32 registers; psi + lambda = 64 in the adjoint sweep)
// and a stage is  [absorbed CNOTs: load address] [sign mask + per-thread phase] [at most one 2x2 per register bit, fully
// unrolled DFMA code, one switch case per shape] [absorbed CNOTs: store address].  Same planner form (plan.cpp:
// schedule_flat with packed = false), same descriptor / table layout, same batched Pauli reduction as flat64.cuh.
#pragma once
#include "flat64.cuh"

namespace qb {
namespace fd {

constexpr int NA = 8;      // amplitudes per thread
constexpr int kMatD = 8;   // per op: the raw 2x2 (ar, ai, br, bi, cr, ci, dr, di) as doubles; adjoint in the backward sweep

using fl::SDesc;
using fl::kHasPhase;
using fl::kNeedIb;
using fl::kXThread;
using fl::kOffBase;
using fl::kOffBuf;
using fl::kOffDesc;
using fl::kOffExtc;
using fl::kOffExtd;
using fl::kOffHik;
using fl::kOffStab;
using fl::kOffTtab;

// byte offset of amplitude i in a tile buffer (16-byte units, same GF(2)-linear swizzle as the complex64 units)
__device__ __forceinline__ uint32_t slot128(uint32_t i) { return (i ^ ((i >> 3) & 7u)) << 4; }
// bits j (of the thread's 8 amplitudes) whose register bit r is set
__device__ __forceinline__ uint32_t reg_pattern8(int r) { return (0xF0CCAAu >> (8 * r)) & 0xFFu; }

// 2x2 on register bit RBIT: out0 = a x + b y, out1 = c x + d y
template <int RBIT>
__device__ __forceinline__ void u1_d(double2 (&V)[NA], const double* M) {
  const double2 m0 = reinterpret_cast<const double2*>(M)[0], m1 = reinterpret_cast<const double2*>(M)[1];
  const double2 m2 = reinterpret_cast<const double2*>(M)[2], m3 = reinterpret_cast<const double2*>(M)[3];
  const double ar = m0.x, ai = m0.y, br = m1.x, bi = m1.y, cr = m2.x, ci = m2.y, dr = m3.x, di = m3.y;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    if (j & (1 << RBIT)) continue;
    const int k = j | (1 << RBIT);
    const double2 x = V[j], y = V[k];
    V[j].x = fma(-bi, y.y, fma(br, y.x, fma(-ai, x.y, ar * x.x)));
    V[j].y = fma(bi, y.x, fma(br, y.y, fma(ai, x.x, ar * x.y)));
    V[k].x = fma(-di, y.y, fma(dr, y.x, fma(-ci, x.y, cr * x.x)));
    V[k].y = fma(di, y.x, fma(dr, y.y, fma(ci, x.x, cr * x.y)));
  }
}

template <int RBIT>
__device__ __forceinline__ void pauli_d(const double2 (&V)[NA], const double2 (&Lm)[NA], double& sx, double& sy, double& sz) {
  sx = sy = sz = 0.0;
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    if (j & (1 << RBIT)) continue;
    const int k = j | (1 << RBIT);
    pauli_acc<double2, double>(sx, sy, sz, V[j], V[k], Lm[j], Lm[k]);
  }
}

// The 2x2s of one stage shape (bits 0..2: register bits) + the shared-memory store; see fl::shape_body
template <bool BWD, int SHAPE, bool FULL>
__device__ __forceinline__ void shape_body_d(double2 (&V)[NA], double2 (&Lm)[NA], const uint4 dw1, const double* smats, double* wacc,
                                             bool active, unsigned char* pbuf, unsigned char* lbuf, uint32_t sb,
                                             const uint32_t* tab_st) {
  const double* M0 = smats + (dw1.x & 0xFFFFu);
  const double* M1 = smats + (dw1.x >> 16);
  const double* M2 = smats + (dw1.y & 0xFFFFu);
  if constexpr (BWD && SHAPE != 0) {
    constexpr int NU = fl::popc4(SHAPE);
    constexpr int P = NU == 1 ? 4 : (NU == 2 ? 8 : 16);
    double v[P];
    int ks[3] = {-1, -1, -1};
    int u = 0;
#pragma unroll
    for (int i = 0; i < P; ++i) v[i] = 0.0;
    if constexpr (SHAPE & 1) {
      pauli_d<0>(V, Lm, v[4 * u], v[4 * u + 1], v[4 * u + 2]);
      ks[u++] = (int)(int16_t)(dw1.z & 0xFFFFu);
    }
    if constexpr (SHAPE & 2) {
      pauli_d<1>(V, Lm, v[4 * u], v[4 * u + 1], v[4 * u + 2]);
      ks[u++] = (int)(int16_t)(dw1.z >> 16);
    }
    if constexpr (SHAPE & 4) {
      pauli_d<2>(V, Lm, v[4 * u], v[4 * u + 1], v[4 * u + 2]);
      ks[u++] = (int)(int16_t)(dw1.w & 0xFFFFu);
    }
    if (!FULL && !active) {
#pragma unroll
      for (int i = 0; i < P; ++i) v[i] = 0.0;
    }
    const double total = fl::warp_transpose_reduce<P, double>(v);
    constexpr int SH = P == 4 ? 3 : (P == 8 ? 2 : 1);  // lanes per value = 1 << SH
    const int lane = threadIdx.x & 31, vi = lane >> SH, uu = vi >> 2, comp = vi & 3;
    const int kslot = uu == 0 ? ks[0] : (uu == 1 ? ks[1] : (uu == 2 ? ks[2] : -1));
    if ((lane & ((1 << SH) - 1)) == 0 && comp < 3 && kslot >= 0) wacc[kslot * kAcc + comp] += total;
  }
  if constexpr (SHAPE & 1) u1_d<0>(V, M0);
  if constexpr (SHAPE & 2) u1_d<1>(V, M1);
  if constexpr (SHAPE & 4) u1_d<2>(V, M2);
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint32_t tw[NA];
  fl::load_tab<!BWD>(tab_st, tw);
  if (FULL || active) {  // adjoint: psi is stored before lambda's 2x2s (see fl::shape_body)
#pragma unroll
    for (int j = 0; j < NA; ++j) {
      const uint32_t o = sb ^ tw[j];
      *reinterpret_cast<double2*>(FULL ? smem_raw + kOffBuf + o : pbuf + o) = V[j];
    }
  }
  if constexpr (BWD) {
    if constexpr (SHAPE & 1) u1_d<0>(Lm, M0);
    if constexpr (SHAPE & 2) u1_d<1>(Lm, M1);
    if constexpr (SHAPE & 4) u1_d<2>(Lm, M2);
    if (FULL || active) {
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        const uint32_t o = sb ^ tw[j];
        *reinterpret_cast<double2*>(FULL ? smem_raw + kOffBuf + fl::kFullBufBytes + o : lbuf + o) = Lm[j];
      }
    }
  }
}

__device__ __forceinline__ void negate_masked(double2 (&V)[NA], uint32_t M) {
#pragma unroll
  for (int j = 0; j < NA; ++j) {
    const int s = (int)((M << (31 - j)) & 0x80000000u);
    V[j].x = __hiloint2double(__double2hiint(V[j].x) ^ s, __double2loint(V[j].x));
    V[j].y = __hiloint2double(__double2hiint(V[j].y) ^ s, __double2loint(V[j].y));
  }
}

// All stages of one tile (execution order; the adjoint sweep has its own list); not inlined, see fl::run_stages
// FULL (2^11 tile on 256 threads): constant buffer offsets and no partial-warp handling, as fl::run_stages
template <bool BWD, bool FULL>
__device__ __noinline__ void run_stages_d(unsigned char* pbuf, unsigned char* lbuf, const int n_stages, const uint32_t n_groups,
                                          const uint64_t gbase, const double tdot, const double* smats, double* wacc,
                                          const KOp* sops) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int tid = threadIdx.x;
  const uint16_t* ttab = reinterpret_cast<const uint16_t*>(smem_raw + kOffTtab);
  const bool warp_busy = FULL || (uint32_t)(tid & ~31) < n_groups;  // whole warps idle when the tile is small
  const bool active = FULL || (uint32_t)tid < n_groups;             // idle lanes of a partial warp shadow the last group
  const uint32_t my_g = active ? (uint32_t)tid : n_groups - 1;
  const uint32_t* tt_lo = reinterpret_cast<const uint32_t*>(ttab) + (my_g & 15);  // {load, store} base units as one 32-bit load
  const uint32_t* tt_hi = reinterpret_cast<const uint32_t*>(ttab) + 16 + (my_g >> 4);
  for (int si = 0; si < n_stages; ++si) {
    const unsigned char* sp = smem_raw + si * 32;
    const uint4 dw0 = *reinterpret_cast<const uint4*>(sp + kOffDesc);
    const uint2 ks = *reinterpret_cast<const uint2*>(sp + kOffDesc + 16);
    const uint4 dw1 = {dw0.z, dw0.w, ks.x, ks.y};
    const int shape = (dw0.y >> 16) & 0xFF, flags = dw0.y >> 24;
    const uint32_t* tab_ld = reinterpret_cast<const uint32_t*>(sp + si * 32 + kOffStab);
    const uint32_t* tab_st = tab_ld + NA;
    const uint2 ex = *reinterpret_cast<const uint2*>(smem_raw + kOffExtc + si * 8);
    const uint32_t ttx = tt_lo[si * 32] ^ tt_hi[si * 32];  // base unit of this thread's group: load side | store side << 16
    double2 V[NA], Lm[NA];
    if (warp_busy) {
      const uint32_t sbl = ((ttx & 0xFFFFu) << 4) ^ ex.x;
      uint32_t tw[NA];
      fl::load_tab<!BWD>(tab_ld, tw);
#pragma unroll
      for (int j = 0; j < NA; ++j) {
        const uint32_t o = sbl ^ tw[j];
        V[j] = *reinterpret_cast<const double2*>(FULL ? smem_raw + kOffBuf + o : pbuf + o);
        if (BWD) Lm[j] = *reinterpret_cast<const double2*>(FULL ? smem_raw + kOffBuf + fl::kFullBufBytes + o : lbuf + o);
      }
    }
    // absorbed CNOTs whose target is a thread bit move amplitudes between threads: all loads before the first store
    if (flags & kXThread) fl::group_barrier((flags >> fl::kXNarrowShift) & 3);
    if (warp_busy) {
      if (flags & kNeedIb) {  // sign mask, per-thread phase (+ its gradients)
        const int la_end = dw0.x >> 16, d_end = dw0.y & 0xFFFF;
        const uint32_t rbw = *reinterpret_cast<const uint32_t*>(sp + kOffDesc + 24);  // regbits[0..2]
        uint32_t ib = my_g;
        ib = ins0(ib, rbw & 0xFF);
        ib = ins0(ib, (rbw >> 8) & 0xFF);
        ib = ins0(ib, (rbw >> 16) & 0xFF);
        uint32_t M = 0;
        double2 ph = {1.0, 0.0};
        double gsum = 0.0;
        if (BWD && (flags & kHasPhase)) {
#pragma unroll
          for (int j = 0; j < NA; ++j) gsum += Lm[j].x * V[j].y - Lm[j].y * V[j].x;  // Im(conj(lam) psi)
          if (!FULL && !active) gsum = 0.0;
        }
        for (int i = la_end; i < d_end; ++i) {
          const KOp& o = sops[i];
          const int kind = o.kind;
          if (kind == K_D1 || kind == K_D1_EXT) {
            const double* Mf = smats + (size_t)i * kMatD;
            const bool one = kind == K_D1 ? ((ib >> o.a) & 1u) : ((gbase >> o.ext_bit) & 1ull);
            const double2 d = one ? double2{Mf[6], Mf[7]} : double2{Mf[0], Mf[1]};
            ph = cmul(ph, d);
            if (BWD && o.kslot >= 0) {
              if (kind == K_D1)
                warp_accumulate1<double>(one ? -gsum : gsum, wacc + o.kslot * kAcc);
              else if (tid == 0)
                wacc[o.kslot * kAcc + 2] += one ? -tdot : tdot;
            }
          } else {
            uint32_t ok = 1u, ma = 0xFFu, mc = 0xFFu;
            if (kind != K_CZ) ok = ((gbase & o.ext_mask) == o.ext_mask) ? 1u : 0u;
            if (kind != K_CZ_EXT2) {
              if (o.r >= 0)
                ma = reg_pattern8(o.r);
              else
                ok &= (ib >> o.a) & 1u;
            }
            if (kind == K_CZ) {
              if (o.rc >= 0)
                mc = reg_pattern8(o.rc);
              else
                ok &= (ib >> o.c) & 1u;
            }
            M ^= ok ? (ma & mc) : 0u;
          }
        }
        if (M) {
          negate_masked(V, M);
          if (BWD) negate_masked(Lm, M);
        }
        if (flags & kHasPhase) {
#pragma unroll
          for (int j = 0; j < NA; ++j) {
            V[j] = cmul(V[j], ph);
            if (BWD) Lm[j] = cmul(Lm[j], ph);
          }
        }
      }
      const uint32_t sbs = ((ttx >> 16) << 4) ^ ex.y;
#define QB_SHAPE_D(S) \
  case S: shape_body_d<BWD, S, FULL>(V, Lm, dw1, smats, wacc, active, pbuf, lbuf, sbs, tab_st); break;
      switch (shape & 7) {
        QB_SHAPE_D(0) QB_SHAPE_D(1) QB_SHAPE_D(2) QB_SHAPE_D(3) QB_SHAPE_D(4) QB_SHAPE_D(5) QB_SHAPE_D(6)
        default: shape_body_d<BWD, 7, FULL>(V, Lm, dw1, smats, wacc, active, pbuf, lbuf, sbs, tab_st); break;
      }
#undef QB_SHAPE_D
    }
    fl::group_barrier((flags >> fl::kEndNarrowShift) & 3);
  }
}

// shared-memory layout (dynamic): as flat64.cuh with 16-byte amplitude units and double tables
__host__ __device__ inline size_t flat128_smem_bytes(int m, int L, int n_ops, int n_kslots, int n_stages, bool backward) {
  auto al = [](size_t x) { return (x + 15) & ~size_t(15); };
  size_t b = kOffBuf + (size_t(1) << m) * 16 * 2;  // forward: two psi buffers (prefetch); backward: psi + lambda
  b += size_t(n_ops) * kMatD * 8;
  if (backward) b += size_t(kMaxWarps) * n_kslots * kAcc * 8 + size_t(kMaxWarps) * 8;
  b = al(b);
  b = al(b + (size_t(1) << (m - L)) * 4);
  b = al(b + size_t(n_ops) * sizeof(KOp));
  return b;
}

// threads per CTA: one per 8 amplitudes of the tile, at most 256 (flat complex128 tiles are at most 2^11 amplitudes)
__host__ __device__ inline int flat128_threads(int m, int L) {
  int t = 1 << (m > 3 ? m - 3 : 0);
  if (t > kSweepThreads) t = kSweepThreads;
  if (t < 64) t = 64;
  if (t < (1 << L)) t = 1 << L;
  return t;
}

template <bool BWD, bool FULL = false>
__global__ void __launch_bounds__(kSweepThreads, BWD ? 2 : 3) sweep_flat128_kernel(const __grid_constant__ pk::PackedArgs PA) {
  const SweepArgs& A = PA.s;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = A.m, L = A.L;
  const int tid = threadIdx.x, nthr = blockDim.x;
  const int n_stages = PA.n_stages;
  const uint32_t buf_bytes = 16u << m;  // one tile of 16-byte amplitudes
  unsigned char* buf0 = smem_raw + kOffBuf;
  unsigned char* buf1 = smem_raw + kOffBuf + buf_bytes;
  double* smats = reinterpret_cast<double*>(smem_raw + kOffBuf + size_t(buf_bytes) * 2);
  double* wacc_all = smats + size_t(A.n_ops) * kMatD;
  double* wred = wacc_all + (BWD ? size_t(kMaxWarps) * A.n_kslots * kAcc : 0);
  auto al = [](size_t x) { return (x + 15) & ~size_t(15); };
  size_t off = kOffBuf + size_t(buf_bytes) * 2 + size_t(A.n_ops) * kMatD * 8;
  if (BWD) off += (size_t(kMaxWarps) * A.n_kslots * kAcc + size_t(kMaxWarps)) * 8;
  off = al(off);
  uint32_t* hi_off = reinterpret_cast<uint32_t*>(smem_raw + off);
  off = al(off + (size_t(1) << (m - L)) * 4);
  KOp* sops = reinterpret_cast<KOp*>(smem_raw + off);
  SDesc* sdesc = reinterpret_cast<SDesc*>(smem_raw + kOffDesc);
  uint32_t* stab = reinterpret_cast<uint32_t*>(smem_raw + kOffStab);
  uint32_t* extc = reinterpret_cast<uint32_t*>(smem_raw + kOffExtc);
  fl::ExtD* extd = reinterpret_cast<fl::ExtD*>(smem_raw + kOffExtd);
  uint16_t* ttab = reinterpret_cast<uint16_t*>(smem_raw + kOffTtab);
  uint64_t* hik = reinterpret_cast<uint64_t*>(smem_raw + kOffHik);
  uint64_t* sbase = reinterpret_cast<uint64_t*>(smem_raw + kOffBase);

  const int b = blockIdx.x / A.cps;
  const int c = blockIdx.x % A.cps;
  const uint32_t n_groups = 1u << (m - 3);  // <= blockDim

  // ---- per-CTA setup --------------------------------------------------------------------------------------
  for (int i = tid; i < A.n_ops; i += nthr) {
    const KOp kop = A.ops[i];
    sops[i] = kop;
    const int mat = kop.mat;
    double M[8] = {1, 0, 0, 0, 0, 0, 1, 0};
    if (mat >= 0) {
      const double* src = (mat & 1) ? reinterpret_cast<const double*>(A.mats_batch) + ((size_t)b * A.n_groups_batch + (mat >> 1)) * 8
                                    : reinterpret_cast<const double*>(A.mats_shared) + (size_t)(mat >> 1) * 8;
#pragma unroll
      for (int k = 0; k < 8; ++k) M[k] = src[k];
    }
    double* o = smats + (size_t)i * kMatD;
    if (BWD) {  // adjoint
      o[0] = M[0], o[1] = -M[1], o[2] = M[4], o[3] = -M[5], o[4] = M[2], o[5] = -M[3], o[6] = M[6], o[7] = -M[7];
    } else {
#pragma unroll
      for (int k = 0; k < 8; ++k) o[k] = M[k];
    }
  }
  for (int i = tid; i < n_stages; i += nthr) {
    const Stage& st = PA.stages[i];
    SDesc d;
    d.la_begin = (uint16_t)st.pre_end;
    d.la_end = (uint16_t)st.la_end;
    d.d_end = (uint16_t)st.d_end;
    d.shape = (uint8_t)st.shape;
    d.flags = (uint8_t)(((st.xthread & 1) ? kXThread : 0) | (st.n_phase > 0 ? kHasPhase : 0) | (st.d_end > st.la_end ? kNeedIb : 0) |
                         (((st.xthread >> 4) & 3) << fl::kEndNarrowShift) | (((st.xthread >> 8) & 3) << fl::kXNarrowShift));
    for (int r = 0; r < 4; ++r) {
      d.u_mat[r] = (uint16_t)(st.u_op[r] >= 0 ? st.u_op[r] * kMatD : 0);
      d.u_kslot[r] = (int16_t)(st.u_op[r] >= 0 ? A.ops[st.u_op[r]].kslot : -1);
      d.regbits[r] = (uint8_t)(st.regbits[r] >= 0 ? st.regbits[r] : 0);
    }
    fl::fill_ext_ranges(d, st, A.ops);
    sdesc[i] = d;
  }
  for (int i = tid; i < n_stages * 2; i += nthr) extd[i] = fl::make_extd(PA.stages[i >> 1], A.ops, i & 1, [](uint32_t x) { return slot128(x); });
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
  for (int i = tid; i < n_stages * 2 * NA; i += nthr) {
    const Stage& st = PA.stages[i / (2 * NA)];
    const int side = (i / NA) & 1, j = i % NA;
    uint32_t x = 0;
    for (int k = 0; k < 3; ++k)
      if ((j >> k) & 1) x |= 1u << st.regbits[k];
    if (side == 0)
      x = pk::absorb_maps<false>(x, A.ops, st.op_begin, st.pre_end, true, 0, false);
    else
      x = pk::absorb_maps<false>(x, A.ops, st.suf_begin, st.op_end, false, 0, false);
    fl::store_tab<!BWD>(stab, i / NA, j, slot128(x));
  }
  // ... and to the thread-group index g, split into nibbles
  for (int i = tid; i < n_stages * 2 * 32; i += nthr) {
    const int si = i >> 6, side = (i >> 5) & 1, e = i & 31;
    const Stage& st = PA.stages[si];
    uint32_t x = e < 16 ? (uint32_t)e : (uint32_t)(e - 16) << 4;
    x = ins0(x, st.regbits[0]);
    x = ins0(x, st.regbits[1]);
    x = ins0(x, st.regbits[2]);
    if (side == 0)
      x = pk::absorb_maps<false>(x, A.ops, st.op_begin, st.pre_end, true, 0, false);
    else
      x = pk::absorb_maps<false>(x, A.ops, st.suf_begin, st.op_end, false, 0, false);
    ttab[((si * 32 + e) << 1) + side] = (uint16_t)(slot128(x) >> 4);  // u32 entry e of stage si: load side low half, store side high half
  }
  __syncthreads();
  double* wacc = wacc_all + (BWD ? size_t(tid >> 5) * A.n_kslots * kAcc : 0);

  const double2* gpsi = reinterpret_cast<const double2*>(A.psi) + ((uint64_t)b << A.n_local);
  double2* gpsi_w = reinterpret_cast<double2*>(A.psi) + ((uint64_t)b << A.n_local);
  double2* glam_w = BWD ? reinterpret_cast<double2*>(A.lam) + ((uint64_t)b << A.n_local) : nullptr;
  const uint32_t n_tiles = 1u << (A.n_local - m);
  const int n_vec = 1 << m;  // 16-byte vectors (1 amplitude) per tile
  const int vpc_log = L;     // vectors per contiguous HBM chunk
  // tile <-> HBM: thread t moves vectors v = t + nthr k (address split as in flat64.cuh)
  const int n_slab = (n_vec + nthr - 1) / nthr;
  if (tid < n_slab) hik[tid] = ((uint64_t)hi_off[(tid * nthr) >> vpc_log] << L) * sizeof(double2);
  const uint64_t my_goff = n_vec > tid ? (((uint64_t)hi_off[tid >> vpc_log] << L) + (uint64_t)(tid & ((1 << vpc_log) - 1))) : 0u;
  const uint32_t my_slot = slot128((uint32_t)tid);
  const bool mover = tid < n_vec;
  if (tid < 2) sbase[tid] = (uint32_t)c + tid * A.cps < n_tiles ? tile_base(A, c + tid * A.cps) : 0;
  __syncthreads();

  auto prefetch_tile = [&](unsigned char* dst, const double2* gsrc, uint64_t base_) {
    if (mover) {
      const char* g0p = reinterpret_cast<const char*>(gsrc + base_ + my_goff);
      uint32_t d = (uint32_t)__cvta_generic_to_shared(dst) + my_slot;
#pragma unroll 4
      for (int k = 0; k < n_slab; ++k, d += nthr * 16)
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(g0p + hik[k]));
    }
  };
  if (!BWD && (uint32_t)c < n_tiles) {
    prefetch_tile(buf0, gpsi, sbase[0]);
    pk::cp_async_commit();
  }
  int it = 0;
  for (uint32_t tau = c; tau < n_tiles; tau += A.cps, ++it) {
    const bool has_next = tau + A.cps < n_tiles;
    const uint64_t base = sbase[it & 3];
    const uint64_t base_next = sbase[(it + 1) & 3];
    if (tid == 0 && tau + 2 * A.cps < n_tiles) sbase[(it + 2) & 3] = tile_base(A, tau + 2 * A.cps);
    const uint64_t gbase = base | A.rank_bits;
    unsigned char* pbuf;
    unsigned char* lbuf;
    if (BWD) {
      pbuf = buf0;
      lbuf = buf1;
      prefetch_tile(pbuf, gpsi, base);
      prefetch_tile(lbuf, glam_w, base);
      pk::cp_async_commit();
    } else {
      pbuf = (it & 1) ? buf1 : buf0;
      lbuf = nullptr;
      if (has_next) {
        prefetch_tile((it & 1) ? buf0 : buf1, gpsi, base_next);
        pk::cp_async_commit();
      }
    }
    const uint32_t bufsel = (FULL && !BWD && (it & 1)) ? fl::kFullBufBytes : 0u;  // which forward buffer holds this tile
    for (int i = tid; i < n_stages * 2; i += nthr)
      extc[i] = fl::tile_extc(extd, sdesc, PA.stages, sops, i, gbase, [](uint32_t x) { return slot128(x); }) ^ bufsel;
    if (!BWD && has_next)
      pk::cp_async_wait<1>();
    else
      pk::cp_async_wait<0>();
    __syncthreads();
    double tdot = 0;
    if (BWD && A.need_tile_dot) {  // Im <lam|psi> over the tile (same slots in both buffers)
      double s = 0;
      for (uint32_t q = tid; q < (1u << m); q += nthr) {
        const double2 pu = *reinterpret_cast<const double2*>(pbuf + q * 16), lu = *reinterpret_cast<const double2*>(lbuf + q * 16);
        s += lu.x * pu.y - lu.y * pu.x;
      }
      s = warp_sum(s);
      if ((tid & 31) == 0) wred[tid >> 5] = s;
      __syncthreads();
      for (int w = 0; w < (nthr >> 5); ++w) tdot += wred[w];
    }
    run_stages_d<BWD, FULL>(pbuf, lbuf, n_stages, n_groups, gbase, tdot, smats, wacc, sops);
    if (mover) {
      char* p0 = reinterpret_cast<char*>(gpsi_w + base + my_goff);
      char* l0 = BWD ? reinterpret_cast<char*>(glam_w + base + my_goff) : nullptr;
      const unsigned char* ps = pbuf + my_slot;
      const unsigned char* ls = BWD ? lbuf + my_slot : nullptr;
#pragma unroll 4
      for (int k = 0; k < n_slab; ++k) {
        const uint64_t go = hik[k];
        __stcs(reinterpret_cast<float4*>(p0 + go), *reinterpret_cast<const float4*>(ps + k * (nthr * 16)));
        if (BWD) __stcs(reinterpret_cast<float4*>(l0 + go), *reinterpret_cast<const float4*>(ls + k * (nthr * 16)));
      }
    }
    __syncthreads();
  }
  if (BWD) {
    double* out = reinterpret_cast<double*>(A.partials) + (size_t)blockIdx.x * A.n_kslots * kAcc;
    for (int i = tid; i < A.n_kslots * kAcc; i += nthr) {
      double s = 0;
      for (int w = 0; w < (nthr >> 5); ++w) s += wacc_all[(size_t)w * A.n_kslots * kAcc + i];
      out[i] = s;
    }
  }
}

}  // namespace fd
}  // namespace qb
