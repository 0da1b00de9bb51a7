// sm_100a kernels of the state-vector engine.  See plan.h for the sweep model and DESIGN.md for the
// roofline of each kernel.  Everything here is templated on the real type T (float: complex64 states,
// double: complex128 states); amplitudes are interleaved (re, im) = T2.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "plan.h"

namespace qb {

template <typename T> struct Vec2;
template <> struct Vec2<float> { using type = float2; };
template <> struct Vec2<double> { using type = double2; };

constexpr int kSweepThreads = 256;
constexpr int kMaxWarps = kSweepThreads / 32;

struct SweepArgs {
  void* psi;            // [B][2^n_local] complex
  void* lam;            // backward only
  const KOp* ops;
  const void* mats_shared;  // [n_groups_shared][8]
  const void* mats_batch;   // [B][n_groups_batch][8]
  void* partials;           // backward: [B*cps][n_kslots][8]
  uint64_t rank_bits;       // rank << n_local
  int32_t n_ops;
  int32_t n_groups_batch;
  int32_t n_kslots;
  int32_t m, L, n_local;
  int32_t cps;              // CTAs per sample
  int32_t need_tile_dot;    // backward: some K_D1_EXT op carries a gradient
  int8_t tile_bits[16];
  int8_t nontile_bits[48];
};

// ---------------------------------------------------------------------------------------------------
// small complex helpers
template <typename T2> __device__ __forceinline__ T2 cmul(T2 a, T2 b) {
  T2 r;
  r.x = a.x * b.x - a.y * b.y;
  r.y = a.x * b.y + a.y * b.x;
  return r;
}
template <typename T2> __device__ __forceinline__ T2 cfma(T2 a, T2 b, T2 c) {  // a*b + c
  T2 r;
  r.x = fma(a.x, b.x, fma(-a.y, b.y, c.x));
  r.y = fma(a.x, b.y, fma(a.y, b.x, c.y));
  return r;
}
template <typename T2> __device__ __forceinline__ T2 cconj(T2 a) {
  a.y = -a.y;
  return a;
}
// acc += a * conj(b)
template <typename T2> __device__ __forceinline__ void cacc_conj(T2& acc, T2 a, T2 b) {
  acc.x = fma(a.x, b.x, fma(a.y, b.y, acc.x));
  acc.y = fma(a.y, b.x, fma(-a.x, b.y, acc.y));
}

__device__ __forceinline__ uint32_t ins0(uint32_t k, int p) { return ((k >> p) << (p + 1)) | (k & ((1u << p) - 1u)); }

template <typename T> __device__ __forceinline__ T warp_sum(T v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// ---------------------------------------------------------------------------------------------------
// INTERNAL STATE LAYOUT.  complex128: interleaved (re, im).  complex64: "pack-planar" 16-byte units
// (re[2k], re[2k+1], im[2k], im[2k+1]) -- one unit = one amplitude pair (i, i|1) = the two float2 operands of the packed
// FFMA2 kernels, so shared-memory and HBM accesses are single 128-bit ops with no register shuffling.  The two
// layouts differ by swapping the middle words of every 16-byte unit (an involution): unit_fix<float>.
template <typename T> __device__ __forceinline__ int4 unit_fix(int4 v) { return v; }
template <> __device__ __forceinline__ int4 unit_fix<float>(int4 v) { return int4{v.x, v.z, v.y, v.w}; }

// amplitude idx of a state in the internal layout
template <typename T>
__device__ __forceinline__ typename Vec2<T>::type load_amp(const typename Vec2<T>::type* s, uint64_t idx) {
  if constexpr (sizeof(T) == 4) {
    const float* f = reinterpret_cast<const float*>(s) + ((idx >> 1) << 2) + (idx & 1);
    return float2{f[0], f[2]};
  } else {
    return s[idx];
  }
}

// ---------------------------------------------------------------------------------------------------
// tile <-> HBM.  A tile is 2^m amplitudes; its chunk h (2^L contiguous amplitudes) lives at element offset
// base + (hi_off[h] << L).  16-byte vector accesses, consecutive threads -> consecutive vectors.
template <typename T>
__device__ __forceinline__ void tile_load(typename Vec2<T>::type* tile, const typename Vec2<T>::type* g, uint64_t base,
                                          const uint32_t* hi_off, int m, int L) {
  constexpr int EPV = 16 / sizeof(typename Vec2<T>::type);  // amplitudes per 16-byte vector: 2 (c64) / 1 (c128)
  constexpr int LEPV = EPV == 2 ? 1 : 0;
  constexpr int UN = 8;  // independent 16-byte requests in flight per thread
  const int n_vec = (1 << m) >> LEPV;
  const int vpc_log = L - LEPV;
  for (int v0 = 0; v0 < n_vec; v0 += blockDim.x * UN) {
    int4 buf[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int v = v0 + u * blockDim.x + threadIdx.x;
      if (v < n_vec) {
        const int h = v >> vpc_log, w = v & ((1 << vpc_log) - 1);
        buf[u] = __ldcs(reinterpret_cast<const int4*>(g + base + ((uint64_t)hi_off[h] << L) + ((uint64_t)w << LEPV)));
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int v = v0 + u * blockDim.x + threadIdx.x;
      if (v < n_vec) {
        const int h = v >> vpc_log, w = v & ((1 << vpc_log) - 1);
        *reinterpret_cast<int4*>(tile + ((h << L) + (w << LEPV))) = unit_fix<T>(buf[u]);
      }
    }
  }
}

template <typename T>
__device__ __forceinline__ void tile_store(const typename Vec2<T>::type* tile, typename Vec2<T>::type* g, uint64_t base,
                                           const uint32_t* hi_off, int m, int L) {
  constexpr int EPV = 16 / sizeof(typename Vec2<T>::type);
  constexpr int LEPV = EPV == 2 ? 1 : 0;
  const int n_vec = (1 << m) >> LEPV;
  const int vpc_log = L - LEPV;
  for (int v = threadIdx.x; v < n_vec; v += blockDim.x) {
    int h = v >> vpc_log;
    int w = v & ((1 << vpc_log) - 1);
    uint64_t e = base + ((uint64_t)hi_off[h] << L) + ((uint64_t)w << LEPV);
    int4 val = *reinterpret_cast<const int4*>(tile + ((h << L) + (w << LEPV)));
    __stcs(reinterpret_cast<int4*>(g + e), unit_fix<T>(val));
  }
}

// ---------------------------------------------------------------------------------------------------
// shared-memory layout of a sweep CTA (dynamic):
//   [tile psi: 2^m T2][tile lam: 2^m T2 (backward)][smats: n_ops*8 T][acc: n_kslots*kAcc T (bwd)]
//   [wpart: 2*kMaxWarps*kAcc T (bwd)][hi_off: 2^(m-L) u32][ops: n_ops KOp]
__host__ __device__ inline size_t sweep_smem_bytes(int m, int L, int n_ops, int n_kslots, bool backward, size_t szT) {
  size_t b = (size_t(1) << m) * 2 * szT * (backward ? 2 : 1);
  b += size_t(n_ops) * 8 * szT;
  if (backward) b += size_t(n_kslots) * 4 * szT + size_t(2 * kMaxWarps * 4) * szT;  // kAcc = 4
  b = (b + 15) & ~size_t(15);
  b += (size_t(1) << (m - L)) * 4;
  b = (b + 15) & ~size_t(15);
  b += size_t(n_ops) * sizeof(KOp);
  return b;
}

template <typename T>
__device__ __forceinline__ void sweep_setup(const SweepArgs& A, int b, T* smats, uint32_t* hi_off, KOp* sops) {
  for (int i = threadIdx.x; i < A.n_ops; i += blockDim.x) sops[i] = A.ops[i];
  for (int i = threadIdx.x; i < A.n_ops * 8; i += blockDim.x) {
    int op = i >> 3, j = i & 7;
    int mat = A.ops[op].mat;
    T v = 0;
    if (mat >= 0) {
      int idx = mat >> 1;
      v = (mat & 1) ? reinterpret_cast<const T*>(A.mats_batch)[((size_t)b * A.n_groups_batch + idx) * 8 + j]
                    : reinterpret_cast<const T*>(A.mats_shared)[(size_t)idx * 8 + j];
    }
    smats[i] = v;
  }
  const int nh = 1 << (A.m - A.L);
  for (int h = threadIdx.x; h < nh; h += blockDim.x) {
    uint64_t off = 0;
    for (int k = 0; k < A.m - A.L; ++k) off |= (uint64_t)((h >> k) & 1) << A.tile_bits[A.L + k];
    hi_off[h] = (uint32_t)(off >> A.L);
  }
}

__device__ __forceinline__ uint64_t tile_base(const SweepArgs& A, uint32_t tau) {
  uint64_t base = 0;
  const int nn = A.n_local - A.m;
  for (int k = 0; k < nn; ++k) base |= (uint64_t)((tau >> k) & 1) << A.nontile_bits[k];
  return base;
}

// ---------------------------------------------------------------------------------------------------
// forward sweep: psi_tile <- G_last ... G_first psi_tile
template <typename T>
__global__ void __launch_bounds__(kSweepThreads) sweep_forward_kernel(const __grid_constant__ SweepArgs A) {
  using T2 = typename Vec2<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = A.m, L = A.L;
  T2* tile = reinterpret_cast<T2*>(smem_raw);
  T* smats = reinterpret_cast<T*>(tile + (size_t(1) << m));
  size_t off = ((size_t(1) << m) * sizeof(T2) + size_t(A.n_ops) * 8 * sizeof(T) + 15) & ~size_t(15);
  uint32_t* hi_off = reinterpret_cast<uint32_t*>(smem_raw + off);
  off = (off + (size_t(1) << (m - L)) * 4 + 15) & ~size_t(15);
  KOp* sops = reinterpret_cast<KOp*>(smem_raw + off);

  const int b = blockIdx.x / A.cps;
  const int c = blockIdx.x % A.cps;
  sweep_setup<T>(A, b, smats, hi_off, sops);
  __syncthreads();

  T2* gstate = reinterpret_cast<T2*>(A.psi) + ((uint64_t)b << A.n_local);
  const uint32_t n_tiles = 1u << (A.n_local - m);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t half = 1u << (m - 1), quarter = m >= 2 ? (1u << (m - 2)) : 0u, full = 1u << m;

  for (uint32_t tau = c; tau < n_tiles; tau += A.cps) {
    const uint64_t base = tile_base(A, tau);
    const uint64_t gbase = base | A.rank_bits;
    tile_load<T>(tile, gstate, base, hi_off, m, L);
    __syncthreads();
    for (int oi = 0; oi < A.n_ops; ++oi) {
      const KOp op = sops[oi];
      const T* M = smats + oi * 8;
      switch (op.kind) {
        case K_U1: {
          const T2 u00 = {M[0], M[1]}, u01 = {M[2], M[3]}, u10 = {M[4], M[5]}, u11 = {M[6], M[7]};
          const int a = op.a;
          for (uint32_t k = tid; k < half; k += nthr) {
            uint32_t i0 = ins0(k, a), i1 = i0 | (1u << a);
            T2 x = tile[i0], y = tile[i1];
            tile[i0] = cfma(u01, y, cmul(u00, x));
            tile[i1] = cfma(u11, y, cmul(u10, x));
          }
          break;
        }
        case K_D1: {
          const T2 d0 = {M[0], M[1]}, d1 = {M[6], M[7]};
          const int a = op.a;
          for (uint32_t i = tid; i < full; i += nthr) tile[i] = cmul(tile[i], ((i >> a) & 1u) ? d1 : d0);
          break;
        }
        case K_D1_EXT: {
          const T2 d = ((gbase >> op.ext_bit) & 1ull) ? T2{M[6], M[7]} : T2{M[0], M[1]};
          for (uint32_t i = tid; i < full; i += nthr) tile[i] = cmul(tile[i], d);
          break;
        }
        case K_CX: {
          const int lo = min(op.a, op.c), hi = max(op.a, op.c);
          for (uint32_t k = tid; k < quarter; k += nthr) {
            uint32_t i0 = ins0(ins0(k, lo), hi) | (1u << op.c), i1 = i0 | (1u << op.a);
            T2 x = tile[i0];
            tile[i0] = tile[i1];
            tile[i1] = x;
          }
          break;
        }
        case K_CX_EXT: {
          if ((gbase & op.ext_mask) == op.ext_mask) {
            for (uint32_t k = tid; k < half; k += nthr) {
              uint32_t i0 = ins0(k, op.a), i1 = i0 | (1u << op.a);
              T2 x = tile[i0];
              tile[i0] = tile[i1];
              tile[i1] = x;
            }
          }
          break;
        }
        case K_CZ: {
          const int lo = min(op.a, op.c), hi = max(op.a, op.c);
          for (uint32_t k = tid; k < quarter; k += nthr) {
            uint32_t i = ins0(ins0(k, lo), hi) | (1u << op.a) | (1u << op.c);
            T2 x = tile[i];
            x.x = -x.x;
            x.y = -x.y;
            tile[i] = x;
          }
          break;
        }
        case K_CZ_EXT1: {
          if ((gbase & op.ext_mask) == op.ext_mask) {
            for (uint32_t k = tid; k < half; k += nthr) {
              uint32_t i = ins0(k, op.a) | (1u << op.a);
              T2 x = tile[i];
              x.x = -x.x;
              x.y = -x.y;
              tile[i] = x;
            }
          }
          break;
        }
        case K_CZ_EXT2: {
          if ((gbase & op.ext_mask) == op.ext_mask) {
            for (uint32_t i = tid; i < full; i += nthr) {
              T2 x = tile[i];
              x.x = -x.x;
              x.y = -x.y;
              tile[i] = x;
            }
          }
          break;
        }
        case K_SWAP: {
          const int lo = min(op.a, op.c), hi = max(op.a, op.c);
          for (uint32_t k = tid; k < quarter; k += nthr) {
            uint32_t i = ins0(ins0(k, lo), hi);
            uint32_t i0 = i | (1u << op.a), i1 = i | (1u << op.c);
            T2 x = tile[i0];
            tile[i0] = tile[i1];
            tile[i1] = x;
          }
          break;
        }
        default:
          break;
      }
      __syncthreads();
    }
    tile_store<T>(tile, gstate, base, hi_off, m, L);
    __syncthreads();
  }
}

// ---------------------------------------------------------------------------------------------------
// gradient accumulators: per parametrised fused group the backward accumulates the Pauli vector
//   s_P = sum over pairs Im <lam| sigma_P |psi>,  P in {X, Y, Z}      (kAcc = 4 values, the 4th is padding)
// of the states right after the group; dtheta_k = a_k . s with a_k the Pauli vector of V_k P_k V_k^+
// (finalize_grads_kernel).  Per pair (x, y) = psi at bit 0 / 1, (lx, ly) = lambda:
//   sX += Im(conj(lx) y + conj(ly) x),  sY += Re(conj(ly) x) - Re(conj(lx) y),  sZ += Im(conj(lx) x - conj(ly) y)
constexpr int kAcc = 4;

template <typename T2, typename T>
__device__ __forceinline__ void pauli_acc(T& sx, T& sy, T& sz, T2 x, T2 y, T2 lx, T2 ly) {
  sx = fma(lx.x, y.y, fma(-lx.y, y.x, fma(ly.x, x.y, fma(-ly.y, x.x, sx))));
  sy = fma(ly.x, x.x, fma(ly.y, x.y, fma(-lx.x, y.x, fma(-lx.y, y.y, sy))));
  sz = fma(lx.x, x.y, fma(-lx.y, x.x, fma(-ly.x, y.y, fma(ly.y, y.x, sz))));
}
// Im(conj(l) p)
template <typename T2> __device__ __forceinline__ auto im_conj_mul(T2 l, T2 p) { return l.x * p.y - l.y * p.x; }

// ---------------------------------------------------------------------------------------------------
// generic backward sweep (adjoint-state method): ops in reverse; for each op G:
//   accumulate s_P (parametrised groups only);  psi <- G^+ psi,  lam <- G^+ lam
template <typename T>
__device__ __forceinline__ void block_accumulate3(T a0, T a1, T a2, T* wpart, T* acc_slot, int parity) {
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* wp = wpart + parity * (kMaxWarps * kAcc);
  if (lane == 0) {
    wp[warp * kAcc + 0] = a0;
    wp[warp * kAcc + 1] = a1;
    wp[warp * kAcc + 2] = a2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    T s = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) s += wp[w * kAcc + threadIdx.x];
    acc_slot[threadIdx.x] += s;
  }
}

template <typename T>
__global__ void __launch_bounds__(kSweepThreads) sweep_backward_kernel(const __grid_constant__ SweepArgs A) {
  using T2 = typename Vec2<T>::type;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = A.m, L = A.L;
  T2* tp = reinterpret_cast<T2*>(smem_raw);
  T2* tl = tp + (size_t(1) << m);
  T* smats = reinterpret_cast<T*>(tl + (size_t(1) << m));
  T* acc = smats + size_t(A.n_ops) * 8;
  T* wpart = acc + size_t(A.n_kslots) * kAcc;
  size_t off = ((size_t(2) << m) * sizeof(T2) +
                (size_t(A.n_ops) * 8 + size_t(A.n_kslots) * kAcc + size_t(2 * kMaxWarps * kAcc)) * sizeof(T) + 15) &
               ~size_t(15);
  uint32_t* hi_off = reinterpret_cast<uint32_t*>(smem_raw + off);
  off = (off + (size_t(1) << (m - L)) * 4 + 15) & ~size_t(15);
  KOp* sops = reinterpret_cast<KOp*>(smem_raw + off);

  const int b = blockIdx.x / A.cps;
  const int c = blockIdx.x % A.cps;
  sweep_setup<T>(A, b, smats, hi_off, sops);
  for (int i = threadIdx.x; i < A.n_kslots * kAcc; i += blockDim.x) acc[i] = 0;
  __syncthreads();

  T2* gpsi = reinterpret_cast<T2*>(A.psi) + ((uint64_t)b << A.n_local);
  T2* glam = reinterpret_cast<T2*>(A.lam) + ((uint64_t)b << A.n_local);
  const uint32_t n_tiles = 1u << (A.n_local - m);
  const int tid = threadIdx.x, nthr = blockDim.x;
  const uint32_t half = 1u << (m - 1), quarter = m >= 2 ? (1u << (m - 2)) : 0u, full = 1u << m;
  int parity = 0;

  for (uint32_t tau = c; tau < n_tiles; tau += A.cps) {
    const uint64_t base = tile_base(A, tau);
    const uint64_t gbase = base | A.rank_bits;
    tile_load<T>(tp, gpsi, base, hi_off, m, L);
    tile_load<T>(tl, glam, base, hi_off, m, L);
    __syncthreads();
    // Im<lam|psi> over the tile is invariant under the in-tile unitaries: computed once for all K_D1_EXT grads
    T tdot = 0;
    if (A.need_tile_dot) {
      T s = 0;
      for (uint32_t i = tid; i < full; i += nthr) s += im_conj_mul(tl[i], tp[i]);
      s = warp_sum(s);
      T* wp = wpart + parity * (kMaxWarps * kAcc);
      if ((tid & 31) == 0) wp[(tid >> 5) * kAcc] = s;
      __syncthreads();
      const int nw = nthr >> 5;
      for (int w = 0; w < nw; ++w) tdot += wp[w * kAcc];
      parity ^= 1;
    }
    for (int oi = A.n_ops - 1; oi >= 0; --oi) {
      const KOp op = sops[oi];
      const T* M = smats + oi * 8;
      bool synced = false;
      switch (op.kind) {
        case K_U1: {
          // adjoint: rows/cols swapped + conjugate
          const T2 a00 = {M[0], -M[1]}, a01 = {M[4], -M[5]}, a10 = {M[2], -M[3]}, a11 = {M[6], -M[7]};
          const int a = op.a;
          T sx = 0, sy = 0, sz = 0;
          for (uint32_t k = tid; k < half; k += nthr) {
            uint32_t i0 = ins0(k, a), i1 = i0 | (1u << a);
            T2 x = tp[i0], y = tp[i1], lx = tl[i0], ly = tl[i1];
            pauli_acc(sx, sy, sz, x, y, lx, ly);
            tp[i0] = cfma(a01, y, cmul(a00, x));
            tp[i1] = cfma(a11, y, cmul(a10, x));
            tl[i0] = cfma(a01, ly, cmul(a00, lx));
            tl[i1] = cfma(a11, ly, cmul(a10, lx));
          }
          if (op.kslot >= 0) {
            block_accumulate3<T>(sx, sy, sz, wpart, acc + op.kslot * kAcc, parity);
            parity ^= 1;
            synced = true;
          }
          break;
        }
        case K_D1: {
          const T2 d0 = {M[0], -M[1]}, d1 = {M[6], -M[7]};
          const int a = op.a;
          T sz = 0;
          for (uint32_t i = tid; i < full; i += nthr) {
            T2 x = tp[i], lx = tl[i];
            T im = im_conj_mul(lx, x);
            if ((i >> a) & 1u) {
              sz -= im;
              tp[i] = cmul(x, d1);
              tl[i] = cmul(lx, d1);
            } else {
              sz += im;
              tp[i] = cmul(x, d0);
              tl[i] = cmul(lx, d0);
            }
          }
          if (op.kslot >= 0) {
            block_accumulate3<T>((T)0, (T)0, sz, wpart, acc + op.kslot * kAcc, parity);
            parity ^= 1;
            synced = true;
          }
          break;
        }
        case K_D1_EXT: {
          const bool one = (gbase >> op.ext_bit) & 1ull;
          const T2 d = one ? T2{M[6], -M[7]} : T2{M[0], -M[1]};
          for (uint32_t i = tid; i < full; i += nthr) {
            tp[i] = cmul(tp[i], d);
            tl[i] = cmul(tl[i], d);
          }
          if (op.kslot >= 0 && tid == 3) acc[op.kslot * kAcc + 2] += one ? -tdot : tdot;
          break;
        }
        case K_CX: {
          const int lo = min(op.a, op.c), hi = max(op.a, op.c);
          for (uint32_t k = tid; k < quarter; k += nthr) {
            uint32_t i0 = ins0(ins0(k, lo), hi) | (1u << op.c), i1 = i0 | (1u << op.a);
            T2 x = tp[i0];
            tp[i0] = tp[i1];
            tp[i1] = x;
            x = tl[i0];
            tl[i0] = tl[i1];
            tl[i1] = x;
          }
          break;
        }
        case K_CX_EXT: {
          if ((gbase & op.ext_mask) == op.ext_mask) {
            for (uint32_t k = tid; k < half; k += nthr) {
              uint32_t i0 = ins0(k, op.a), i1 = i0 | (1u << op.a);
              T2 x = tp[i0];
              tp[i0] = tp[i1];
              tp[i1] = x;
              x = tl[i0];
              tl[i0] = tl[i1];
              tl[i1] = x;
            }
          }
          break;
        }
        case K_CZ: {
          const int lo = min(op.a, op.c), hi = max(op.a, op.c);
          for (uint32_t k = tid; k < quarter; k += nthr) {
            uint32_t i = ins0(ins0(k, lo), hi) | (1u << op.a) | (1u << op.c);
            T2 x = tp[i];
            tp[i] = T2{-x.x, -x.y};
            x = tl[i];
            tl[i] = T2{-x.x, -x.y};
          }
          break;
        }
        case K_CZ_EXT1: {
          if ((gbase & op.ext_mask) == op.ext_mask) {
            for (uint32_t k = tid; k < half; k += nthr) {
              uint32_t i = ins0(k, op.a) | (1u << op.a);
              T2 x = tp[i];
              tp[i] = T2{-x.x, -x.y};
              x = tl[i];
              tl[i] = T2{-x.x, -x.y};
            }
          }
          break;
        }
        case K_CZ_EXT2: {
          if ((gbase & op.ext_mask) == op.ext_mask) {
            for (uint32_t i = tid; i < full; i += nthr) {
              T2 x = tp[i];
              tp[i] = T2{-x.x, -x.y};
              x = tl[i];
              tl[i] = T2{-x.x, -x.y};
            }
          }
          break;
        }
        case K_SWAP: {
          const int lo = min(op.a, op.c), hi = max(op.a, op.c);
          for (uint32_t k = tid; k < quarter; k += nthr) {
            uint32_t i = ins0(ins0(k, lo), hi);
            uint32_t i0 = i | (1u << op.a), i1 = i | (1u << op.c);
            T2 x = tp[i0];
            tp[i0] = tp[i1];
            tp[i1] = x;
            x = tl[i0];
            tl[i0] = tl[i1];
            tl[i1] = x;
          }
          break;
        }
        default:
          break;
      }
      if (!synced) __syncthreads();
    }
    tile_store<T>(tp, gpsi, base, hi_off, m, L);
    tile_store<T>(tl, glam, base, hi_off, m, L);
    __syncthreads();
  }
  // flush this CTA's accumulators (plain stores: deterministic)
  T* out = reinterpret_cast<T*>(A.partials) + (size_t)blockIdx.x * A.n_kslots * kAcc;
  for (int i = threadIdx.x; i < A.n_kslots * kAcc; i += blockDim.x) out[i] = acc[i];
}

// ===================================================================================================
// Register-blocked ("staged") sweeps.  plan.h: Stage.  Each thread owns the R = 2^RB amplitudes that differ in the
// stage's register bits, applies the stage's ops in registers and writes them back: one shared-memory round
// trip per stage.  The tile is stored swizzled by 16-byte pieces (piece p lives at p ^ ((p >> 3) & 7)) so that
// both access shapes are bank-conflict free: a `low` stage reads 8 consecutive pieces per thread (128-bit
// accesses, thread stride 128 B), any other stage reads single amplitudes with consecutive threads on
// consecutive amplitudes.
template <typename T> struct StageCfg;
template <> struct StageCfg<float> { static constexpr int RB = 4, LE = 1; };   // 16 amplitudes / thread, 2 per piece
template <> struct StageCfg<double> { static constexpr int RB = 3, LE = 0; };  // 8 amplitudes / thread, 1 per piece

__device__ __forceinline__ uint32_t swz_piece(uint32_t p) { return p ^ ((p >> 3) & 7u); }
template <int LE> __device__ __forceinline__ uint32_t swz_amp(uint32_t i) {
  return (swz_piece(i >> LE) << LE) | (i & ((1u << LE) - 1u));
}

template <typename T>
__device__ __forceinline__ void tile_load_swz(typename Vec2<T>::type* tile, const typename Vec2<T>::type* g, uint64_t base,
                                              const uint32_t* hi_off, int m, int L) {
  constexpr int LE = StageCfg<T>::LE;
  constexpr int UN = 8;
  const int n_vec = (1 << m) >> LE;
  const int vpc_log = L - LE;
  int4* tp = reinterpret_cast<int4*>(tile);
  for (int v0 = 0; v0 < n_vec; v0 += blockDim.x * UN) {
    int4 buf[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int v = v0 + u * blockDim.x + threadIdx.x;
      if (v < n_vec) {
        const int h = v >> vpc_log, w = v & ((1 << vpc_log) - 1);
        buf[u] = __ldcs(reinterpret_cast<const int4*>(g + base + ((uint64_t)hi_off[h] << L) + ((uint64_t)w << LE)));
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const int v = v0 + u * blockDim.x + threadIdx.x;
      if (v < n_vec) tp[swz_piece(v)] = unit_fix<T>(buf[u]);
    }
  }
}

template <typename T>
__device__ __forceinline__ void tile_store_swz(const typename Vec2<T>::type* tile, typename Vec2<T>::type* g, uint64_t base,
                                               const uint32_t* hi_off, int m, int L) {
  constexpr int LE = StageCfg<T>::LE;
  const int n_vec = (1 << m) >> LE;
  const int vpc_log = L - LE;
  const int4* tp = reinterpret_cast<const int4*>(tile);
  for (int v = threadIdx.x; v < n_vec; v += blockDim.x) {
    int h = v >> vpc_log;
    int w = v & ((1 << vpc_log) - 1);
    uint64_t e = base + ((uint64_t)hi_off[h] << L) + ((uint64_t)w << LE);
    __stcs(reinterpret_cast<int4*>(g + e), unit_fix<T>(tp[swz_piece(v)]));
  }
}

// ---- register-array primitives (all register indices are compile-time) -----------------------------------
template <int RBIT, int R, typename T2>
__device__ __forceinline__ void reg_u1(T2 (&v)[R], T2 u00, T2 u01, T2 u10, T2 u11) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    if (j & (1 << RBIT)) continue;
    const T2 x = v[j], y = v[j | (1 << RBIT)];
    v[j] = cfma(u01, y, cmul(u00, x));
    v[j | (1 << RBIT)] = cfma(u11, y, cmul(u10, x));
  }
}

template <int RBIT, int R, typename T2, typename T>
__device__ __forceinline__ void reg_pauli(const T2 (&v)[R], const T2 (&l)[R], T& sx, T& sy, T& sz) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    if (j & (1 << RBIT)) continue;
    pauli_acc(sx, sy, sz, v[j], v[j | (1 << RBIT)], l[j], l[j | (1 << RBIT)]);
  }
}

// swap the pair (j, j | 1<<RBIT) where the control predicate holds: thread_ok && (j & cmask) == cmask
template <int RBIT, int R, typename T2>
__device__ __forceinline__ void reg_cx(T2 (&v)[R], bool thread_ok, uint32_t cmask) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    if (j & (1 << RBIT)) continue;
    const bool p = thread_ok && ((j & cmask) == cmask);
    const T2 x = v[j], y = v[j | (1 << RBIT)];
    v[j].x = p ? y.x : x.x;
    v[j].y = p ? y.y : x.y;
    v[j | (1 << RBIT)].x = p ? x.x : y.x;
    v[j | (1 << RBIT)].y = p ? x.y : y.y;
  }
}

template <int R, typename T2>
__device__ __forceinline__ void reg_negate_where(T2 (&v)[R], bool thread_ok, uint32_t jmask) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const bool p = thread_ok && ((j & jmask) == jmask);
    v[j].x = p ? -v[j].x : v[j].x;
    v[j].y = p ? -v[j].y : v[j].y;
  }
}

// v[j] *= (bit set ? d1 : d0), bit = register bit r (r >= 0) or a per-thread bit
template <int R, typename T2>
__device__ __forceinline__ void reg_diag(T2 (&v)[R], int r, bool thread_bit, T2 d0, T2 d1) {
#pragma unroll
  for (int j = 0; j < R; ++j) {
    const bool one = r >= 0 ? ((j >> r) & 1) : thread_bit;
    const T2 d = {one ? d1.x : d0.x, one ? d1.y : d0.y};
    v[j] = cmul(v[j], d);
  }
}

#define QB_DISPATCH_RBIT(RBv, r, CALL)            \
  do {                                            \
    if constexpr (RBv == 4) {                     \
      switch (r) {                                \
        case 0: { constexpr int RBIT = 0; CALL; } break; \
        case 1: { constexpr int RBIT = 1; CALL; } break; \
        case 2: { constexpr int RBIT = 2; CALL; } break; \
        default: { constexpr int RBIT = 3; CALL; } break; \
      }                                           \
    } else {                                      \
      switch (r) {                                \
        case 0: { constexpr int RBIT = 0; CALL; } break; \
        case 1: { constexpr int RBIT = 1; CALL; } break; \
        default: { constexpr int RBIT = 2; CALL; } break; \
      }                                           \
    }                                             \
  } while (0)

// per-warp gradient accumulators: lane 0 of each warp owns wacc[warp][kslot][0..2] (no atomics, no block barrier)
template <typename T>
__device__ __forceinline__ void warp_accumulate3(T a0, T a1, T a2, T* wacc_slot) {
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  if ((threadIdx.x & 31) == 0) {
    wacc_slot[0] += a0;
    wacc_slot[1] += a1;
    wacc_slot[2] += a2;
  }
}
template <typename T> __device__ __forceinline__ void warp_accumulate1(T a2, T* wacc_slot) {
  a2 = warp_sum(a2);
  if ((threadIdx.x & 31) == 0) wacc_slot[2] += a2;
}

// Apply ops [ob, oe) of a stage to the register arrays.  BWD: reverse order, adjoint matrices, on psi (v) and
// lambda (l), accumulating the Pauli vectors.  i_base: the thread's local index with the register bits cleared.
template <typename T, bool BWD>
__device__ __forceinline__ void stage_apply(typename Vec2<T>::type (&v)[1 << StageCfg<T>::RB],
                                            typename Vec2<T>::type (&l)[1 << StageCfg<T>::RB], const KOp* sops,
                                            const T* smats, int ob, int oe, uint32_t i_base, uint64_t gbase, T* wacc,
                                            T tdot) {
  using T2 = typename Vec2<T>::type;
  constexpr int RB = StageCfg<T>::RB;
  constexpr int R = 1 << RB;
  const int n = oe - ob;
  for (int q = 0; q < n; ++q) {
    const int oi = BWD ? (oe - 1 - q) : (ob + q);
    const KOp op = sops[oi];
    const T* M = smats + oi * 8;
    switch (op.kind) {
      case K_U1: {
        T2 u00, u01, u10, u11;
        if (BWD) {
          u00 = {M[0], -M[1]};
          u01 = {M[4], -M[5]};
          u10 = {M[2], -M[3]};
          u11 = {M[6], -M[7]};
          if (op.kslot >= 0) {
            T sx = 0, sy = 0, sz = 0;
            QB_DISPATCH_RBIT(RB, op.r, (reg_pauli<RBIT, R>(v, l, sx, sy, sz)));
            warp_accumulate3<T>(sx, sy, sz, wacc + op.kslot * kAcc);
          }
        } else {
          u00 = {M[0], M[1]};
          u01 = {M[2], M[3]};
          u10 = {M[4], M[5]};
          u11 = {M[6], M[7]};
        }
        QB_DISPATCH_RBIT(RB, op.r, (reg_u1<RBIT, R>(v, u00, u01, u10, u11)));
        if (BWD) QB_DISPATCH_RBIT(RB, op.r, (reg_u1<RBIT, R>(l, u00, u01, u10, u11)));
        break;
      }
      case K_D1:
      case K_D1_EXT: {
        const T sgn = BWD ? (T)-1 : (T)1;
        const T2 d0 = {M[0], sgn * M[1]}, d1 = {M[6], sgn * M[7]};
        int r = -1;
        bool tb;
        if (op.kind == K_D1) {
          r = op.r;
          tb = (i_base >> op.a) & 1u;
        } else {
          tb = (gbase >> op.ext_bit) & 1ull;
        }
        if (BWD && op.kslot >= 0) {
          if (op.kind == K_D1) {
            T sz = 0;
#pragma unroll
            for (int j = 0; j < R; ++j) {
              const bool one = r >= 0 ? ((j >> r) & 1) : tb;
              const T im = im_conj_mul(l[j], v[j]);
              sz += one ? -im : im;
            }
            warp_accumulate1<T>(sz, wacc + op.kslot * kAcc);
          } else if (threadIdx.x == 0 && i_base == 0) {
            // tile-level Im<lam|psi> (computed once per tile); thread 0's first group only
            wacc[op.kslot * kAcc + 2] += tb ? -tdot : tdot;
          }
        }
        reg_diag<R>(v, r, tb, d0, d1);
        if (BWD) reg_diag<R>(l, r, tb, d0, d1);
        break;
      }
      case K_CX:
      case K_CX_EXT: {
        bool ok;
        uint32_t cmask = 0;
        if (op.kind == K_CX) {
          if (op.rc >= 0) {
            ok = true;
            cmask = 1u << op.rc;
          } else {
            ok = (i_base >> op.c) & 1u;
          }
        } else {
          ok = (gbase & op.ext_mask) == op.ext_mask;
        }
        QB_DISPATCH_RBIT(RB, op.r, (reg_cx<RBIT, R>(v, ok, cmask)));
        if (BWD) QB_DISPATCH_RBIT(RB, op.r, (reg_cx<RBIT, R>(l, ok, cmask)));
        break;
      }
      case K_CZ:
      case K_CZ_EXT1:
      case K_CZ_EXT2: {
        // merged run of sign flips (plan.cpp marks runs through ext_bit): one R-bit mask for the whole run
        const int run = op.ext_bit > 0 ? op.ext_bit : 1;
        uint32_t Msk = 0;
        for (int u = 0; u < run; ++u) {
          const KOp o2 = sops[BWD ? (oi - u) : (oi + u)];
          uint32_t ok = 1u, ma = 0xFFFFu, mc = 0xFFFFu;
          if (o2.kind != K_CZ) ok = ((gbase & o2.ext_mask) == o2.ext_mask) ? 1u : 0u;
          if (o2.kind != K_CZ_EXT2) {
            if (o2.r >= 0)
              ma = (uint32_t)(0xFF00F0F0CCCCAAAAull >> (16 * o2.r)) & 0xFFFFu;
            else
              ok &= (i_base >> o2.a) & 1u;
          }
          if (o2.kind == K_CZ) {
            if (o2.rc >= 0)
              mc = (uint32_t)(0xFF00F0F0CCCCAAAAull >> (16 * o2.rc)) & 0xFFFFu;
            else
              ok &= (i_base >> o2.c) & 1u;
          }
          Msk ^= ok ? (ma & mc) : 0u;
        }
        q += run - 1;
#pragma unroll
        for (int j = 0; j < R; ++j) {
          const bool neg = (Msk >> j) & 1u;
          v[j].x = neg ? -v[j].x : v[j].x;
          v[j].y = neg ? -v[j].y : v[j].y;
          if (BWD) {
            l[j].x = neg ? -l[j].x : l[j].x;
            l[j].y = neg ? -l[j].y : l[j].y;
          }
        }
        break;
      }
      default:
        break;
    }
  }
}

struct StagedArgs {
  SweepArgs s;
  const Stage* stages;
  int32_t n_stages;
};

__host__ __device__ inline size_t staged_smem_bytes(int m, int L, int n_ops, int n_kslots, int n_stages, bool backward,
                                                    size_t szT) {
  size_t b = (size_t(1) << m) * 2 * szT * (backward ? 2 : 1);
  b += size_t(n_ops) * 8 * szT;
  if (backward) b += size_t(kMaxWarps) * n_kslots * kAcc * szT + size_t(kMaxWarps) * szT;
  b = (b + 15) & ~size_t(15);
  b += (size_t(1) << (m - L)) * 4;
  b = (b + 15) & ~size_t(15);
  b += size_t(n_ops) * sizeof(KOp);
  b = (b + 15) & ~size_t(15);
  b += size_t(n_stages) * sizeof(Stage);
  b = (b + 15) & ~size_t(15);
  b += size_t(n_stages) * 16 * sizeof(uint16_t);  // per-stage swizzled byte offsets of the register amplitudes
  return b;
}

template <typename T, bool BWD>
__global__ void __launch_bounds__(kSweepThreads, BWD ? 2 : 3) sweep_staged_kernel(const __grid_constant__ StagedArgs SA) {
  using T2 = typename Vec2<T>::type;
  constexpr int RB = StageCfg<T>::RB, LE = StageCfg<T>::LE;
  constexpr int R = 1 << RB;
  constexpr int NP = R >> LE;  // 16-byte pieces per thread in a low stage (8)
  const SweepArgs& A = SA.s;
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int m = A.m, L = A.L;
  T2* tp = reinterpret_cast<T2*>(smem_raw);
  T2* tl = tp + (BWD ? (size_t(1) << m) : 0);
  T* smats = reinterpret_cast<T*>(tl + (size_t(1) << m));
  T* wacc_all = smats + size_t(A.n_ops) * 8;
  T* wred = wacc_all + (BWD ? size_t(kMaxWarps) * A.n_kslots * kAcc : 0);
  size_t off = (size_t(1) << m) * sizeof(T2) * (BWD ? 2 : 1) + size_t(A.n_ops) * 8 * sizeof(T);
  if (BWD) off += (size_t(kMaxWarps) * A.n_kslots * kAcc + size_t(kMaxWarps)) * sizeof(T);
  off = (off + 15) & ~size_t(15);
  uint32_t* hi_off = reinterpret_cast<uint32_t*>(smem_raw + off);
  off = (off + (size_t(1) << (m - L)) * 4 + 15) & ~size_t(15);
  KOp* sops = reinterpret_cast<KOp*>(smem_raw + off);
  off = (off + size_t(A.n_ops) * sizeof(KOp) + 15) & ~size_t(15);
  Stage* sst = reinterpret_cast<Stage*>(smem_raw + off);
  off = (off + size_t(SA.n_stages) * sizeof(Stage) + 15) & ~size_t(15);
  uint16_t* soff = reinterpret_cast<uint16_t*>(smem_raw + off);  // [n_stages][16]

  const int b = blockIdx.x / A.cps;
  const int c = blockIdx.x % A.cps;
  const int tid = threadIdx.x, nthr = blockDim.x;
  sweep_setup<T>(A, b, smats, hi_off, sops);
  for (int i = tid; i < SA.n_stages; i += nthr) sst[i] = SA.stages[i];
  // swz_amp is linear over GF(2) and i_base / the register-bit pattern are bit-disjoint, so the byte offset of
  // register amplitude j is  swz(i_base)*sizeof(T2)  XOR  soff[stage][j]
  for (int i = tid; i < SA.n_stages * 16; i += nthr) {
    const Stage& stg = SA.stages[i >> 4];
    const int j = i & 15;
    uint32_t jm = 0;
    for (int k = 0; k < RB; ++k)
      if ((j >> k) & 1) jm |= 1u << stg.regbits[k];
    soff[i] = (j < R) ? (uint16_t)(swz_amp<LE>(jm) * sizeof(T2)) : (uint16_t)0;
  }
  if (BWD)
    for (int i = tid; i < kMaxWarps * A.n_kslots * kAcc; i += nthr) wacc_all[i] = 0;
  __syncthreads();
  T* wacc = wacc_all + (BWD ? size_t(tid >> 5) * A.n_kslots * kAcc : 0);

  T2* gpsi = reinterpret_cast<T2*>(A.psi) + ((uint64_t)b << A.n_local);
  T2* glam = BWD ? reinterpret_cast<T2*>(A.lam) + ((uint64_t)b << A.n_local) : nullptr;
  const uint32_t n_tiles = 1u << (A.n_local - m);
  const uint32_t n_groups = 1u << (m - RB);

  for (uint32_t tau = c; tau < n_tiles; tau += A.cps) {
    const uint64_t base = tile_base(A, tau);
    const uint64_t gbase = base | A.rank_bits;
    tile_load_swz<T>(tp, gpsi, base, hi_off, m, L);
    if (BWD) tile_load_swz<T>(tl, glam, base, hi_off, m, L);
    __syncthreads();
    T tdot = 0;
    if (BWD && A.need_tile_dot) {
      T s = 0;
      for (uint32_t i = tid; i < (1u << m); i += nthr) s += im_conj_mul(tl[i], tp[i]);  // same swizzle on both tiles
      s = warp_sum(s);
      if ((tid & 31) == 0) wred[tid >> 5] = s;
      __syncthreads();
      for (int w = 0; w < (nthr >> 5); ++w) tdot += wred[w];
    }
    for (int sq = 0; sq < SA.n_stages; ++sq) {
      const int si = BWD ? (SA.n_stages - 1 - sq) : sq;
      const Stage st = sst[si];
      // warp-uniform trip count: stage_apply contains full-warp shuffles
      for (uint32_t g0 = 0; g0 < n_groups; g0 += nthr) {
        const uint32_t g = g0 + tid;
        const bool active = g < n_groups;
        T2 v[R], l[R];
        uint32_t i_base = 0, sb = 0;
        uint32_t offs[R / 2];  // packed uint16 byte offsets
        if (!active) {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            v[j] = T2{0, 0};
            l[j] = T2{0, 0};
          }
        } else if (st.low) {
          i_base = g << RB;
          const int4* pp = reinterpret_cast<const int4*>(tp) + (size_t)g * NP;
          int4* vv = reinterpret_cast<int4*>(v);
#pragma unroll
          for (int k = 0; k < NP; ++k) vv[k] = pp[k ^ (g & 7u)];
          if (BWD) {
            const int4* pl = reinterpret_cast<const int4*>(tl) + (size_t)g * NP;
            int4* lv = reinterpret_cast<int4*>(l);
#pragma unroll
            for (int k = 0; k < NP; ++k) lv[k] = pl[k ^ (g & 7u)];
          }
        } else {
          i_base = g;
#pragma unroll
          for (int k = 0; k < RB; ++k) i_base = ins0(i_base, st.regbits[k]);
          sb = swz_amp<LE>(i_base) * (uint32_t)sizeof(T2);
          const uint4* tab = reinterpret_cast<const uint4*>(soff + si * 16);
#pragma unroll
          for (int q4 = 0; q4 < R / 8; ++q4) {
            const uint4 t4 = tab[q4];
            offs[q4 * 4 + 0] = t4.x;
            offs[q4 * 4 + 1] = t4.y;
            offs[q4 * 4 + 2] = t4.z;
            offs[q4 * 4 + 3] = t4.w;
          }
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const uint32_t o = sb ^ ((offs[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu);
            v[j] = *reinterpret_cast<const T2*>(reinterpret_cast<const char*>(tp) + o);
            if (BWD) l[j] = *reinterpret_cast<const T2*>(reinterpret_cast<const char*>(tl) + o);
          }
        }
        stage_apply<T, BWD>(v, l, sops, smats, st.op_begin, st.op_end, i_base, gbase, wacc, tdot);
        if (!active) continue;
        if (st.low) {
          int4* pp = reinterpret_cast<int4*>(tp) + (size_t)g * NP;
          const int4* vv = reinterpret_cast<const int4*>(v);
#pragma unroll
          for (int k = 0; k < NP; ++k) pp[k ^ (g & 7u)] = vv[k];
          if (BWD) {
            int4* pl = reinterpret_cast<int4*>(tl) + (size_t)g * NP;
            const int4* lv = reinterpret_cast<const int4*>(l);
#pragma unroll
            for (int k = 0; k < NP; ++k) pl[k ^ (g & 7u)] = lv[k];
          }
        } else {
#pragma unroll
          for (int j = 0; j < R; ++j) {
            const uint32_t o = sb ^ ((offs[j >> 1] >> ((j & 1) * 16)) & 0xFFFFu);
            *reinterpret_cast<T2*>(reinterpret_cast<char*>(tp) + o) = v[j];
            if (BWD) *reinterpret_cast<T2*>(reinterpret_cast<char*>(tl) + o) = l[j];
          }
        }
      }
      __syncthreads();
    }
    tile_store_swz<T>(tp, gpsi, base, hi_off, m, L);
    if (BWD) tile_store_swz<T>(tl, glam, base, hi_off, m, L);
    __syncthreads();
  }
  if (BWD) {
    T* out = reinterpret_cast<T*>(A.partials) + (size_t)blockIdx.x * A.n_kslots * kAcc;
    for (int i = tid; i < A.n_kslots * kAcc; i += nthr) {
      T s = 0;
      for (int w = 0; w < (nthr >> 5); ++w) s += wacc_all[(size_t)w * A.n_kslots * kAcc + i];
      out[i] = s;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// reduce the per-CTA partials of one backward sweep into the plan-wide accumulators (double sums, fixed order)
//   shared slot : K_shared[k_index][j]    = sum over all CTAs
//   batch slot  : K_batch[b][k_index][j]  = sum over the sample's cps CTAs
template <typename T>
__global__ void reduce_partials_kernel(const T* __restrict__ partials, const KSlot* __restrict__ kslots, int n_kslots,
                                       int B, int cps, T* __restrict__ k_shared, T* __restrict__ k_batch, int n_k_batch) {
  const int s = blockIdx.x;  // local slot
  const KSlot ks = kslots[s];
  __shared__ double red[kAcc][65];
  if (!ks.batch) {
    const int j = threadIdx.x & (kAcc - 1), lane = threadIdx.x / kAcc, nl = blockDim.x / kAcc;  // 256 threads -> 64 lanes
    double sum = 0;
    const long total = (long)B * cps;
    for (long r = lane; r < total; r += nl) sum += (double)partials[((size_t)r * n_kslots + s) * kAcc + j];
    red[j][lane] = sum;
    __syncthreads();
    if (threadIdx.x < kAcc) {
      double t = 0;
      for (int i = 0; i < nl; ++i) t += red[threadIdx.x][i];
      k_shared[(size_t)ks.k_index * kAcc + threadIdx.x] = (T)t;
    }
  } else {
    for (long idx = threadIdx.x; idx < (long)B * kAcc; idx += blockDim.x) {
      long bb = idx / kAcc;
      int j = idx % kAcc;
      double sum = 0;
      for (int cc = 0; cc < cps; ++cc) sum += (double)partials[(((size_t)bb * cps + cc) * n_kslots + s) * kAcc + j];
      k_batch[((size_t)bb * n_k_batch + ks.k_index) * kAcc + j] = (T)sum;
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// 2x2 helpers in double for the tiny prepare / finalize kernels
struct C2 {
  double x, y;
};
struct M22 {
  C2 m[4];
};
__device__ __forceinline__ C2 c2mul(C2 a, C2 b) { return {a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x}; }
__device__ __forceinline__ C2 c2add(C2 a, C2 b) { return {a.x + b.x, a.y + b.y}; }
__device__ __forceinline__ M22 m22mul(const M22& A, const M22& B) {
  M22 R;
  R.m[0] = c2add(c2mul(A.m[0], B.m[0]), c2mul(A.m[1], B.m[2]));
  R.m[1] = c2add(c2mul(A.m[0], B.m[1]), c2mul(A.m[1], B.m[3]));
  R.m[2] = c2add(c2mul(A.m[2], B.m[0]), c2mul(A.m[3], B.m[2]));
  R.m[3] = c2add(c2mul(A.m[2], B.m[1]), c2mul(A.m[3], B.m[3]));
  return R;
}
__device__ __forceinline__ M22 m22dag(const M22& A) {
  M22 R;
  R.m[0] = {A.m[0].x, -A.m[0].y};
  R.m[1] = {A.m[2].x, -A.m[2].y};
  R.m[2] = {A.m[1].x, -A.m[1].y};
  R.m[3] = {A.m[3].x, -A.m[3].y};
  return R;
}


// member matrix in column-vector convention (operators.py:368-395, t = angle/2)
template <typename T>
__device__ __forceinline__ M22 member_matrix(const Member& mb, const T* shared_angles, const T* batch_row, const T* fixed_mats) {
  M22 R;
  if (mb.kind == M_U) {
    const T* f = fixed_mats + (size_t)mb.slot * 8;
    for (int i = 0; i < 4; ++i) R.m[i] = {(double)f[2 * i], (double)f[2 * i + 1]};
    return R;
  }
  T ang = mb.batch ? batch_row[mb.slot] : shared_angles[mb.slot];
  // t = angle / 2 (operators.py:267, 271; exact in any binary precision).  sin / cos are evaluated in DOUBLE also for complex64
  // states: a float sincosf is 1-2 ulp off, i.e. the 2x2 is non-unitary at 1e-7, the SAME factor on every amplitude -- over the
  // ~500 rotations of a BASELINE circuit that coherent norm error was the engine's whole deviation from the float64 oracle
  // (1e-5 relative; 1e-6 with exactly rounded matrices).  The fused group product is formed in double and rounded once.
  double sd, cd;
  sincos(0.5 * (double)ang, &sd, &cd);
  if (mb.kind == M_RX) {
    R.m[0] = {cd, 0};
    R.m[1] = {0, -sd};
    R.m[2] = {0, -sd};
    R.m[3] = {cd, 0};
  } else if (mb.kind == M_RY) {
    R.m[0] = {cd, 0};
    R.m[1] = {-sd, 0};
    R.m[2] = {sd, 0};
    R.m[3] = {cd, 0};
  } else {
    R.m[0] = {cd, -sd};
    R.m[1] = {0, 0};
    R.m[2] = {0, 0};
    R.m[3] = {cd, sd};
  }
  return R;
}

// build the fused 2x2 of every group: U = M_r ... M_1
template <typename T>
__global__ void build_mats_kernel(const Group* __restrict__ groups, const Member* __restrict__ members, int n_groups,
                                  const T* __restrict__ shared_angles, const T* __restrict__ batch_angles,
                                  int n_batch_cols, const T* __restrict__ fixed_mats, T* __restrict__ mats_shared,
                                  T* __restrict__ mats_batch, int n_groups_batch, long B) {
  // thread space: shared groups once, batch groups per sample
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)n_groups * B;
  if (idx >= total) return;
  const int g = (int)(idx % n_groups);
  const long b = idx / n_groups;
  const Group grp = groups[g];
  if (!grp.batch && b != 0) return;
  const T* brow = batch_angles ? batch_angles + (size_t)b * n_batch_cols : nullptr;
  M22 U;
  U.m[0] = {1, 0};
  U.m[1] = {0, 0};
  U.m[2] = {0, 0};
  U.m[3] = {1, 0};
  for (int k = 0; k < grp.member_count; ++k) {
    M22 Mk = member_matrix<T>(members[grp.member_begin + k], shared_angles, brow, fixed_mats);
    U = m22mul(Mk, U);
  }
  T* out = grp.batch ? mats_batch + ((size_t)b * n_groups_batch + grp.mat_index) * 8 : mats_shared + (size_t)grp.mat_index * 8;
  for (int i = 0; i < 4; ++i) {
    out[2 * i] = (T)U.m[i].x;
    out[2 * i + 1] = (T)U.m[i].y;
  }
}

// dL/dtheta_k = Im <lam| V_k P_k V_k^+ |psi> = a_k . s,  V_k = M_r ... M_{k+1},  a_k = Pauli vector of the traceless
// Hermitian matrix V_k P_k V_k^+,  s = accumulated (sX, sY, sZ)   (DESIGN.md "fused-group adjoint")
template <typename T>
__global__ void finalize_grads_kernel(const Group* __restrict__ groups, const Member* __restrict__ members, int n_groups,
                                      const T* __restrict__ shared_angles, const T* __restrict__ batch_angles,
                                      int n_batch_cols, const T* __restrict__ fixed_mats, const T* __restrict__ k_shared,
                                      const T* __restrict__ k_batch, int n_k_batch, T* __restrict__ grad_shared,
                                      T* __restrict__ grad_batch, long B) {
  const long idx = blockIdx.x * (long)blockDim.x + threadIdx.x;
  const long total = (long)n_groups * B;
  if (idx >= total) return;
  const int g = (int)(idx % n_groups);
  const long b = idx / n_groups;
  const Group grp = groups[g];
  if (!grp.has_param) return;
  if (!grp.batch && b != 0) return;
  const T* brow = batch_angles ? batch_angles + (size_t)b * n_batch_cols : nullptr;
  const T* kp = grp.batch ? k_batch + ((size_t)b * n_k_batch + grp.k_index) * kAcc : k_shared + (size_t)grp.k_index * kAcc;
  const double sx = (double)kp[0], sy = (double)kp[1], sz = (double)kp[2];
  M22 V;
  V.m[0] = {1, 0};
  V.m[1] = {0, 0};
  V.m[2] = {0, 0};
  V.m[3] = {1, 0};
  for (int k = grp.member_count - 1; k >= 0; --k) {
    const Member mb = members[grp.member_begin + k];
    if (mb.kind != M_U) {
      M22 P;
      if (mb.kind == M_RX) {
        P.m[0] = {0, 0};
        P.m[1] = {1, 0};
        P.m[2] = {1, 0};
        P.m[3] = {0, 0};
      } else if (mb.kind == M_RY) {
        P.m[0] = {0, 0};
        P.m[1] = {0, -1};
        P.m[2] = {0, 1};
        P.m[3] = {0, 0};
      } else {
        P.m[0] = {1, 0};
        P.m[1] = {0, 0};
        P.m[2] = {0, 0};
        P.m[3] = {-1, 0};
      }
      M22 Am = m22mul(m22mul(V, P), m22dag(V));
      // A = ax X + ay Y + az Z:  A10 = ax + i ay,  A00 = az
      T gval = (T)(Am.m[2].x * sx + Am.m[2].y * sy + Am.m[0].x * sz);
      if (mb.batch)
        atomicAdd(grad_batch + (size_t)b * n_batch_cols + mb.slot, gval);
      else
        atomicAdd(grad_shared + mb.slot, gval);
    }
    if (k > 0) {
      M22 Mk = member_matrix<T>(mb, shared_angles, brow, fixed_mats);
      V = m22mul(V, Mk);
    }
  }
}

// ---------------------------------------------------------------------------------------------------
// |0...0>
template <typename T>
__global__ void init_zero_kernel(typename Vec2<T>::type* state, int n_local, long B, int is_rank0) {
  using T2 = typename Vec2<T>::type;
  const uint64_t total = (uint64_t)B << n_local;
  const uint64_t mask = (uint64_t(1) << n_local) - 1;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    T2 v = {0, 0};
    if (is_rank0 && (i & mask) == 0) v.x = 1;
    state[i] = v;
  }
}

// ---------------------------------------------------------------------------------------------------
// MeasureProbability (measurements.py:113-123), one pass over the state.
// CTA = 256 threads x 4 consecutive amplitudes = one 1024-amplitude segment per iteration.  Per thread
// running sums of its own |amp|^2 pattern give the 10 low index bits after one Walsh-Hadamard butterfly
// at the end; the higher bits are accumulated from per-segment totals by one thread per bit.
// part layout per CTA: [0] total, [1..10] W_j = S0_j - S1_j for local bits 0..9, [11 + j] S1 of local bit 10+j.
constexpr int kProbSegBits = 10;
constexpr int kProbPartStride = 11 + 40;

template <typename T>
__global__ void __launch_bounds__(256) probs_partial_kernel(const typename Vec2<T>::type* __restrict__ state, int n_local,
                                                            int cps, double* __restrict__ part) {
  using T2 = typename Vec2<T>::type;
  const int b = blockIdx.x / cps, c = blockIdx.x % cps;
  const T2* s = state + ((uint64_t)b << n_local);
  const uint64_t n_amp = uint64_t(1) << n_local;
  const uint64_t n_seg = (n_amp + 1023) >> kProbSegBits;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int n_high = n_local > kProbSegBits ? n_local - kProbSegBits : 0;
  __shared__ double wtot[2][8];
  __shared__ double fin[8][8];
  double tot = 0, w0 = 0, w1 = 0, hi_acc = 0;
  int parity = 0;
  // contiguous range of segments per CTA
  const uint64_t per = (n_seg + cps - 1) / cps;
  const uint64_t seg0 = (uint64_t)c * per, seg1 = min(n_seg, seg0 + per);
  for (uint64_t seg = seg0; seg < seg1; ++seg) {
    const uint64_t i = (seg << kProbSegBits) + (uint64_t)tid * 4;
    double p[4];
    if (i + 3 < n_amp) {
      // 16-byte vector loads (the segment is 16-byte aligned: i is a multiple of 4)
      constexpr int NV = 4 * sizeof(T2) / 16;
      int4 raw[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) raw[v] = unit_fix<T>(__ldcs(reinterpret_cast<const int4*>(s + i) + v));
      const T2* vv = reinterpret_cast<const T2*>(raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) p[e] = (double)vv[e].x * (double)vv[e].x + (double)vv[e].y * (double)vv[e].y;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        T2 v = {0, 0};
        if (i + e < n_amp) v = load_amp<T>(s, i + e);
        p[e] = (double)v.x * (double)v.x + (double)v.y * (double)v.y;
      }
    }
    const double t = (p[0] + p[1]) + (p[2] + p[3]);
    tot += t;
    w0 += (p[0] - p[1]) + (p[2] - p[3]);
    w1 += (p[0] + p[1]) - (p[2] + p[3]);
    if (n_high > 0) {
      double wt = warp_sum(t);
      if (lane == 0) wtot[parity][warp] = wt;
      __syncthreads();
      if (tid < n_high && ((seg >> tid) & 1ull)) {
        double st = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) st += wtot[parity][w];
        hi_acc += st;
      }
      parity ^= 1;
    }
  }
  // Walsh-Hadamard butterfly over the 32 lanes of tot: lane 0 -> sum, lane 2^j -> W for index bit 2+j
  double h = tot;
#pragma unroll
  for (int j = 0; j < 5; ++j) {
    double o = __shfl_xor_sync(0xffffffffu, h, 1 << j);
    h = ((lane >> j) & 1) ? (o - h) : (h + o);
  }
  w0 = warp_sum(w0);
  w1 = warp_sum(w1);
  // per warp: 8 values [total, W0, W1, W2..W6]
  if (lane == 0) {
    fin[warp][0] = h;
    fin[warp][1] = w0;
    fin[warp][2] = w1;
  }
  if (lane == 1 || lane == 2 || lane == 4 || lane == 8 || lane == 16) fin[warp][3 + (31 - __clz(lane))] = h;
  __syncthreads();
  double* out = part + (size_t)blockIdx.x * kProbPartStride;
  if (tid < 8) {
    double sum = 0;
    for (int w = 0; w < 8; ++w) sum += fin[w][tid];
    out[tid] = sum;  // [0] total, [1..7] W for bits 0..6
  } else if (tid < 11) {
    const int j = tid - 8;  // warp-index bit j -> index bit 7 + j
    double sum = 0;
    for (int w = 0; w < 8; ++w) sum += ((w >> j) & 1) ? -fin[w][0] : fin[w][0];
    out[tid] = sum;
  }
  if (tid < n_high) out[11 + tid] = hi_acc;
}

// probs_out[b][q] = P(q = 0) restricted to this rank's amplitudes
template <typename T>
__global__ void probs_finalize_kernel(const double* __restrict__ part, int cps, int n_qubits, int n_local,
                                      const int32_t* __restrict__ final_pos, int rank, T* __restrict__ probs_out) {
  const int b = blockIdx.x;
  const int q = threadIdx.x;
  if (q >= n_qubits) return;
  const int p = final_pos[q];
  double total = 0, val = 0;
  for (int c = 0; c < cps; ++c) {
    const double* pp = part + ((size_t)b * cps + c) * kProbPartStride;
    total += pp[0];
    if (p < kProbSegBits)
      val += pp[1 + p];
    else if (p < n_local)
      val += pp[11 + (p - kProbSegBits)];
  }
  double p0;
  if (p >= n_local)
    p0 = ((rank >> (p - n_local)) & 1) ? 0.0 : total;
  else if (p < kProbSegBits)
    p0 = 0.5 * (total + val);  // val = S0 - S1
  else
    p0 = total - val;  // val = S1
  probs_out[(size_t)b * n_qubits + q] = (T)p0;
}

// MeasureJointProbability (measurements.py:78-79); state in the internal layout, output in index order
template <typename T>
__global__ void joint_kernel(const typename Vec2<T>::type* __restrict__ state, T* __restrict__ out, uint64_t total) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    auto v = load_amp<T>(state, i);
    out[i] = v.x * v.x + v.y * v.y;
  }
}

// adjoint seeds  lambda = dL/dpsi*   (all states in the internal layout)
//   probs: lambda_i = (sum_q g_q [bit_pos(q)(i) == 0]) psi_i
template <typename T>
__global__ void __launch_bounds__(256) seed_probs_kernel(const typename Vec2<T>::type* __restrict__ state,
                                                         const T* __restrict__ grad, typename Vec2<T>::type* __restrict__ lam,
                                                         int n_qubits, int n_local, const int32_t* __restrict__ final_pos,
                                                         int rank, int cps) {
  using T2 = typename Vec2<T>::type;
  const int b = blockIdx.x / cps, c = blockIdx.x % cps;
  __shared__ T gbit[64];  // gradient per physical bit
  __shared__ T lowtab[1 << 10];
  __shared__ T ext_w;
  if (threadIdx.x < 64) gbit[threadIdx.x] = 0;
  __syncthreads();
  if (threadIdx.x < n_qubits) gbit[final_pos[threadIdx.x]] = grad[(size_t)b * n_qubits + threadIdx.x];
  __syncthreads();
  const int lowb = min(n_local, 10);
  for (int i = threadIdx.x; i < (1 << lowb); i += blockDim.x) {
    T w = 0;
    for (int k = 0; k < lowb; ++k)
      if (!((i >> k) & 1)) w += gbit[k];
    lowtab[i] = w;
  }
  if (threadIdx.x == 0) {
    T w = 0;
    for (int k = n_local; k < n_qubits; ++k)
      if (!((rank >> (k - n_local)) & 1)) w += gbit[k];
    ext_w = w;
  }
  __syncthreads();
  const uint64_t n_amp = uint64_t(1) << n_local;
  const T2* s = state + ((uint64_t)b << n_local);
  T2* l = lam + ((uint64_t)b << n_local);
  auto weight = [&](uint64_t i) {
    T w = ext_w + lowtab[i & ((1u << lowb) - 1u)];
    for (int k = 10; k < n_local; ++k)
      if (!((i >> k) & 1ull)) w += gbit[k];
    return w;
  };
  if constexpr (sizeof(T) == 4) {
    // one 16-byte unit (amplitudes 2u, 2u+1) per iteration
    const uint64_t n_unit = n_amp >> 1;
    const float4* su = reinterpret_cast<const float4*>(s);
    float4* lu = reinterpret_cast<float4*>(l);
    for (uint64_t u = (uint64_t)c * blockDim.x + threadIdx.x; u < n_unit; u += (uint64_t)cps * blockDim.x) {
      const T w0 = weight(2 * u), w1 = weight(2 * u + 1);
      float4 v = su[u];
      lu[u] = float4{v.x * w0, v.y * w1, v.z * w0, v.w * w1};
    }
  } else {
    for (uint64_t i = (uint64_t)c * blockDim.x + threadIdx.x; i < n_amp; i += (uint64_t)cps * blockDim.x) {
      const T w = weight(i);
      T2 v = s[i];
      v.x *= w;
      v.y *= w;
      l[i] = v;
    }
  }
}

//   joint: lambda_i = g_i psi_i   (grad in index order)
template <typename T>
__global__ void seed_joint_kernel(const typename Vec2<T>::type* __restrict__ state, const T* __restrict__ grad,
                                  typename Vec2<T>::type* __restrict__ lam, uint64_t total) {
  if constexpr (sizeof(T) == 4) {
    const float4* su = reinterpret_cast<const float4*>(state);
    float4* lu = reinterpret_cast<float4*>(lam);
    for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < (total >> 1); u += (uint64_t)gridDim.x * blockDim.x) {
      const float g0 = grad[2 * u], g1 = grad[2 * u + 1];
      const float4 v = su[u];
      lu[u] = float4{v.x * g0, v.y * g1, v.z * g0, v.w * g1};
    }
  } else {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
      auto v = state[i];
      T g = grad[i];
      v.x *= g;
      v.y *= g;
      lam[i] = v;
    }
  }
}

//   state: lambda = grad / 2   (torch's complex gradient convention is 2 dL/dpsi*); grad is interleaved complex (a
//   user tensor), lambda is written in the internal layout
template <typename T>
__global__ void seed_state_kernel(const typename Vec2<T>::type* __restrict__ grad, typename Vec2<T>::type* __restrict__ lam,
                                  uint64_t total) {
  if constexpr (sizeof(T) == 4) {
    const float4* gu = reinterpret_cast<const float4*>(grad);
    float4* lu = reinterpret_cast<float4*>(lam);
    for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < (total >> 1); u += (uint64_t)gridDim.x * blockDim.x) {
      const float4 g = gu[u];  // (re0, im0, re1, im1)
      lu[u] = float4{0.5f * g.x, 0.5f * g.z, 0.5f * g.y, 0.5f * g.w};
    }
  } else {
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
      auto v = grad[i];
      v.x *= (T)0.5;
      v.y *= (T)0.5;
      lam[i] = v;
    }
  }
}

// interleaved <-> internal layout, in place (complex64 only; an involution)
__global__ void convert_layout_kernel(int4* __restrict__ units, uint64_t n_units) {
  for (uint64_t u = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; u < n_units; u += (uint64_t)gridDim.x * blockDim.x) {
    const int4 v = units[u];
    units[u] = int4{v.x, v.z, v.y, v.w};
  }
}

// ---------------------------------------------------------------------------------------------------
// Global<->local qubit swap over NVLink peer memory (amplitude sharding, north_star (d)).
// An exchange step swaps rank bit j with local index bit pos[j] (plan.h: Exchange): the amplitude at (rank R, local index x) trades
// places with the one at (rank B(x), x with the bits pos[j] replaced by R's bits), B(x) = the bits of x at pos[j].  Amplitudes with
// B(x) = R stay.  Every rank works on every pair {R, C} it belongs to: of the 16-byte vectors involved, the lower rank swaps those
// whose `hbit` (a low index bit outside the exchanged ones) is 0 and the higher rank the others, reading one side from its own HBM
// and the other through the peer mapping and writing both back crosswise -- in place, no staging buffer, both NVLink directions
// busy.  Callers bracket the launch with a cross-rank barrier.  With pos = the top bits this is the all-to-all over contiguous chunks.
//
// Work index i = (sample, high free bits, peer, low free bits): the PEER varies faster than the 64 KB granule of low free bits and is
// counted from rank + 1, so at any moment the grid's traffic is spread over all peers and rank r starts with peer r + 1.  (Walking the
// peers one after the other -- every rank the same order -- is an incast on one GPU at a time: 417 GB/s per direction on 8 GPUs
// where the interleaved order reaches 647.)
struct PeerPtrs {
  void* p[16];
};

struct ExchangeBits {
  int32_t g;          // rank bits
  int32_t vbits;      // log2(16-byte vectors per shard)
  int32_t vpos[4];    // vector-index bit of rank bit j
  int32_t hbit;       // vector-index bit that splits a pair's work between its two ranks
  int32_t ins[5];     // vpos[] and hbit, ascending: where zero bits are inserted into the free index
};

__global__ void __launch_bounds__(256) exchange_p2p_kernel(const PeerPtrs peers, const ExchangeBits E, int rank, int world, int64_t batch) {
  int4* local = reinterpret_cast<int4*>(peers.p[rank]);
  const int fbits = E.vbits - E.g - 1;                          // free bits of a vector index
  const uint64_t n_free = uint64_t(1) << fbits;
  const uint64_t gran = fbits < 12 ? n_free : uint64_t(4096);    // 64 KB of vectors
  const uint64_t n_gran = n_free / gran;
  const uint64_t total = (uint64_t)batch * (uint64_t)(world - 1) * n_free;
  uint64_t mine = 0;  // my rank's bits at the exchanged positions
#pragma unroll
  for (int j = 0; j < 4; ++j)
    if (j < E.g) mine |= (uint64_t)((rank >> j) & 1) << E.vpos[j];
  constexpr int UN = 4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * UN) {
    int4 a[UN], b[UN];
    uint64_t lo[UN], ro[UN];
    int4* rp[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < total) {
        const uint64_t fl = i % gran;
        uint64_t t = i / gran;
        const int cc = (int)(t % (uint64_t)(world - 1));
        t /= (uint64_t)(world - 1);
        uint64_t x = (t % n_gran) * gran + fl;  // free index
        const uint64_t bb = t / n_gran;
        const int c = (rank + 1 + cc) % world;  // peer: never `rank`
#pragma unroll
        for (int k = 0; k < 5; ++k)
          if (k <= E.g) x = ((x >> E.ins[k]) << (E.ins[k] + 1)) | (x & ((uint64_t(1) << E.ins[k]) - 1));  // zero bit at ins[k]
        if (rank > c) x |= uint64_t(1) << E.hbit;
        uint64_t theirs = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (j < E.g) theirs |= (uint64_t)((c >> j) & 1) << E.vpos[j];
        lo[u] = (bb << E.vbits) + (x | theirs);  // mine: destination-rank bits = c
        ro[u] = (bb << E.vbits) + (x | mine);    // the peer's: destination-rank bits = rank
        rp[u] = reinterpret_cast<int4*>(peers.p[c]);
        a[u] = local[lo[u]];
        b[u] = rp[u][ro[u]];
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < total) {
        local[lo[u]] = b[u];
        rp[u][ro[u]] = a[u];
      }
    }
  }
}

// out = scale * in (used for grad_init_state = 2 * lambda_0)
template <typename T>
__global__ void scale_kernel(const T* in, T* out, T scale, uint64_t total) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * scale;
}

}  // namespace qb
