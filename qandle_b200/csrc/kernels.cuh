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
// tile <-> HBM.  A tile is 2^m amplitudes; its chunk h (2^L contiguous amplitudes) lives at element offset
// base + (hi_off[h] << L).  16-byte vector accesses, consecutive threads -> consecutive vectors.
template <typename T>
__device__ __forceinline__ void tile_load(typename Vec2<T>::type* tile, const typename Vec2<T>::type* g, uint64_t base,
                                          const uint32_t* hi_off, int m, int L) {
  constexpr int EPV = 16 / sizeof(typename Vec2<T>::type);  // amplitudes per 16-byte vector: 2 (c64) / 1 (c128)
  constexpr int LEPV = EPV == 2 ? 1 : 0;
  const int n_vec = (1 << m) >> LEPV;
  const int vpc_log = L - LEPV;
  for (int v = threadIdx.x; v < n_vec; v += blockDim.x) {
    int h = v >> vpc_log;
    int w = v & ((1 << vpc_log) - 1);
    uint64_t e = base + ((uint64_t)hi_off[h] << L) + ((uint64_t)w << LEPV);
    int4 val = __ldcs(reinterpret_cast<const int4*>(g + e));
    *reinterpret_cast<int4*>(tile + ((h << L) + (w << LEPV))) = val;
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
    __stcs(reinterpret_cast<int4*>(g + e), val);
  }
}

// ---------------------------------------------------------------------------------------------------
// shared-memory layout of a sweep CTA (dynamic):
//   [tile psi: 2^m T2][tile lam: 2^m T2 (backward)][smats: n_ops*8 T][acc: n_kslots*8 T (bwd)]
//   [wpart: 2*kMaxWarps*8 T (bwd)][hi_off: 2^(m-L) u32][ops: n_ops KOp]
__host__ __device__ inline size_t sweep_smem_bytes(int m, int L, int n_ops, int n_kslots, bool backward, size_t szT) {
  size_t b = (size_t(1) << m) * 2 * szT * (backward ? 2 : 1);
  b += size_t(n_ops) * 8 * szT;
  if (backward) b += size_t(n_kslots) * 8 * szT + size_t(2 * kMaxWarps * 8) * szT;
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
// backward sweep (adjoint-state method): ops in reverse; for each op G:
//   K'[i][j] += sum psi_out[i] conj(lam_out[j])   (parametrised groups only)
//   psi <- G^+ psi,  lam <- G^+ lam
template <typename T> struct Acc8 {
  T v[8];
};

template <typename T>
__device__ __forceinline__ void block_accumulate8(Acc8<T>& r, T* wpart, T* acc_slot, int parity) {
#pragma unroll
  for (int j = 0; j < 8; ++j) r.v[j] = warp_sum(r.v[j]);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  T* wp = wpart + parity * (kMaxWarps * 8);
  if (lane == 0) {
#pragma unroll
    for (int j = 0; j < 8; ++j) wp[warp * 8 + j] = r.v[j];
  }
  __syncthreads();
  if (threadIdx.x < 8) {
    T s = 0;
    const int nw = blockDim.x >> 5;
    for (int w = 0; w < nw; ++w) s += wp[w * 8 + threadIdx.x];
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
  T* wpart = acc + size_t(A.n_kslots) * 8;
  size_t off = ((size_t(2) << m) * sizeof(T2) +
                (size_t(A.n_ops) * 8 + size_t(A.n_kslots) * 8 + size_t(2 * kMaxWarps * 8)) * sizeof(T) + 15) &
               ~size_t(15);
  uint32_t* hi_off = reinterpret_cast<uint32_t*>(smem_raw + off);
  off = (off + (size_t(1) << (m - L)) * 4 + 15) & ~size_t(15);
  KOp* sops = reinterpret_cast<KOp*>(smem_raw + off);

  const int b = blockIdx.x / A.cps;
  const int c = blockIdx.x % A.cps;
  sweep_setup<T>(A, b, smats, hi_off, sops);
  for (int i = threadIdx.x; i < A.n_kslots * 8; i += blockDim.x) acc[i] = 0;
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
    // <lam|psi> over the tile is invariant under the in-tile unitaries: compute it once for all K_D1_EXT grads
    T2 tdot = {0, 0};
    if (A.need_tile_dot) {
      T2 s = {0, 0};
      for (uint32_t i = tid; i < full; i += nthr) cacc_conj(s, tp[i], tl[i]);
      Acc8<T> r;
      r.v[0] = s.x;
      r.v[1] = s.y;
#pragma unroll
      for (int j = 2; j < 8; ++j) r.v[j] = 0;
      // reuse the block reduction; result broadcast through wpart
      r.v[0] = warp_sum(r.v[0]);
      r.v[1] = warp_sum(r.v[1]);
      T* wp = wpart + parity * (kMaxWarps * 8);
      if ((tid & 31) == 0) {
        wp[(tid >> 5) * 8 + 0] = r.v[0];
        wp[(tid >> 5) * 8 + 1] = r.v[1];
      }
      __syncthreads();
      const int nw = nthr >> 5;
      for (int w = 0; w < nw; ++w) {
        tdot.x += wp[w * 8 + 0];
        tdot.y += wp[w * 8 + 1];
      }
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
          if (op.kslot >= 0) {
            T2 k00 = {0, 0}, k01 = {0, 0}, k10 = {0, 0}, k11 = {0, 0};
            for (uint32_t k = tid; k < half; k += nthr) {
              uint32_t i0 = ins0(k, a), i1 = i0 | (1u << a);
              T2 x = tp[i0], y = tp[i1], lx = tl[i0], ly = tl[i1];
              cacc_conj(k00, x, lx);
              cacc_conj(k01, x, ly);
              cacc_conj(k10, y, lx);
              cacc_conj(k11, y, ly);
              tp[i0] = cfma(a01, y, cmul(a00, x));
              tp[i1] = cfma(a11, y, cmul(a10, x));
              tl[i0] = cfma(a01, ly, cmul(a00, lx));
              tl[i1] = cfma(a11, ly, cmul(a10, lx));
            }
            Acc8<T> r = {{k00.x, k00.y, k01.x, k01.y, k10.x, k10.y, k11.x, k11.y}};
            block_accumulate8<T>(r, wpart, acc + op.kslot * 8, parity);
            parity ^= 1;
            synced = true;
          } else {
            for (uint32_t k = tid; k < half; k += nthr) {
              uint32_t i0 = ins0(k, a), i1 = i0 | (1u << a);
              T2 x = tp[i0], y = tp[i1], lx = tl[i0], ly = tl[i1];
              tp[i0] = cfma(a01, y, cmul(a00, x));
              tp[i1] = cfma(a11, y, cmul(a10, x));
              tl[i0] = cfma(a01, ly, cmul(a00, lx));
              tl[i1] = cfma(a11, ly, cmul(a10, lx));
            }
          }
          break;
        }
        case K_D1: {
          const T2 d0 = {M[0], -M[1]}, d1 = {M[6], -M[7]};
          const int a = op.a;
          T2 k00 = {0, 0}, k11 = {0, 0};
          for (uint32_t i = tid; i < full; i += nthr) {
            T2 x = tp[i], lx = tl[i];
            if ((i >> a) & 1u) {
              cacc_conj(k11, x, lx);
              tp[i] = cmul(x, d1);
              tl[i] = cmul(lx, d1);
            } else {
              cacc_conj(k00, x, lx);
              tp[i] = cmul(x, d0);
              tl[i] = cmul(lx, d0);
            }
          }
          if (op.kslot >= 0) {
            Acc8<T> r = {{k00.x, k00.y, 0, 0, 0, 0, k11.x, k11.y}};
            block_accumulate8<T>(r, wpart, acc + op.kslot * 8, parity);
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
          if (op.kslot >= 0 && tid == 0) {
            T* s = acc + op.kslot * 8 + (one ? 6 : 0);
            s[0] += tdot.x;
            s[1] += tdot.y;
          }
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
  T* out = reinterpret_cast<T*>(A.partials) + (size_t)blockIdx.x * A.n_kslots * 8;
  for (int i = threadIdx.x; i < A.n_kslots * 8; i += blockDim.x) out[i] = acc[i];
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
  __shared__ double red[8][33];
  if (!ks.batch) {
    const int j = threadIdx.x & 7, lane8 = threadIdx.x >> 3, n8 = blockDim.x >> 3;
    double sum = 0;
    const long total = (long)B * cps;
    for (long r = lane8; r < total; r += n8) sum += (double)partials[((size_t)r * n_kslots + s) * 8 + j];
    // blockDim.x == 256 -> 32 partial sums per j
    red[j][lane8] = sum;
    __syncthreads();
    if (threadIdx.x < 8) {
      double t = 0;
      for (int i = 0; i < n8; ++i) t += red[threadIdx.x][i];
      k_shared[(size_t)ks.k_index * 8 + threadIdx.x] = (T)t;
    }
  } else {
    // one thread per (b, j)
    for (long idx = threadIdx.x; idx < (long)B * 8; idx += blockDim.x) {
      long bb = idx >> 3;
      int j = idx & 7;
      double sum = 0;
      for (int cc = 0; cc < cps; ++cc) sum += (double)partials[(((size_t)bb * cps + cc) * n_kslots + s) * 8 + j];
      k_batch[((size_t)bb * n_k_batch + ks.k_index) * 8 + j] = (T)sum;
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

__device__ __forceinline__ void sincos_T(float a, float* s, float* c) { sincosf(a, s, c); }
__device__ __forceinline__ void sincos_T(double a, double* s, double* c) { sincos(a, s, c); }

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
  T half = ang / (T)2;  // halve in the state's precision, like the reference (operators.py:267, 271)
  T s, c;
  sincos_T(half, &s, &c);
  double cd = (double)c, sd = (double)s;
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

// dL/dtheta_k = Im tr(V_k P_k V_k^+ K'),  V_k = M_r ... M_{k+1}   (DESIGN.md "fused-group adjoint")
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
  const T* kp = grp.batch ? k_batch + ((size_t)b * n_k_batch + grp.k_index) * 8 : k_shared + (size_t)grp.k_index * 8;
  M22 K;
  for (int i = 0; i < 4; ++i) K.m[i] = {(double)kp[2 * i], (double)kp[2 * i + 1]};
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
      // tr(A K) = sum_ij A_ij K_ji
      C2 tr = c2add(c2add(c2mul(Am.m[0], K.m[0]), c2mul(Am.m[1], K.m[2])), c2add(c2mul(Am.m[2], K.m[1]), c2mul(Am.m[3], K.m[3])));
      T gval = (T)tr.y;
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
      for (int v = 0; v < NV; ++v) raw[v] = __ldcs(reinterpret_cast<const int4*>(s + i) + v);
      const T2* vv = reinterpret_cast<const T2*>(raw);
#pragma unroll
      for (int e = 0; e < 4; ++e) p[e] = (double)vv[e].x * (double)vv[e].x + (double)vv[e].y * (double)vv[e].y;
    } else {
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        T2 v = {0, 0};
        if (i + e < n_amp) v = s[i + e];
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

// MeasureJointProbability (measurements.py:78-79)
template <typename T>
__global__ void joint_kernel(const typename Vec2<T>::type* __restrict__ state, T* __restrict__ out, uint64_t total) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    auto v = state[i];
    out[i] = v.x * v.x + v.y * v.y;
  }
}

// adjoint seeds  lambda = dL/dpsi*
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
  for (uint64_t i = (uint64_t)c * blockDim.x + threadIdx.x; i < n_amp; i += (uint64_t)cps * blockDim.x) {
    T w = ext_w + lowtab[i & ((1u << lowb) - 1u)];
    for (int k = 10; k < n_local; ++k)
      if (!((i >> k) & 1ull)) w += gbit[k];
    T2 v = s[i];
    v.x *= w;
    v.y *= w;
    l[i] = v;
  }
}

//   joint: lambda_i = g_i psi_i
template <typename T>
__global__ void seed_joint_kernel(const typename Vec2<T>::type* __restrict__ state, const T* __restrict__ grad,
                                  typename Vec2<T>::type* __restrict__ lam, uint64_t total) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    auto v = state[i];
    T g = grad[i];
    v.x *= g;
    v.y *= g;
    lam[i] = v;
  }
}

//   state: lambda = grad / 2   (torch's complex gradient convention is 2 dL/dpsi*)
template <typename T>
__global__ void seed_state_kernel(const typename Vec2<T>::type* __restrict__ grad, typename Vec2<T>::type* __restrict__ lam,
                                  uint64_t total) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    auto v = grad[i];
    v.x *= (T)0.5;
    v.y *= (T)0.5;
    lam[i] = v;
  }
}

// out = scale * in (used for grad_init_state = 2 * lambda_0)
template <typename T>
__global__ void scale_kernel(const T* in, T* out, T scale, uint64_t total) {
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x)
    out[i] = in[i] * scale;
}

}  // namespace qb
