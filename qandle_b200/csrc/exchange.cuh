// Global<->local qubit swap of amplitude-sharded states, PUSH form (sm_100a): NVLink carries only posted writes.
//
// The exchange step of a sharded plan swaps the top log2(world) local index bits with the rank bits: chunk c of rank r <-> chunk r
// of rank c (plan.cpp; shard layout [batch][world][chunk]).  exchange_p2p_kernel (kernels.cuh) does it in place by PULLING: every
// rank reads half of each chunk pair through the peer mapping.  Remote reads are latency-bound -- measured on 8 B200s: 417 GB/s per
// direction on 2 GiB shards, 246 GB/s on 64 GiB shards (0.54 / 0.32 of the 770 GB/s a peer copy reaches), and the 7 exchanges of the
// 36-qubit config were 54 % of its forward.  Here a rank only WRITES to its peers:
//   phase 0 (push)    my chunk c, piece p  ->  slot `rank` of peer c's staging buffer       (NVLink, posted writes)
//   -- cross-rank barrier --
//   phase 1 (unpack)  slot c of my staging buffer  ->  my chunk c, piece p                  (local HBM copy)
// with the chunk cut into `pieces` so the staging buffer is world * chunk / pieces (36 qubits on 8 GPUs: 8 GiB next to the 2 x 64
// GiB of psi and lambda).  The push is done by the TMA engine: one thread per CTA streams 32 KB blocks HBM -> shared memory
// (cp.async.bulk ... mbarrier::complete_tx) -> peer memory (cp.async.bulk.global.shared::cta), four blocks in flight per CTA, no
// register traffic at all (SASS: UBLKCP, SYNCS).  exchange_push_ldst_kernel is the same copy with 16-byte loads / stores (fallback,
// A/B).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"

namespace qb {
namespace ex {

constexpr int kTmaBlock = 32 * 1024;  // bytes per bulk copy
constexpr int kTmaStages = 4;
constexpr int kTmaThreads = 32;
constexpr size_t kTmaSmem = (size_t)kTmaBlock * kTmaStages + 64;

struct PushArgs {
  PeerPtrs staging;      // staging buffer of every rank, peer-mapped: [world slots][batch][piece_vec] 16-byte vectors
  const int4* local;     // this rank's state shard: [batch][world][chunk_vec]
  int rank, world;
  int64_t batch;
  uint64_t chunk_vec;    // 16-byte vectors per chunk
  uint64_t piece_vec;    // 16-byte vectors per piece of a chunk
  uint64_t piece_off;    // first vector of this piece inside the chunk
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "QB_MBAR_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra QB_MBAR_DONE;\n"
      "bra QB_MBAR_WAIT;\n"
      "QB_MBAR_DONE:\n"
      "}\n" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(void* smem_dst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(smem_dst)),
               "l"(gsrc), "r"(bytes), "r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void bulk_s2g(void* gdst, const void* smem_src, uint32_t bytes) {
  asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(smem_src)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void bulk_wait_read() {
  asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

// block `g` of the push: (source in my shard, destination in the peer's staging buffer, bytes)
struct PushBlock {
  const char* src;
  char* dst;
  uint32_t bytes;
};
__device__ __forceinline__ PushBlock push_block(const PushArgs& A, uint64_t g, uint64_t blocks_per_seg) {
  const uint64_t piece_bytes = A.piece_vec * 16;
  // the PEER is the fastest index and counted from rank + 1: consecutive blocks (= consecutive CTAs) go to different peers and rank r
  // starts with peer r + 1, so no GPU is the target of everybody at once
  const int cc = (int)(g % (uint64_t)(A.world - 1));
  const uint64_t rest = g / (uint64_t)(A.world - 1);
  const uint64_t blk = rest % blocks_per_seg, b = rest / blocks_per_seg;
  const int c = (A.rank + 1 + cc) % A.world;
  const uint64_t off = blk * (uint64_t)kTmaBlock;
  PushBlock r;
  r.bytes = (uint32_t)(piece_bytes - off < (uint64_t)kTmaBlock ? piece_bytes - off : (uint64_t)kTmaBlock);
  r.src = reinterpret_cast<const char*>(A.local + ((b * A.world + c) * A.chunk_vec + A.piece_off)) + off;
  r.dst = reinterpret_cast<char*>(A.staging.p[c]) + (((uint64_t)A.rank * A.batch + b) * piece_bytes) + off;
  return r;
}

// phase 0 through the TMA engine: one issuing thread per CTA, kTmaStages blocks of 32 KB in flight
__global__ void __launch_bounds__(kTmaThreads) exchange_push_tma_kernel(const __grid_constant__ PushArgs A) {
  extern __shared__ __align__(128) unsigned char xsmem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(xsmem + (size_t)kTmaBlock * kTmaStages);
  if (threadIdx.x != 0) return;
  for (int s = 0; s < kTmaStages; ++s) mbar_init(&bars[s], 1);
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  const uint64_t piece_bytes = A.piece_vec * 16;
  const uint64_t blocks_per_seg = (piece_bytes + kTmaBlock - 1) / kTmaBlock;
  const uint64_t n_blocks = (uint64_t)A.batch * (uint64_t)(A.world - 1) * blocks_per_seg;
  // my blocks: blockIdx.x, + gridDim.x, ...  -- k-th of them lives in stage k % kTmaStages
  uint64_t n_mine = n_blocks > blockIdx.x ? (n_blocks - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
  auto load = [&](uint64_t k) {
    const PushBlock pb = push_block(A, blockIdx.x + k * gridDim.x, blocks_per_seg);
    const int s = (int)(k % kTmaStages);
    mbar_expect_tx(&bars[s], pb.bytes);
    bulk_g2s(xsmem + (size_t)s * kTmaBlock, pb.src, pb.bytes, &bars[s]);
  };
  for (uint64_t k = 0; k < (uint64_t)(kTmaStages - 1) && k < n_mine; ++k) load(k);
  for (uint64_t k = 0; k < n_mine; ++k) {
    const int s = (int)(k % kTmaStages);
    mbar_wait(&bars[s], (uint32_t)((k / kTmaStages) & 1));
    const PushBlock pb = push_block(A, blockIdx.x + k * gridDim.x, blocks_per_seg);
    bulk_s2g(pb.dst, xsmem + (size_t)s * kTmaBlock, pb.bytes);
    bulk_commit();
    if (k + kTmaStages - 1 < n_mine) {
      // the next load goes into the stage store k-1 read from: at most the store just committed may still be reading
      bulk_wait_read<1>();
      load(k + kTmaStages - 1);
    }
  }
  bulk_wait_all();  // every write has left this SM before the kernel (and the cross-rank barrier after it) completes
}

// phase 0 with ordinary 16-byte loads / stores (fallback; A/B against the TMA form)
__global__ void __launch_bounds__(256) exchange_push_ldst_kernel(const __grid_constant__ PushArgs A) {
  const uint64_t total = (uint64_t)A.batch * (uint64_t)(A.world - 1) * A.piece_vec;
  constexpr int UN = 4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * UN) {
    int4 v[UN];
    int4* d[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < total) {
        // (sample, granule, peer, vector in the granule): peers interleaved at 64 KB, counted from rank + 1 (see push_block)
        const uint64_t gran = A.piece_vec < 4096 ? A.piece_vec : 4096;
        const uint64_t wl = i % gran;
        uint64_t t = i / gran;
        const int c = (A.rank + 1 + (int)(t % (uint64_t)(A.world - 1))) % A.world;
        t /= (uint64_t)(A.world - 1);
        const uint64_t n_gran = A.piece_vec / gran;
        const uint64_t w = (t % n_gran) * gran + wl, b = t / n_gran;
        v[u] = __ldcs(A.local + (b * A.world + c) * A.chunk_vec + A.piece_off + w);
        d[u] = reinterpret_cast<int4*>(A.staging.p[c]) + ((uint64_t)A.rank * A.batch + b) * A.piece_vec + w;
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
      if (i0 + u * stride < total) *d[u] = v[u];
  }
}

// phase 1: slot c of my staging buffer -> my chunk c, piece p (all c != rank); a local HBM copy
__global__ void __launch_bounds__(256) exchange_unpack_kernel(const int4* __restrict__ staging, int4* __restrict__ local, int rank, int world,
                                                              int64_t batch, uint64_t chunk_vec, uint64_t piece_vec, uint64_t piece_off) {
  const uint64_t total = (uint64_t)batch * (uint64_t)(world - 1) * piece_vec;
  constexpr int UN = 4;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i0 = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i0 < total; i0 += stride * UN) {
    int4 v[UN];
    uint64_t o[UN];
#pragma unroll
    for (int u = 0; u < UN; ++u) {
      const uint64_t i = i0 + u * stride;
      if (i < total) {
        const uint64_t w = i % piece_vec, t = i / piece_vec;
        int c = (int)(t % (uint64_t)(world - 1));
        const uint64_t b = t / (uint64_t)(world - 1);
        if (c >= rank) ++c;
        v[u] = __ldcs(staging + ((uint64_t)c * batch + b) * piece_vec + w);
        o[u] = (b * world + c) * chunk_vec + piece_off + w;
      }
    }
#pragma unroll
    for (int u = 0; u < UN; ++u)
      if (i0 + u * stride < total) local[o[u]] = v[u];
  }
}

}  // namespace ex
}  // namespace qb
