// Micro-benchmark: do the shared-memory phase and the FP32 phase of a sweep STAGE overlap between the CTAs of an SM?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/phase_overlap tools/ubench/phase_overlap.cu
//   run:   tools/ubench/phase_overlap
// A stage of the flat sweep kernels (flat64.cuh) is  [8 x LDS.128 per state] [G 2x2s = 64 G FFMA2 on 16 amplitudes] [8 x STS.128]
// [CTA barrier].  Measured on config 2 / 20 qubits, a stage costs about (FP time) + (shared-memory time), as if the three resident CTAs
// of an SM ran their phases in lock-step.  This benchmark runs the same phase structure on synthetic data in three forms:
//   mode 0  3 CTAs x 256 threads per SM, free-running (what the sweep kernels do)
//   mode 1  1 CTA x 768 threads per SM = 3 groups of 256 threads, each group its own tile and its own 256-thread named barrier
//           (same as mode 0 but inside one CTA: checks that the grouping itself costs nothing)
//   mode 2  mode 1 + an FP TOKEN passed round-robin between the groups with named barriers (bar.arrive / bar.sync): only one group is
//           in its FP phase at a time, the other two do their shared-memory phases meanwhile
// and prints cycles per (CTA-)stage per SM next to the FP-only and shared-memory-only floors.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

constexpr int NP = 8;

__device__ __forceinline__ float2 f2fma(float2 a, float2 b, float2 c) { return __ffma2_rn(a, b, c); }
__device__ __forceinline__ float2 f2mul(float2 a, float2 b) { return __fmul2_rn(a, b); }

template <int RBIT>
__device__ __forceinline__ void u1(float2 (&R)[NP], float2 (&I)[NP], const float* M) {
  const float4 m0 = reinterpret_cast<const float4*>(M)[0], m1 = reinterpret_cast<const float4*>(M)[1];
  const float2 ar = {m0.x, m0.x}, ai = {m0.y, m0.y}, br = {m0.z, m0.z}, bi = {m0.w, m0.w};
  const float2 cr = {m1.x, m1.x}, ci = {m1.y, m1.y}, dr = {m1.z, m1.z}, di = {m1.w, m1.w};
  const float2 nai = {-m0.y, -m0.y}, nbi = {-m0.w, -m0.w}, nci = {-m1.y, -m1.y}, ndi = {-m1.w, -m1.w};
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

__device__ __forceinline__ void bar_sync(int id, int n) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(n) : "memory"); }
// the same with an immediate barrier id (SASS: BAR.SYNC 0x1, 0x100 instead of BAR.SYNC R44, 0x100)
__device__ __forceinline__ void bar_sync_group(int group) {
  if (group == 0)
    asm volatile("bar.sync 1, 256;" ::: "memory");
  else if (group == 1)
    asm volatile("bar.sync 2, 256;" ::: "memory");
  else
    asm volatile("bar.sync 3, 256;" ::: "memory");
}
__device__ __forceinline__ void bar_arrive(int id, int n) { asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(n) : "memory"); }

// PASSES: shared-memory round trips per stage (1 = forward stage; the streaming adjoint stage moves 2.5x as much)
// GATES:  2x2s per FP phase (3 = forward stage; ~8 = adjoint: psi, lambda and the Pauli sums)
// DO_FP / DO_SMEM: switch a phase off to measure the other one's floor
template <int MODE, int GATES, int PASSES, bool DO_FP, bool DO_SMEM>
__global__ void __launch_bounds__(MODE == 0 ? 256 : 768, MODE == 0 ? 3 : 1) k_phase(float* out, const float* mats, int stages, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int group = MODE == 0 ? 0 : threadIdx.x >> 8;
  const int t = threadIdx.x & 255;
  unsigned char* tile = smem + (MODE == 0 ? 0 : group * (32 << 10));
  float* sm = reinterpret_cast<float*>(smem + (MODE == 0 ? (32 << 10) : 3 * (32 << 10)));
  for (int i = threadIdx.x; i < 32; i += blockDim.x) sm[i] = mats[i];
  for (int j = 0; j < NP; ++j)
    *reinterpret_cast<float4*>(tile + (t + 256 * j) * 16) = make_float4(1.f + t * 1e-3f, 0.5f, -0.25f + j, 1e-2f * t);
  __syncthreads();
  if (MODE == 2 && group == 2) bar_arrive(4 + 0, 512);  // the token starts at group 0
  float2 R[NP], I[NP];
#pragma unroll
  for (int j = 0; j < NP; ++j) R[j] = I[j] = make_float2(0.f, 0.f);
  const long long t0 = clock64();
  for (int s = 0; s < stages; ++s) {
    // load phase: conflict-free LDS.128 (consecutive threads -> consecutive units), the partner pattern changes with the stage
    const unsigned rot = (unsigned)(s * 37) & 255u;
    if (DO_SMEM) {
#pragma unroll
      for (int p = 0; p < PASSES; ++p)
#pragma unroll
        for (int j = 0; j < NP; ++j) {
          const float4 v = *reinterpret_cast<const float4*>(tile + ((((unsigned)t + rot + 31u * p) & 255u) + 256 * j) * 16);
          if (p == 0) {
            R[j] = make_float2(v.x, v.y), I[j] = make_float2(v.z, v.w);
          } else {  // a streamed second state: two FFMA2 per unit (the adjoint sweep's Pauli sums use the streamed lambda like this)
            R[j] = f2fma(make_float2(v.x, v.y), make_float2(1e-3f, 1e-3f), R[j]);
            I[j] = f2fma(make_float2(v.z, v.w), make_float2(1e-3f, 1e-3f), I[j]);
          }
        }
    }
    if (MODE == 2) bar_sync(4 + group, 512);  // wait for the FP token
    if (DO_FP) {
#pragma unroll
      for (int g = 0; g < GATES; ++g) {
        if (g % 3 == 0) u1<0>(R, I, sm + 8 * (g & 3));
        if (g % 3 == 1) u1<1>(R, I, sm + 8 * (g & 3));
        if (g % 3 == 2) u1<2>(R, I, sm + 8 * (g & 3));
      }
    }
    if (MODE == 2) bar_arrive(4 + (group + 1) % 3, 512);  // pass it on
    if (DO_SMEM) {
      if (!DO_FP) {  // keep the round trip alive when the FP phase is switched off
#pragma unroll
        for (int j = 0; j < NP; ++j) R[j].x = __uint_as_float(__float_as_uint(R[j].x) ^ 0x80000000u);
      }
#pragma unroll
      for (int p = 0; p < PASSES; ++p)
#pragma unroll
        for (int j = 0; j < NP; ++j)
          *reinterpret_cast<float4*>(tile + ((((unsigned)t + rot + 31u * p) & 255u) + 256 * j) * 16) = make_float4(R[j].x, R[j].y, I[j].x, I[j].y);
    }
    if (MODE == 0)
      __syncthreads();
    else if (MODE == 3)
      bar_sync_group(group);
    else if (MODE != 4)
      bar_sync(1 + group, 256);
  }
  const long long t1 = clock64();
  if (MODE == 2 && group == 0) bar_sync(4 + 0, 512);  // absorb the last token so that no barrier is left half-arrived
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < NP; ++j) acc += R[j].x + R[j].y + I[j].x + I[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int GATES, int PASSES, bool DO_FP, bool DO_SMEM>
double run(int sms, int stages, float* out, const float* mats, long long* cyc) {
  auto kern = k_phase<MODE, GATES, PASSES, DO_FP, DO_SMEM>;
  const int threads = MODE == 0 ? 256 : 768;
  const int smem = (MODE == 0 ? (32 << 10) : 3 * (32 << 10)) + 1024 + (MODE == 0 ? (36 << 10) : 0);  // mode 0: pad so that exactly 3 CTAs fit
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = MODE == 0 ? sms * 3 : sms;
  kern<<<grid, threads, smem>>>(out, mats, 8, cyc);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0), cudaEventCreate(&e1);
  cudaEventRecord(e0);
  kern<<<grid, threads, smem>>>(out, mats, stages, cyc);
  cudaEventRecord(e1);
  cudaDeviceSynchronize();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(err));
    exit(1);
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (long long c : h) mean += (double)c;
  mean /= grid;
  // cycles per CTA-stage per SM: 3 tile-stages run concurrently on an SM, so an SM finishes one every (elapsed / stages / 3)
  return mean / stages / 3.0;
}


// ---- the same stage with the real kernels' stage-start overheads, switched on one by one (3 CTAs x 256 threads, free-running) -----------
// TAB:  per-stage descriptor (2 LDS.128, uniform address) -> thread-group nibble tables (2 LDS.U16, dependent on nothing but the stage
//       index) + per-tile XOR constant (LDS.64) -> load / store address tables (2 + 2 LDS.128) -> 16 LOP3 for the unit addresses, as
//       fl::run_stages does (flat64.cuh)
// XB:   a second CTA barrier between the loads and the stores (stages whose absorbed CNOTs move amplitudes between threads)
// SW:   the 2x2s behind a switch over the stage's shape (read from the descriptor): an indirect branch per stage
template <int GATES, bool TAB, bool XB, bool SW>
__global__ void __launch_bounds__(256, 3) k_stage(float* out, const float* mats, int stages, long long* cyc) {
  extern __shared__ __align__(16) unsigned char smem[];
  const int t = threadIdx.x;
  unsigned char* tile = smem;
  float* sm = reinterpret_cast<float*>(smem + (32 << 10));
  uint4* desc = reinterpret_cast<uint4*>(smem + (33 << 10));              // [16][2]
  unsigned short* ttab = reinterpret_cast<unsigned short*>(smem + (34 << 10));  // [16 stages][2 sides][32]
  unsigned* stab = reinterpret_cast<unsigned*>(smem + (36 << 10));        // [16][2][8]
  uint2* extc = reinterpret_cast<uint2*>(smem + (38 << 10));              // [16]
  for (int i = t; i < 32; i += 256) sm[i] = mats[i];
  for (int i = t; i < 16 * 2 * 32; i += 256) {
    const int si = i >> 6, e = i & 31;
    const unsigned rot = (unsigned)(si * 37 + ((i >> 5) & 1) * 11) & 255u;
    ttab[i] = (unsigned short)(e < 16 ? ((unsigned)e ^ (rot & 15u)) : ((((unsigned)e - 16u) << 4) ^ (rot & 0xF0u)));
  }
  for (int i = t; i < 16 * 2 * 8; i += 256) stab[i] = (unsigned)(i & 7) * 256u * 16u;
  for (int i = t; i < 16; i += 256) {
    extc[i] = make_uint2(0u, 0u);
    desc[2 * i] = make_uint4(0, (unsigned)(i & 3) << 16, 0, 8);
    desc[2 * i + 1] = make_uint4(16, 24, 0, 0);
  }
  for (int j = 0; j < NP; ++j)
    *reinterpret_cast<float4*>(tile + (t + 256 * j) * 16) = make_float4(1.f + t * 1e-3f, 0.5f, -0.25f + j, 1e-2f * t);
  __syncthreads();
  const unsigned short* tt_lo = ttab + (t & 15);
  const unsigned short* tt_hi = ttab + 16 + (t >> 4);
  float2 R[NP], I[NP];
  const long long t0 = clock64();
  for (int s = 0; s < stages; ++s) {
    const int si = s & 15;
    unsigned sbl, sbs, twl[NP], tws[NP];
    int shape = si & 3;
    const float *M0 = sm, *M1 = sm + 8, *M2 = sm + 16, *M3 = sm + 24;
    if (TAB) {
      const uint4 dw0 = desc[2 * si], dw1 = desc[2 * si + 1];
      shape = (dw0.y >> 16) & 3;
      M0 = sm + (dw0.z & 0xFFFFu), M1 = sm + (dw0.w & 0xFFFFu), M2 = sm + (dw1.x & 0xFFFFu), M3 = sm + (dw1.y & 0xFFFFu);
      const uint2 ex = extc[si];
      sbl = ((unsigned)(tt_lo[si * 64] ^ tt_hi[si * 64]) << 4) ^ ex.x;
      sbs = ((unsigned)(tt_lo[si * 64 + 32] ^ tt_hi[si * 64 + 32]) << 4) ^ ex.y;
      const uint4 a = reinterpret_cast<const uint4*>(stab + si * 16)[0], b = reinterpret_cast<const uint4*>(stab + si * 16)[1];
      const uint4 c = reinterpret_cast<const uint4*>(stab + si * 16)[2], d = reinterpret_cast<const uint4*>(stab + si * 16)[3];
      twl[0] = a.x, twl[1] = a.y, twl[2] = a.z, twl[3] = a.w, twl[4] = b.x, twl[5] = b.y, twl[6] = b.z, twl[7] = b.w;
      tws[0] = c.x, tws[1] = c.y, tws[2] = c.z, tws[3] = c.w, tws[4] = d.x, tws[5] = d.y, tws[6] = d.z, tws[7] = d.w;
    } else {
      sbl = (((unsigned)t + (unsigned)(s * 37)) & 255u) * 16u;
      sbs = (((unsigned)t + (unsigned)(s * 37 + 11)) & 255u) * 16u;
#pragma unroll
      for (int j = 0; j < NP; ++j) twl[j] = tws[j] = (unsigned)j * 256u * 16u;
    }
#pragma unroll
    for (int j = 0; j < NP; ++j) {
      const float4 v = *reinterpret_cast<const float4*>(tile + (TAB ? (sbl ^ twl[j]) : (sbl + twl[j])));
      R[j] = make_float2(v.x, v.y), I[j] = make_float2(v.z, v.w);
    }
    if (XB) __syncthreads();
    if (SW) {
      switch (shape) {
        case 0: u1<0>(R, I, M0); u1<1>(R, I, M1); if (GATES > 2) u1<2>(R, I, M2); break;
        case 1: u1<1>(R, I, M0); u1<2>(R, I, M1); if (GATES > 2) u1<0>(R, I, M2); break;
        case 2: u1<2>(R, I, M0); u1<0>(R, I, M1); if (GATES > 2) u1<1>(R, I, M3); break;
        default: u1<0>(R, I, M3); u1<2>(R, I, M1); if (GATES > 2) u1<1>(R, I, M2); break;
      }
    } else {
      u1<0>(R, I, M0); u1<1>(R, I, M1); if (GATES > 2) u1<2>(R, I, M2);
    }
#pragma unroll
    for (int j = 0; j < NP; ++j)
      *reinterpret_cast<float4*>(tile + (TAB ? (sbs ^ tws[j]) : (sbs + tws[j]))) = make_float4(R[j].x, R[j].y, I[j].x, I[j].y);
    __syncthreads();
  }
  const long long t1 = clock64();
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < NP; ++j) acc += R[j].x + R[j].y + I[j].x + I[j].y;
  out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int GATES, bool TAB, bool XB, bool SW>
double run_stage(int sms, int stages, float* out, const float* mats, long long* cyc) {
  auto kern = k_stage<GATES, TAB, XB, SW>;
  const int smem = 69 << 10;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  const int grid = sms * 3;
  kern<<<grid, 256, smem>>>(out, mats, 16, cyc);
  kern<<<grid, 256, smem>>>(out, mats, stages, cyc);
  cudaDeviceSynchronize();
  cudaError_t err = cudaGetLastError();
  if (err != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(err));
    exit(1);
  }
  std::vector<long long> h(grid);
  cudaMemcpy(h.data(), cyc, grid * sizeof(long long), cudaMemcpyDeviceToHost);
  double mean = 0;
  for (long long c : h) mean += (double)c;
  return mean / grid / stages / 3.0;
}

template <int GATES, int PASSES>
void report(const char* name, int sms, int stages, float* out, const float* mats, long long* cyc) {
  const double fp = run<0, GATES, PASSES, true, false>(sms, stages, out, mats, cyc);
  const double sm = run<0, GATES, PASSES, false, true>(sms, stages, out, mats, cyc);
  const double m0 = run<0, GATES, PASSES, true, true>(sms, stages, out, mats, cyc);
  const double m1 = run<1, GATES, PASSES, true, true>(sms, stages, out, mats, cyc);
  const double m2 = run<2, GATES, PASSES, true, true>(sms, stages, out, mats, cyc);
  const double fp1 = run<1, GATES, PASSES, true, false>(sms, stages, out, mats, cyc);
  const double sm1 = run<1, GATES, PASSES, false, true>(sms, stages, out, mats, cyc);
  const double fp3 = run<3, GATES, PASSES, true, false>(sms, stages, out, mats, cyc);
  const double fp4 = run<4, GATES, PASSES, true, false>(sms, stages, out, mats, cyc);
  const double m3 = run<3, GATES, PASSES, true, true>(sms, stages, out, mats, cyc);
  printf("    768-thread CTA: FP only with immediate barrier ids %7.1f, without any barrier %7.1f; both phases, immediate ids %7.1f\n", fp3, fp4, m3);
  printf("%-34s cycles per tile-stage per SM: FP only %7.1f (768-thr CTA %7.1f)  smem only %7.1f (%7.1f)  | 3 CTAs free-running %7.1f  "
         "1 CTA x 3 groups %7.1f  + FP token %7.1f   (sum %7.1f, max %7.1f)\n",
         name, fp, fp1, sm, sm1, m0, m1, m2, fp + sm, fp > sm ? fp : sm);
}

int main() {
  cudaDeviceProp prop;
  cudaGetDeviceProperties(&prop, 0);
  const int sms = prop.multiProcessorCount;
  float *out, *mats;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 3 * 768);
  cudaMalloc(&mats, sizeof(float) * 32);
  cudaMalloc(&cyc, sizeof(long long) * sms * 3);
  std::vector<float> hm(32);
  for (int i = 0; i < 32; ++i) hm[i] = (i % 8 == 0 || i % 8 == 6) ? 0.7071f : ((i % 8 == 2 || i % 8 == 4) ? 0.5f : 0.001f * i);
  cudaMemcpy(mats, hm.data(), sizeof(float) * 32, cudaMemcpyHostToDevice);
  printf("%s, %d SMs\n", prop.name, sms);
  const int stages = 2000;
  report<3, 1>("forward stage  (3 2x2s, 1 round trip)", sms, stages, out, mats, cyc);
  report<4, 1>("forward stage  (4 2x2s, 1 round trip)", sms, stages, out, mats, cyc);
  report<2, 1>("forward stage  (2 2x2s, 1 round trip)", sms, stages, out, mats, cyc);
  report<8, 2>("adjoint-like   (8 2x2s, 2 round trips)", sms, stages, out, mats, cyc);
  report<8, 3>("adjoint-like   (8 2x2s, 3 round trips)", sms, stages, out, mats, cyc);
  printf("stage-start overheads (3 2x2s, 3 CTAs x 256 threads; cycles per tile-stage per SM):\n");
  printf("  plain                         %7.1f\n", run_stage<3, false, false, false>(sms, stages, out, mats, cyc));
  printf("  + tables                      %7.1f\n", run_stage<3, true, false, false>(sms, stages, out, mats, cyc));
  printf("  + second barrier              %7.1f\n", run_stage<3, false, true, false>(sms, stages, out, mats, cyc));
  printf("  + shape switch                %7.1f\n", run_stage<3, false, false, true>(sms, stages, out, mats, cyc));
  printf("  + tables + switch             %7.1f\n", run_stage<3, true, false, true>(sms, stages, out, mats, cyc));
  printf("  + tables + barrier + switch   %7.1f\n", run_stage<3, true, true, true>(sms, stages, out, mats, cyc));
  return 0;
}
