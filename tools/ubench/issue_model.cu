// Micro-benchmark for the sweep kernels' performance model: how do FFMA2 and the "other" instructions of
// a stage (LOP3 / IADD3 address arithmetic, LDS.128) share an SM sub-partition's issue port on B200?
//   build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/ubench/issue_model tools/ubench/issue_model.cu
//   run:   tools/ubench/issue_model            (prints warp-instructions per cycle per SM sub-partition)
// Each warp runs an unrolled loop of K independent FFMA2 chains, optionally interleaved with X independent integer ops
// or shared-memory loads per FFMA2.  If the extra instructions are free while FFMA2 (2 cycles per issue on the FMA pipe)
// is the bottleneck, the FFMA2 rate stays at 0.5 / cycle; if they take issue slots away, it drops.
#include <cuda_runtime.h>

#include <cstdio>
#include <vector>

template <int MODE, int XPER>  // MODE 0: FFMA2 only; 1: + XPER LOP3 per FFMA2; 2: + 1 LDS.128 per (XPER) FFMA2; 3: scalar FFMA; 4: LOP3 only;
                               // 5: FFMA2 with a scalar-broadcast operand (SASS `Rn.F32`: the form the sweep kernels' 2x2s use)
__global__ void __launch_bounds__(1024, 1) k_issue(float* out, long long* cyc, int iters) {
  __shared__ float4 sm[1024];
  sm[threadIdx.x] = make_float4(threadIdx.x, 1.f, 2.f, 3.f);
  __syncthreads();
  float2 a[8];
  unsigned u[8];
  unsigned ld = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    a[i] = make_float2(1.0f + threadIdx.x * 1e-3f + i, 0.5f + i);
    u[i] = threadIdx.x * 2654435761u + i;
  }
  const float2 m = make_float2(0.999f, 1.001f), c = make_float2(1e-3f, -1e-3f);
  const float ms = 0.999f + 1e-6f * (float)(iters & 3);  // scalar operand, unknown at compile time
  const unsigned k1 = 0x9E3779B9u + blockIdx.x;
  unsigned addr = (threadIdx.x * 16) & 16383;
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int r = 0; r < 8; ++r) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        if (MODE == 0 || MODE == 1 || MODE == 2) a[i] = __ffma2_rn(a[i], m, c);
        if (MODE == 5) a[i] = __ffma2_rn(make_float2(ms, ms), a[i], c);
        if (MODE == 3) {
          a[i].x = fmaf(a[i].x, m.x, c.x);
          a[i].y = fmaf(a[i].y, m.y, c.y);
        }
        if (MODE == 1 || MODE == 4) {
#pragma unroll
          for (int x = 0; x < XPER; ++x) u[(i + x) & 7] = __funnelshift_l(u[(i + x) & 7], u[(i + x + 1) & 7], 7) ^ k1;  // SHF + LOP3 (both ALU pipe)
        }
        if (MODE == 2 && (i % XPER) == 0) {
          const float4 v = *reinterpret_cast<const float4*>(reinterpret_cast<const char*>(sm) + addr);
          ld ^= __float_as_uint(v.x);
          addr = (addr + 16 * 33) & 16383;
        }
      }
    }
  }
  const long long t1 = clock64();
  float s = (float)ld;
  unsigned us = 0;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    s += a[i].x + a[i].y;
    us ^= u[i];
  }
  out[blockIdx.x * blockDim.x + threadIdx.x] = s + (float)us;
  if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}

template <int MODE, int XPER>
void run(const char* name, double fp_per_iter, double other_per_iter) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  float* out;
  long long* cyc;
  cudaMalloc(&out, sizeof(float) * sms * 1024);
  cudaMalloc(&cyc, sizeof(long long) * sms);
  const int iters = 2000;
  for (int warps_per_smsp : {1, 2, 4, 6, 8}) {
    const int threads = warps_per_smsp * 4 * 32;
    k_issue<MODE, XPER><<<sms, threads>>>(out, cyc, 10);
    k_issue<MODE, XPER><<<sms, threads>>>(out, cyc, iters);
    cudaDeviceSynchronize();
    std::vector<long long> h(sms);
    cudaMemcpy(h.data(), cyc, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (long long v : h) avg += (double)v;
    avg /= sms;
    const double fp = fp_per_iter * iters * warps_per_smsp / avg, ot = other_per_iter * iters * warps_per_smsp / avg;
    printf("%-34s warps/SMSP=%d  cycles=%.0f  FP inst/cyc/SMSP=%.3f  other inst/cyc/SMSP=%.3f  total=%.3f\n", name, warps_per_smsp,
           avg, fp, ot, fp + ot);
  }
  cudaFree(out);
  cudaFree(cyc);
}

int main() {
  cudaError_t e = cudaSetDevice(0);
  if (e != cudaSuccess) {
    printf("no CUDA device: %s\n", cudaGetErrorString(e));
    return 1;
  }
  run<0, 0>("FFMA2 only (8 chains)", 64, 0);
  run<5, 0>("FFMA2, scalar-broadcast operand", 64, 0);
  run<3, 0>("scalar FFMA x2 (same flops)", 128, 0);
  run<4, 1>("SHF+LOP3 only", 0, 128);
  run<1, 1>("FFMA2 + (SHF+LOP3) each", 64, 128);
  run<1, 2>("FFMA2 + 2 (SHF+LOP3) each", 64, 256);
  run<2, 2>("FFMA2 + (LDS.128+3 int) per 2 FFMA2", 64, 32 * 4);
  run<2, 4>("FFMA2 + (LDS.128+3 int) per 4 FFMA2", 64, 16 * 4);
  return 0;
}
