// Host-side stand-in for <cuda_runtime.h>, used ONLY by tests/kernel_emu: it lets g++ compile the unmodified CUDA kernel
// sources of qandle_b200/csrc as ordinary C++ and run them on the CPU as a FUNCTIONAL SIMULATION -- one cooperative fiber
// per CUDA thread, CTA / named / warp barriers and warp shuffles implemented by the fiber scheduler (kernel_emu.cpp).
// TEST INFRASTRUCTURE: nothing in the product loads or links this; there is no CPU execution path in qandle_b200.
#pragma once
#include <cmath>
#include <cstddef>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>

#define QB_KERNEL_EMU 1
#define __device__
#define __host__
#define __global__
#define __forceinline__ inline
#define __noinline__  // (also spelled inside __attribute__((...)) by system headers: must expand to nothing)
#define __grid_constant__
#define __launch_bounds__(...)
#define __align__(n) alignas(n)
#define __shared__ static  // static shared arrays: one CTA runs at a time on one OS thread

// ---- vector types ---------------------------------------------------------------------------------------------
struct alignas(8) float2 { float x, y; };
struct alignas(16) float4 { float x, y, z, w; };
struct alignas(16) double2 { double x, y; };
struct alignas(8) int2 { int x, y; };
struct alignas(16) int4 { int x, y, z, w; };
struct alignas(8) uint2 { unsigned x, y; };
struct uint3 { unsigned x, y, z; };
struct alignas(16) uint4 { unsigned x, y, z, w; };
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned a = 1, unsigned b = 1, unsigned c = 1) : x(a), y(b), z(c) {}
};
inline float2 make_float2(float x, float y) { return float2{x, y}; }
inline float4 make_float4(float x, float y, float z, float w) { return float4{x, y, z, w}; }
inline double2 make_double2(double x, double y) { return double2{x, y}; }
inline uint2 make_uint2(unsigned x, unsigned y) { return uint2{x, y}; }
inline uint4 make_uint4(unsigned x, unsigned y, unsigned z, unsigned w) { return uint4{x, y, z, w}; }

// ---- runtime API subset (device memory is host memory) ---------------------------------------------------------------
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0 };
enum cudaMemcpyKind { cudaMemcpyHostToHost, cudaMemcpyHostToDevice, cudaMemcpyDeviceToHost, cudaMemcpyDeviceToDevice, cudaMemcpyDefault };
enum { cudaFuncAttributeMaxDynamicSharedMemorySize = 8 };
enum { cudaStreamNonBlocking = 1 };
struct cudaDeviceProp {
  int multiProcessorCount = 148;
  size_t sharedMemPerBlockOptin = 227 * 1024;
  size_t totalGlobalMem = size_t(180) << 30;
  int major = 10, minor = 0;
  char name[64] = "kernel_emu (CPU fibers)";
};
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = std::aligned_alloc(256, (n + 255) / 256 * 256 + 256); return *p ? 0 : 2; }
template <class T> inline cudaError_t cudaMalloc(T** p, size_t n) { return cudaMalloc(reinterpret_cast<void**>(p), n); }
inline cudaError_t cudaFree(void* p) { std::free(p); return 0; }
inline cudaError_t cudaMemcpy(void* d, const void* s, size_t n, cudaMemcpyKind) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, cudaMemcpyKind, cudaStream_t = nullptr) { std::memmove(d, s, n); return 0; }
inline cudaError_t cudaMemsetAsync(void* d, int v, size_t n, cudaStream_t = nullptr) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaMemset(void* d, int v, size_t n) { std::memset(d, v, n); return 0; }
inline cudaError_t cudaGetLastError() { return 0; }
inline const char* cudaGetErrorString(cudaError_t) { return "kernel_emu error"; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return 0; }
inline cudaError_t cudaDeviceSynchronize() { return 0; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, unsigned) { *s = nullptr; return 0; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return 0; }
inline cudaError_t cudaGetDevice(int* d) { *d = 0; return 0; }
inline cudaError_t cudaGetDeviceProperties(cudaDeviceProp* p, int) { *p = cudaDeviceProp(); return 0; }
template <class F> inline cudaError_t cudaFuncSetAttribute(F, int, int) { return 0; }
struct cudaFuncAttributes { size_t sharedSizeBytes = 0; int numRegs = 0; };
template <class F> inline cudaError_t cudaFuncGetAttributes(cudaFuncAttributes* a, F) { *a = cudaFuncAttributes(); return 0; }

// ---- execution model ------------------------------------------------------------------------------------------------
namespace kemu {
struct ThreadCtx {
  uint3 tid, bid;
  dim3 bdim, gdim;
};
ThreadCtx& cur();                                    // the running fiber's indices
unsigned char* dyn_smem();                           // dynamic shared memory of the running CTA (16-byte aligned)
void barrier(int id, int count);                     // CTA barrier `id`; count <= 0: all live threads of the CTA
void warp_barrier();                                 // live lanes of the running fiber's warp
uint64_t shfl_xor_bits(uint64_t v, int lane_mask);   // exchange 64 bits with lane ^ lane_mask
void launch(dim3 grid, dim3 block, size_t smem_bytes, const std::function<void()>& body);
}  // namespace kemu

#define threadIdx (kemu::cur().tid)
#define blockIdx (kemu::cur().bid)
#define blockDim (kemu::cur().bdim)
#define gridDim (kemu::cur().gdim)

inline void __syncthreads() { kemu::barrier(0, 0); }
inline void __syncwarp(unsigned = 0xffffffffu) { kemu::warp_barrier(); }
template <class T> inline T __shfl_xor_sync(unsigned, T v, int lane_mask) {
  static_assert(sizeof(T) <= 8, "shuffle operand");
  uint64_t b = 0;
  std::memcpy(&b, &v, sizeof(T));
  b = kemu::shfl_xor_bits(b, lane_mask);
  T r;
  std::memcpy(&r, &b, sizeof(T));
  return r;
}
inline unsigned __float_as_uint(float f) { unsigned u; std::memcpy(&u, &f, 4); return u; }
inline float __uint_as_float(unsigned u) { float f; std::memcpy(&f, &u, 4); return f; }
inline float2 __ffma2_rn(float2 a, float2 b, float2 c) { return float2{std::fma(a.x, b.x, c.x), std::fma(a.y, b.y, c.y)}; }
inline float2 __fmul2_rn(float2 a, float2 b) { return float2{a.x * b.x, a.y * b.y}; }
template <class T> inline void __stcs(T* p, T v) { *p = v; }
template <class T> inline T __ldcs(const T* p) { return *p; }
template <class T> inline T __ldg(const T* p) { return *p; }
inline int __double2hiint(double d) { int64_t b; std::memcpy(&b, &d, 8); return (int)(b >> 32); }
inline int __double2loint(double d) { int64_t b; std::memcpy(&b, &d, 8); return (int)(b & 0xffffffff); }
inline double __hiloint2double(int hi, int lo) { const uint64_t b = ((uint64_t)(uint32_t)hi << 32) | (uint32_t)lo; double d; std::memcpy(&d, &b, 8); return d; }
inline long long __double_as_longlong(double d) { long long b; std::memcpy(&b, &d, 8); return b; }
inline double __longlong_as_double(long long b) { double d; std::memcpy(&d, &b, 8); return d; }
inline int __clz(int x) { return x == 0 ? 32 : __builtin_clz((unsigned)x); }
inline int __popc(unsigned x) { return __builtin_popcount(x); }
inline int __popcll(unsigned long long x) { return __builtin_popcountll(x); }
inline int __ffsll(long long x) { return __builtin_ffsll(x); }
template <class T> inline T atomicAdd(T* p, T v) { T o = *p; *p = o + v; return o; }  // one OS thread: fibers never pre-empt
using std::fma;
using std::fmaf;
inline int min(int a, int b) { return a < b ? a : b; }
inline int max(int a, int b) { return a > b ? a : b; }
inline unsigned min(unsigned a, unsigned b) { return a < b ? a : b; }
inline unsigned max(unsigned a, unsigned b) { return a > b ? a : b; }
inline uint64_t min(uint64_t a, uint64_t b) { return a < b ? a : b; }
inline uint64_t max(uint64_t a, uint64_t b) { return a > b ? a : b; }
inline int64_t min(int64_t a, int64_t b) { return a < b ? a : b; }
inline int64_t max(int64_t a, int64_t b) { return a > b ? a : b; }
// shared-window addresses (cp.async destinations): offset into the CTA's dynamic shared memory
inline size_t __cvta_generic_to_shared(const void* p) { return (size_t)(reinterpret_cast<const unsigned char*>(p) - kemu::dyn_smem()); }
namespace kemu {
inline void cp_async16(uint32_t smem_off, const void* gsrc) { std::memcpy(dyn_smem() + smem_off, gsrc, 16); }
}
