// Shared helpers for libou_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ou_b200.h"

namespace ou {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return OU_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OU_OK;
}

#define OU_REQUIRE(cond, ...)        \
  do {                               \
    if (!(cond)) {                   \
      ou::set_error(__VA_ARGS__);    \
      return OU_ERR_INVALID;         \
    }                                \
  } while (0)

__host__ __device__ inline int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Blocked ("channels-last in blocks") activation layout: bf16 [B][C/CB][T][CB] with
// CB = largest of {64, 32, 16} dividing C.  One time step of one channel block is a CB*2-byte
// row (128 B for CB = 64) -- exactly a K-major, 128B-swizzle-able tensor-core operand row.
__host__ __device__ inline int cl_cb(int channels) {
  return (channels % 64 == 0) ? 64 : ((channels % 32 == 0) ? 32 : 16);
}
// element offset of channel c (any) of time step t of clip b
__host__ __device__ inline size_t cl_off(int b, int c, int t, int channels, int t_len, int cb) {
  return (((size_t)b * (channels / cb) + c / cb) * t_len + t) * cb + (c % cb);
}

__device__ __forceinline__ float prelu_f(float x, float a) { return x >= 0.f ? x : a * x; }

__device__ __forceinline__ float2 bf2_to_f2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t f2_to_bf2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

}  // namespace ou
