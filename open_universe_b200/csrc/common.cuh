// Shared helpers for libou_b200.so (sm_100a only).
#pragma once
#include <cuda_bf16.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ou_b200.h"

namespace ou {

void set_error(const char* fmt, ...);
extern std::atomic<int64_t> g_launches;
extern std::atomic<int64_t> g_conv_fallbacks;

inline int check_launch(const char* what) {
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) {
    set_error("%s: %s", what, cudaGetErrorString(e));
    return OU_ERR_CUDA;
  }
  g_launches.fetch_add(1, std::memory_order_relaxed);
  return OU_OK;
}

// Per-device caches (a process may drive several GPUs: cudaFuncSetAttribute and the SM count are
// per device, so a per-process `static` would leave device 1 unconfigured after device 0 ran).
constexpr int OU_MAX_DEVICES = 64;
inline int current_device() {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= OU_MAX_DEVICES) dev = 0;
  return dev;
}
inline int num_sms() {
  static std::atomic<int> cache[OU_MAX_DEVICES];
  const int dev = current_device();
  int n = cache[dev].load(std::memory_order_relaxed);
  if (n == 0) {
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cache[dev].store(n, std::memory_order_relaxed);
  }
  return n;
}
struct SmemConfig {
  std::atomic<size_t> configured[OU_MAX_DEVICES];
};
// raise the dynamic shared-memory limit of `kern` on the current device to at least `smem` bytes
template <typename K>
inline int ensure_smem(K kern, size_t smem, SmemConfig& cfg, const char* what) {
  const int dev = current_device();
  if (smem <= cfg.configured[dev].load(std::memory_order_relaxed)) return OU_OK;
  cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  if (e != cudaSuccess) {
    set_error("%s: cudaFuncSetAttribute(%zu): %s", what, smem, cudaGetErrorString(e));
    return OU_ERR_CUDA;
  }
  cfg.configured[dev].store(smem, std::memory_order_relaxed);
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

// fp32 blocked layout [B][N/16][rows][16] (GRU input pre-activations): element offset of column n, row j
__host__ __device__ inline size_t f32blk_off(int b, int n, int j, int n_total, int rows) {
  return (((size_t)b * (n_total >> 4) + (n >> 4)) * rows + j) * 16 + (n & 15);
}

__device__ __forceinline__ float prelu_f(float x, float a) { return x >= 0.f ? x : a * x; }

// ---- storage / tensor-core operand type of every activation and conv weight ("act") ----
// Default: IEEE fp16 -- the 11-bit significand of the TF32 the reference's own CUDA path computes its
// convolutions in (cuDNN, allow_tf32), 8x finer than bf16 at the same tcgen05 kind::f16 throughput and
// the same bytes.  Measured (tools/precision_study.py, profiles/README.md): one score-network
// evaluation deviates 1e-3..5e-3 relative from the fp32 oracle with fp16, 0.9e-2..4e-2 with bf16.
// Conversions to fp16 saturate (cvt.rn.satfinite) instead of producing inf.  -DOU_ACT_BF16 builds the
// bf16 policy for A/B runs.
#ifdef OU_ACT_BF16
typedef __nv_bfloat16 act_t;
#define OU_ACT_IS_BF16 1
#define OU_ACT_PTX "bf16x2"
__device__ __forceinline__ float act_to_f(act_t v) { return __bfloat162float(v); }
__device__ __forceinline__ act_t f_to_act(float v) { return __float2bfloat16(v); }
__device__ __forceinline__ float2 act2_to_f2(uint32_t v) {
  __nv_bfloat162 h = *reinterpret_cast<__nv_bfloat162*>(&v);
  return __bfloat1622float2(h);
}
__device__ __forceinline__ uint32_t f2_to_act2(float a, float b) {
  __nv_bfloat162 h = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&h);
}
#else
typedef __half act_t;
#define OU_ACT_IS_BF16 0
#define OU_ACT_PTX "f16x2"
__device__ __forceinline__ float act_to_f(act_t v) { return __half2float(v); }
__device__ __forceinline__ float2 act2_to_f2(uint32_t v) {
  __half2 h = *reinterpret_cast<__half2*>(&v);
  return __half22float2(h);
}
__device__ __forceinline__ uint32_t f2_to_act2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));
  return r;
}
__device__ __forceinline__ act_t f_to_act(float v) {
  return __ushort_as_half((unsigned short)(f2_to_act2(v, 0.f) & 0xFFFFu));
}
#endif
__device__ __forceinline__ unsigned short act_bits(act_t v) { return *reinterpret_cast<unsigned short*>(&v); }
__device__ __forceinline__ float act_bits_to_f(unsigned short b) {
  act_t v = *reinterpret_cast<act_t*>(&b);
  return act_to_f(v);
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

using ou::act_t;   // the extern "C" entry points live outside the namespace
