// Signal-level kernels: first / last layers of the networks, enhance() pre/post-processing and
// layout converters (see include/ou_b200.h for the contracts and the reference lines replaced).
#include "common.cuh"

namespace ou {

// ------------------------------------------------------------------------------ input conv
// One thread per time step: reads k input samples, writes cout channels as 32-byte vectors.  The
// weights sit in shared memory tap-major ([k][cout], then the bias) so that a 16-byte broadcast load
// feeds four FMAs: with one 4-byte load per FMA the kernel was bound by shared-memory instruction
// issue (96 loads per thread at C = 32), not by its 262 MB of output.
template <int K>
__global__ void __launch_bounds__(256) input_conv_kernel(const float* __restrict__ x, const float* __restrict__ w,
                                                         const float* __restrict__ bias,
                                                         const float* __restrict__ in_scale,
                                                         act_t* __restrict__ out, int t_len, int cout) {
  extern __shared__ __align__(16) float sw[];  // [K][cout] weights then [cout] bias
  for (int i = threadIdx.x; i < cout * K; i += blockDim.x) sw[(i % K) * cout + i / K] = w[i];
  for (int i = threadIdx.x; i < cout; i += blockDim.x) sw[cout * K + i] = bias ? bias[i] : 0.f;
  __syncthreads();
  const int b = blockIdx.y;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= t_len) return;
  const float sc = in_scale ? in_scale[b] : 1.f;
  float xin[K];
#pragma unroll
  for (int i = 0; i < K; i++) {
    const int tt = t + i - K / 2;
    xin[i] = (tt >= 0 && tt < t_len) ? __ldg(x + (size_t)b * t_len + tt) * sc : 0.f;
  }
  const int cb = cl_cb(cout);
  for (int c16 = 0; c16 < cout / 16; c16++) {   // 16 channels = one 32-byte (256-bit) store
    float4 acc[4];
#pragma unroll
    for (int q = 0; q < 4; q++) acc[q] = *reinterpret_cast<const float4*>(sw + cout * K + c16 * 16 + q * 4);
#pragma unroll
    for (int i = 0; i < K; i++) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const float4 wv = *reinterpret_cast<const float4*>(sw + i * cout + c16 * 16 + q * 4);
        acc[q].x = fmaf(wv.x, xin[i], acc[q].x), acc[q].y = fmaf(wv.y, xin[i], acc[q].y);
        acc[q].z = fmaf(wv.z, xin[i], acc[q].z), acc[q].w = fmaf(wv.w, xin[i], acc[q].w);
      }
    }
    act_t* dst = out + cl_off(b, c16 * 16, t, cout, t_len, cb);
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(dst), "r"(f2_to_act2(acc[0].x, acc[0].y)),
                 "r"(f2_to_act2(acc[0].z, acc[0].w)), "r"(f2_to_act2(acc[1].x, acc[1].y)),
                 "r"(f2_to_act2(acc[1].z, acc[1].w)), "r"(f2_to_act2(acc[2].x, acc[2].y)),
                 "r"(f2_to_act2(acc[2].z, acc[2].w)), "r"(f2_to_act2(acc[3].x, acc[3].y)),
                 "r"(f2_to_act2(acc[3].z, acc[3].w))
                 : "memory");
  }
}

// ------------------------------------------------------------------------------ output conv + SDE
// One lane per time step of the SOURCE row it loads (all cin channels, 32-byte loads); it forms the K
// per-tap partial sums p_i[t] = sum_c w[c][i] * v[c][t] of its own row and the output
// net[t] = bias + sum_i p_i[t + i - K/2] is assembled with warp shuffles, so every activation row is
// read from memory once (a warp covers 32 rows and produces 32 - (K - 1) outputs; the first version
// re-read each row K times through L1 and ran at 2.7 TB/s).  Weights tap-major in shared memory as in
// input_conv_kernel.
template <int K>
__global__ void __launch_bounds__(256) output_sde_kernel(const act_t* __restrict__ src,
                                                         const float* __restrict__ w, float bias,
                                                         const float* __restrict__ coef,
                                                         const float* __restrict__ x,
                                                         const float* __restrict__ noise, float* __restrict__ xout,
                                                         float* __restrict__ net_out, int cin, int t_src, int t_sig) {
  extern __shared__ __align__(16) float sw[];  // [K][cin]
  for (int i = threadIdx.x; i < cin * K; i += blockDim.x) sw[(i % K) * cin + i / K] = w[i];
  __syncthreads();
  constexpr int HALO = K / 2, OUT_PER_WARP = 32 - 2 * HALO;
  const int b = blockIdx.y;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int t = (blockIdx.x * (int)(blockDim.x >> 5) + warp) * OUT_PER_WARP + lane - HALO;   // row this lane loads
  float p[K];
#pragma unroll
  for (int i = 0; i < K; i++) p[i] = 0.f;
  if (t >= 0 && t < t_src) {
    const int cb = cl_cb(cin);
    for (int c16 = 0; c16 < cin / 16; c16++) {
      uint32_t pv[8];   // 16 channels = one 32-byte (256-bit) load
      asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                   : "=r"(pv[0]), "=r"(pv[1]), "=r"(pv[2]), "=r"(pv[3]), "=r"(pv[4]), "=r"(pv[5]),
                     "=r"(pv[6]), "=r"(pv[7])
                   : "l"(src + cl_off(b, c16 * 16, t, cin, t_src, cb)));
      float f[16];
#pragma unroll
      for (int q = 0; q < 8; q++) {
        const float2 v = act2_to_f2(pv[q]);
        f[2 * q] = v.x, f[2 * q + 1] = v.y;
      }
#pragma unroll
      for (int i = 0; i < K; i++) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
          const float4 wv = *reinterpret_cast<const float4*>(sw + i * cin + c16 * 16 + q * 4);
          p[i] = fmaf(wv.x, f[4 * q], p[i]);
          p[i] = fmaf(wv.y, f[4 * q + 1], p[i]);
          p[i] = fmaf(wv.z, f[4 * q + 2], p[i]);
          p[i] = fmaf(wv.w, f[4 * q + 3], p[i]);
        }
      }
    }
  }
  // output at this lane's time step: tap i multiplies the row at t + i - HALO, i.e. lane + i - HALO
  float net = bias;
#pragma unroll
  for (int i = 0; i < K; i++) net += __shfl_sync(0xffffffffu, p[i], (lane + i - HALO) & 31);
  if (lane < HALO || lane >= 32 - HALO || t >= t_sig) return;   // halo lanes only feed their neighbours
  if (t >= t_src) net = 0.f;                                     // right padding of the signal
  const size_t o = (size_t)b * t_sig + t;
  if (net_out) net_out[o] = net;
  if (coef) {
    const float ca = coef[b * 3], cb = coef[b * 3 + 1], cc = coef[b * 3 + 2];
    float v = ca * x[o] + cb * net;
    if (noise) v = fmaf(cc, noise[o], v);
    xout[o] = v;
  }
}

// ------------------------------------------------------------------------------ pad + normalise
// One CTA per clip.  Mean and unbiased std over the padded clip, accumulated in double.
__device__ double block_sum(double v, double* red) {
  v = warp_sum(v);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    double s = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    s = warp_sum(s);
    if (lane == 0) red[0] = s;
  }
  __syncthreads();
  return red[0];
}

__global__ void pad_normalize_kernel(const float* __restrict__ mix, float* __restrict__ out,
                                     float* __restrict__ stats, int t, int t_pad, int pad_left,
                                     float level) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  const float* src = mix + (size_t)b * t;
  double s = 0.0;
  for (int i = threadIdx.x; i < t; i += blockDim.x) s += src[i];
  const double mean = block_sum(s, red) / t_pad;
  const float meanf = (float)mean;
  double q = 0.0;
  for (int i = threadIdx.x; i < t; i += blockDim.x) {
    const double d = (double)(src[i] - meanf);
    q += d * d;
  }
  q = block_sum(q, red);
  q += (double)(t_pad - t) * (double)meanf * (double)meanf;
  const float stdv = (float)sqrt(q / (t_pad - 1));
  const float gain = level / fmaxf(stdv, 1e-5f);
  float* dst = out + (size_t)b * t_pad;
  for (int i = threadIdx.x; i < t_pad; i += blockDim.x) {
    const int j = i - pad_left;
    const float v = (j >= 0 && j < t) ? src[j] : 0.f;
    dst[i] = (v - meanf) * gain;
  }
  if (threadIdx.x == 0 && stats) {
    stats[b * 2] = meanf;
    stats[b * 2 + 1] = 1.f / gain;
  }
}

// ------------------------------------------------------------------------------ unpad + limiter
__global__ void unpad_limit_kernel(const float* __restrict__ x, const float* __restrict__ mix_rms,
                                   float* __restrict__ out, int t_pad, int pad_left, int t_valid,
                                   int t) {
  __shared__ double red[32];
  __shared__ float redf[32];
  const int b = blockIdx.x;
  const float* src = x + (size_t)b * t_pad + pad_left;
  float* dst = out + (size_t)b * t;
  float rs = 1.f;
  if (mix_rms) {
    double q = 0.0;
    for (int i = threadIdx.x; i < t_valid && i < t; i += blockDim.x) q += (double)src[i] * src[i];
    q = block_sum(q, red);
    const float x_rms = fmaxf((float)sqrt(q / t), 1e-5f);
    rs = mix_rms[b] / x_rms;
  }
  float m = 0.f;
  for (int i = threadIdx.x; i < t_valid && i < t; i += blockDim.x) m = fmaxf(m, fabsf(src[i] * rs));
  m = warp_max(m);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) redf[warp] = m;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? redf[lane] : 0.f;
    v = warp_max(v);
    if (lane == 0) redf[0] = v;
  }
  __syncthreads();
  const float peak = redf[0];
  for (int i = threadIdx.x; i < t; i += blockDim.x) {
    float v = (i < t_valid) ? src[i] * rs : 0.f;
    if (peak > 1.f) v = v / peak;
    dst[i] = v;
  }
}

// ------------------------------------------------------------------------------ layout converters
__global__ void pack_blocked_kernel(const float* __restrict__ src, act_t* __restrict__ dst,
                                    int channels, int t_len) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y, b = blockIdx.z;
  if (t >= t_len) return;
  float f[8];
#pragma unroll
  for (int e = 0; e < 8; e++) f[e] = src[((size_t)b * channels + c8 * 8 + e) * t_len + t];
  uint4 v = make_uint4(f2_to_act2(f[0], f[1]), f2_to_act2(f[2], f[3]), f2_to_act2(f[4], f[5]),
                       f2_to_act2(f[6], f[7]));
  *reinterpret_cast<uint4*>(dst + cl_off(b, c8 * 8, t, channels, t_len, cl_cb(channels))) = v;
}

__global__ void unpack_blocked_kernel(const act_t* __restrict__ src, float* __restrict__ dst,
                                      int channels, int t_len) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c8 = blockIdx.y, b = blockIdx.z;
  if (t >= t_len) return;
  const uint4 v =
      *reinterpret_cast<const uint4*>(src + cl_off(b, c8 * 8, t, channels, t_len, cl_cb(channels)));
  const uint32_t* pv = reinterpret_cast<const uint32_t*>(&v);
#pragma unroll
  for (int h = 0; h < 4; h++) {
    const float2 f = act2_to_f2(pv[h]);
    dst[((size_t)b * channels + c8 * 8 + h * 2) * t_len + t] = f.x;
    dst[((size_t)b * channels + c8 * 8 + h * 2 + 1) * t_len + t] = f.y;
  }
}

__global__ void film_f32_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                float* __restrict__ out, int channels, int t_len) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int c = blockIdx.y, b = blockIdx.z;
  if (t >= t_len) return;
  const float g = y[(size_t)b * 2 * channels + c], be = y[(size_t)b * 2 * channels + channels + c];
  const size_t o = ((size_t)b * channels + c) * t_len + t;
  out[o] = g * x[o] + be;
}

}  // namespace ou

extern "C" int ou_input_conv(const float* x, const float* w, const float* bias, const float* in_scale,
                             void* out, int batch, int t, int cout, int k, void* stream) {
  OU_REQUIRE(x && w && out, "ou_input_conv: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && cout > 0 && cout % 16 == 0, "ou_input_conv: bad shape");
  OU_REQUIRE(k >= 1 && k <= 7 && (k & 1), "ou_input_conv: kernel size must be odd and <= 7");
  dim3 grid(ou::ceil_div(t, 256), batch);
  const size_t smem = (size_t)(cout * k + cout) * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  act_t* o = (act_t*)out;
  switch (k) {
    case 1: ou::input_conv_kernel<1><<<grid, 256, smem, st>>>(x, w, bias, in_scale, o, t, cout); break;
    case 3: ou::input_conv_kernel<3><<<grid, 256, smem, st>>>(x, w, bias, in_scale, o, t, cout); break;
    case 5: ou::input_conv_kernel<5><<<grid, 256, smem, st>>>(x, w, bias, in_scale, o, t, cout); break;
    default: ou::input_conv_kernel<7><<<grid, 256, smem, st>>>(x, w, bias, in_scale, o, t, cout); break;
  }
  return ou::check_launch("ou_input_conv");
}

extern "C" int ou_output_sde(const void* src, const float* w, float bias, const float* coef,
                             const float* x, const float* noise, float* xout, float* net_out,
                             int batch, int cin, int k, int t_src, int t_sig, void* stream) {
  OU_REQUIRE(src && w, "ou_output_sde: null pointer");
  OU_REQUIRE(batch > 0 && t_src > 0 && t_sig >= t_src && cin % 16 == 0, "ou_output_sde: bad shape");
  OU_REQUIRE(k >= 1 && (k & 1) && k <= 7, "ou_output_sde: kernel size must be odd and <= 7");
  OU_REQUIRE(coef == nullptr || (x && xout), "ou_output_sde: coef needs x and xout");
  OU_REQUIRE(coef || net_out, "ou_output_sde: nothing to write");
  dim3 grid(ou::ceil_div(t_sig, 8 * (32 - 2 * (k / 2))), batch);   // 8 warps x (32 - halo) outputs per CTA
  const size_t smem = (size_t)cin * k * sizeof(float);
  cudaStream_t st = (cudaStream_t)stream;
  const act_t* s = (const act_t*)src;
  switch (k) {
    case 1: ou::output_sde_kernel<1><<<grid, 256, smem, st>>>(s, w, bias, coef, x, noise, xout, net_out, cin, t_src, t_sig); break;
    case 3: ou::output_sde_kernel<3><<<grid, 256, smem, st>>>(s, w, bias, coef, x, noise, xout, net_out, cin, t_src, t_sig); break;
    case 5: ou::output_sde_kernel<5><<<grid, 256, smem, st>>>(s, w, bias, coef, x, noise, xout, net_out, cin, t_src, t_sig); break;
    default: ou::output_sde_kernel<7><<<grid, 256, smem, st>>>(s, w, bias, coef, x, noise, xout, net_out, cin, t_src, t_sig); break;
  }
  return ou::check_launch("ou_output_sde");
}

extern "C" int ou_pad_normalize(const float* mix, float* out, float* stats, int batch, int t,
                                int t_pad, int pad_left, float level, void* stream) {
  OU_REQUIRE(mix && out, "ou_pad_normalize: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && t_pad >= t + pad_left && t_pad > 1 && pad_left >= 0,
             "ou_pad_normalize: bad shape");
  ou::pad_normalize_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(mix, out, stats, t, t_pad,
                                                                     pad_left, level);
  return ou::check_launch("ou_pad_normalize");
}

extern "C" int ou_unpad_limit(const float* x, const float* mix_rms, float* out, int batch, int t_pad,
                              int pad_left, int t_valid, int t, void* stream) {
  OU_REQUIRE(x && out, "ou_unpad_limit: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && t_valid > 0 && pad_left + t_valid <= t_pad,
             "ou_unpad_limit: bad shape");
  ou::unpad_limit_kernel<<<batch, 1024, 0, (cudaStream_t)stream>>>(x, mix_rms, out, t_pad, pad_left,
                                                                   t_valid, t);
  return ou::check_launch("ou_unpad_limit");
}

extern "C" int ou_pack_blocked(const float* src, void* dst, int batch, int channels, int t,
                               void* stream) {
  OU_REQUIRE(src && dst, "ou_pack_blocked: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && channels > 0 && channels % 16 == 0, "ou_pack_blocked: bad shape");
  dim3 grid(ou::ceil_div(t, 256), channels / 8, batch);
  ou::pack_blocked_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(src, (act_t*)dst, channels, t);
  return ou::check_launch("ou_pack_blocked");
}

extern "C" int ou_unpack_blocked(const void* src, float* dst, int batch, int channels, int t,
                                 void* stream) {
  OU_REQUIRE(src && dst, "ou_unpack_blocked: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && channels > 0 && channels % 16 == 0, "ou_unpack_blocked: bad shape");
  dim3 grid(ou::ceil_div(t, 256), channels / 8, batch);
  ou::unpack_blocked_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>((const act_t*)src, dst,
                                                                    channels, t);
  return ou::check_launch("ou_unpack_blocked");
}

extern "C" int ou_film_f32(const float* x, const float* y, float* out, int batch, int channels, int t,
                           void* stream) {
  OU_REQUIRE(x && y && out, "ou_film_f32: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && channels > 0, "ou_film_f32: bad shape");
  dim3 grid(ou::ceil_div(t, 256), channels, batch);
  ou::film_f32_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(x, y, out, channels, t);
  return ou::check_launch("ou_film_f32");
}
