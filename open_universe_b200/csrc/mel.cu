// Mel front-end of the conditioner (condition.py:92-108) and the sigma embedding / small dense
// layers (sigma_block.py, FiLM projections of score.py).  All fp32 on CUDA cores: this work is
// O(1 %) of one enhance() call and is precision-sensitive (power spectrum, sinusoid phases).
#include "common.cuh"

namespace ou {

// One CTA per (frame, clip).  Direct real DFT with a twiddle table in shared memory:
//   X[k] = sum_i xw[i] * (cos, -sin)(2*pi*i*k/N),  power = re^2 + im^2,  mel = power @ fb.
__global__ void mel_power_kernel(const float* __restrict__ x, const float* __restrict__ window,
                                 const float* __restrict__ fb, const float* __restrict__ twiddle,
                                 float* __restrict__ mel, float* __restrict__ energy, int t_len,
                                 int n_fft, int hop, int n_mels, int pad_left, int frames) {
  extern __shared__ float sm[];
  float* xw = sm;                         // [n_fft] windowed frame
  float2* tw = reinterpret_cast<float2*>(sm + n_fft);  // [n_fft] (cos, sin)
  float* pw = sm + 3 * n_fft;             // [n_fft/2+1] power spectrum
  __shared__ float red[32];
  const int m = blockIdx.x, b = blockIdx.y;
  const int n_freq = n_fft / 2 + 1;
  const long start = (long)m * hop - pad_left;
  for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
    const long tt = start + i;
    const float v = (tt >= 0 && tt < t_len) ? x[(size_t)b * t_len + tt] : 0.f;
    xw[i] = v * window[i];
    tw[i] = reinterpret_cast<const float2*>(twiddle)[i];
  }
  __syncthreads();
  for (int k = threadIdx.x; k < n_freq; k += blockDim.x) {
    float re = 0.f, im = 0.f;
    int idx = 0;
    for (int i = 0; i < n_fft; i++) {
      const float2 c = tw[idx];
      const float v = xw[i];
      re = fmaf(v, c.x, re);
      im = fmaf(v, c.y, im);
      idx += k;
      if (idx >= n_fft) idx -= n_fft;
    }
    pw[k] = re * re + im * im;
  }
  __syncthreads();
  float e = 0.f;
  for (int j = threadIdx.x; j < n_mels; j += blockDim.x) {
    float acc = 0.f;
    for (int k = 0; k < n_freq; k++) acc = fmaf(pw[k], fb[(size_t)k * n_mels + j], acc);
    mel[((size_t)b * n_mels + j) * frames + m] = acc;
    e += acc * acc;
  }
  e = warp_sum(e);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = e;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) energy[(size_t)b * frames + m] = v;
  }
}

// One CTA per clip: scale = 1 / max(sqrt(mean_frames energy), 1e-5), applied to every mel bin.
__global__ void mel_finalize_kernel(const float* __restrict__ mel, const float* __restrict__ energy,
                                    float* __restrict__ mel_norm, __nv_bfloat16* __restrict__ blocked,
                                    int n_mels, int frames) {
  __shared__ float red[32];
  __shared__ float s_scale;
  const int b = blockIdx.x;
  float e = 0.f;
  for (int i = threadIdx.x; i < frames; i += blockDim.x) e += energy[(size_t)b * frames + i];
  e = warp_sum(e);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = e;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) s_scale = 1.f / fmaxf(sqrtf(v / frames), 1e-5f);
  }
  __syncthreads();
  const float sc = s_scale;
  const float* src = mel + (size_t)b * n_mels * frames;
  for (int i = threadIdx.x; i < n_mels * frames; i += blockDim.x) {
    const float v = src[i] * sc;
    const int j = i / frames, m = i - j * frames;
    if (mel_norm) mel_norm[(size_t)b * n_mels * frames + i] = v;
    if (blocked)
      blocked[cl_off(b, j, m, n_mels, frames, cl_cb(n_mels))] = __float2bfloat16(v);
  }
}

__global__ void sigma_embed_simple_kernel(const float* __restrict__ ls, float weight, float bias,
                                          float* __restrict__ out, int half) {
  const int r = blockIdx.x;
  const float f = 0.5f / (1.f + expf(-(weight * ls[r] + bias)));
  for (int k = threadIdx.x; k < half; k += blockDim.x) {
    const float ph = 2.0f * 3.14159265358979323846f * f * (float)k;
    out[(size_t)r * 2 * half + k] = sinf(ph);
    out[(size_t)r * 2 * half + half + k] = cosf(ph);
  }
}

__global__ void sigma_embed_rff_kernel(const float* __restrict__ ls, const float* __restrict__ freq,
                                       float* __restrict__ out, int n_rff) {
  const int r = blockIdx.x;
  for (int k = threadIdx.x; k < n_rff; k += blockDim.x) {
    const float ph = 2.0f * 3.14159265358979323846f * freq[k] * ls[r];
    out[(size_t)r * 2 * n_rff + k] = sinf(ph);
    out[(size_t)r * 2 * n_rff + n_rff + k] = cosf(ph);
  }
}

// out[r][c] = act(bias[c] + <w[c,:], in[r,:]>): one warp per output element.
__global__ void linear_f32_kernel(const float* __restrict__ in, const float* __restrict__ w,
                                  const float* __restrict__ bias, float* __restrict__ out, int rows,
                                  int k, int n, int ld_out, int has_prelu, float slope) {
  const int warps_per_block = blockDim.x >> 5;
  const long gw = (long)blockIdx.x * warps_per_block + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (gw >= (long)rows * n) return;
  const int r = gw / n, c = gw - (long)r * n;
  float acc = 0.f;
  for (int i = lane; i < k; i += 32) acc = fmaf(w[(size_t)c * k + i], in[(size_t)r * k + i], acc);
  acc = warp_sum(acc);
  if (lane == 0) {
    acc += bias ? bias[c] : 0.f;
    if (has_prelu) acc = prelu_f(acc, slope);
    out[(size_t)r * ld_out + c] = acc;
  }
}

}  // namespace ou

extern "C" int ou_mel_power(const float* x, const float* window, const float* fb,
                            const float* twiddle, float* mel, float* energy, int batch, int t,
                            int n_fft, int hop, int n_mels, int pad_left, int frames, void* stream) {
  OU_REQUIRE(x && window && fb && twiddle && mel && energy, "ou_mel_power: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && n_fft > 0 && (n_fft & 1) == 0 && hop > 0 && n_mels > 0 &&
                 frames > 0,
             "ou_mel_power: bad shape");
  const size_t smem = (size_t)(3 * n_fft + n_fft / 2 + 1) * sizeof(float);
  OU_REQUIRE(smem <= 48 * 1024, "ou_mel_power: n_fft too large for the direct-DFT kernel");
  dim3 grid(frames, batch);
  ou::mel_power_kernel<<<grid, 352, smem, (cudaStream_t)stream>>>(
      x, window, fb, twiddle, mel, energy, t, n_fft, hop, n_mels, pad_left, frames);
  return ou::check_launch("ou_mel_power");
}

extern "C" int ou_mel_finalize(const float* mel, const float* energy, float* mel_norm,
                               void* mel_blocked, int batch, int n_mels, int frames, void* stream) {
  OU_REQUIRE(mel && energy && (mel_norm || mel_blocked), "ou_mel_finalize: null pointer");
  OU_REQUIRE(batch > 0 && n_mels > 0 && frames > 0, "ou_mel_finalize: bad shape");
  OU_REQUIRE(mel_blocked == nullptr || n_mels % 16 == 0, "ou_mel_finalize: n_mels % 16 != 0");
  ou::mel_finalize_kernel<<<batch, 512, 0, (cudaStream_t)stream>>>(
      mel, energy, mel_norm, (__nv_bfloat16*)mel_blocked, n_mels, frames);
  return ou::check_launch("ou_mel_finalize");
}

extern "C" int ou_sigma_embed_simple(const float* log10_sigma, float weight, float bias, float* out,
                                     int rows, int half, void* stream) {
  OU_REQUIRE(log10_sigma && out && rows > 0 && half > 0, "ou_sigma_embed_simple: bad argument");
  ou::sigma_embed_simple_kernel<<<rows, 256, 0, (cudaStream_t)stream>>>(log10_sigma, weight, bias,
                                                                        out, half);
  return ou::check_launch("ou_sigma_embed_simple");
}

extern "C" int ou_sigma_embed_rff(const float* log10_sigma, const float* freq, float* out, int rows,
                                  int n_rff, void* stream) {
  OU_REQUIRE(log10_sigma && freq && out && rows > 0 && n_rff > 0, "ou_sigma_embed_rff: bad argument");
  ou::sigma_embed_rff_kernel<<<rows, 64, 0, (cudaStream_t)stream>>>(log10_sigma, freq, out, n_rff);
  return ou::check_launch("ou_sigma_embed_rff");
}

extern "C" int ou_linear_f32(const float* in, const float* w, const float* bias, float* out, int rows,
                             int k, int n, int ld_out, int has_prelu, float slope, void* stream) {
  OU_REQUIRE(in && w && out && rows > 0 && k > 0 && n > 0 && ld_out >= n, "ou_linear_f32: bad argument");
  const long warps = (long)rows * n;
  const int blocks = (int)((warps + 7) / 8);
  ou::linear_f32_kernel<<<blocks, 256, 0, (cudaStream_t)stream>>>(in, w, bias, out, rows, k, n,
                                                                  ld_out, has_prelu, slope);
  return ou::check_launch("ou_linear_f32");
}
