// Mel front-end of the conditioner (condition.py:92-108) and the sigma embedding / small dense
// layers (sigma_block.py, FiLM projections of score.py).  All fp32 on CUDA cores: this work is
// O(1 %) of one enhance() call and is precision-sensitive (power spectrum, sinusoid phases).
#include "common.cuh"

namespace ou {

// Power spectrum of all frames as ONE fp32 GEMM on the CUDA cores (fp32 is required: the mel energies
// span many orders of magnitude and parity is held at 2e-5):
//   [B * frames, n_fft] (frames gathered from the zero-padded signal, times the window)
//     x [n_fft, 2 * n_freq] (cos | sin interleaved per bin, built once in double precision on the host)
// 64 x 64 output tile per CTA, BK = 16, 4 x 4 micro-tile per thread; a thread owns (re, im) of two bins, so
// the epilogue writes |X|^2 directly.  (The first version computed a direct DFT per frame with a
// table lookup per multiply-add: 3.8 ms per call, 1.2 % of a whole 64-step enhance().)
constexpr int MEL_BM = 64, MEL_BN = 64, MEL_BK = 16;

__global__ void __launch_bounds__(256) dft_power_kernel(const float* __restrict__ x, const float* __restrict__ window,
                                                        const float* __restrict__ dft, float* __restrict__ power,
                                                        int batch, int t_len, int n_fft, int hop, int pad_left,
                                                        int frames) {
  __shared__ __align__(16) float As[MEL_BK][MEL_BM + 4];
  __shared__ __align__(16) float Bs[MEL_BK][MEL_BN + 4];
  const int n_freq = n_fft / 2 + 1, n2 = 2 * n_freq;
  const int rows = batch * frames;
  const int g0 = blockIdx.y * MEL_BM, n0 = blockIdx.x * MEL_BN;
  const int tid = threadIdx.x, ty = tid >> 4, tx = tid & 15;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; i++)
#pragma unroll
    for (int j = 0; j < 4; j++) acc[i][j] = 0.f;
  // A loader: thread -> (row r = tid / 4, four consecutive samples kk0 = (tid % 4) * 4)
  const int ar = tid >> 2, ak = (tid & 3) * 4;
  const int ag = g0 + ar;
  const int ab = ag < rows ? ag / frames : 0;
  const long a_start = ag < rows ? (long)(ag - ab * frames) * hop - pad_left : 0;
  const float* xa = x + (size_t)ab * t_len;
  // B loader: thread -> (k row kk = tid / 16, four consecutive columns (tid % 16) * 4)
  const int bk = tid >> 4, bc = (tid & 15) * 4;
  for (int k0 = 0; k0 < n_fft; k0 += MEL_BK) {
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int i = k0 + ak + e;
      const long tt = a_start + i;
      float v = 0.f;
      if (ag < rows && tt >= 0 && tt < t_len) v = __ldg(xa + tt) * __ldg(window + i);
      As[ak + e][ar] = v;
    }
#pragma unroll
    for (int e = 0; e < 4; e++) {
      const int c = n0 + bc + e;
      Bs[bk][bc + e] = c < n2 ? __ldg(dft + (size_t)(k0 + bk) * n2 + c) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < MEL_BK; kk++) {
      const float4 av = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 bv = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float a4[4] = {av.x, av.y, av.z, av.w}, b4[4] = {bv.x, bv.y, bv.z, bv.w};
#pragma unroll
      for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 4; j++) acc[i][j] = fmaf(a4[i], b4[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; i++) {
    const int g = g0 + ty * 4 + i;
    if (g >= rows) continue;
#pragma unroll
    for (int j = 0; j < 2; j++) {
      const int bin = (n0 + tx * 4) / 2 + j;
      if (bin < n_freq) power[(size_t)g * n_freq + bin] = acc[i][2 * j] * acc[i][2 * j] + acc[i][2 * j + 1] * acc[i][2 * j + 1];
    }
  }
}

// One CTA per (frame, clip): mel = power @ fb, frame energy = sum_j mel_j^2.
__global__ void mel_project_kernel(const float* __restrict__ power, const float* __restrict__ fb,
                                   float* __restrict__ mel, float* __restrict__ energy, int n_freq, int n_mels,
                                   int frames) {
  // [n_freq] padded to a multiple of 4 floats + 4 (zero filled): the unrolled dot-product loop below is
  // compiled to 12-byte shared loads that may start at the last element (compute-sanitizer memcheck,
  // profiles/r2_sanitize_summary.txt)
  extern __shared__ float pw[];
  __shared__ float red[32];
  const int m = blockIdx.x, b = blockIdx.y;
  const float* prow = power + ((size_t)b * frames + m) * n_freq;
  const int n_pad = ((n_freq + 3) & ~3) + 4;
  for (int k = threadIdx.x; k < n_pad; k += blockDim.x) pw[k] = k < n_freq ? prow[k] : 0.f;
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
                                    float* __restrict__ mel_norm, act_t* __restrict__ blocked,
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
      blocked[cl_off(b, j, m, n_mels, frames, cl_cb(n_mels))] = f_to_act(v);
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

extern "C" int ou_mel_power(const float* x, const float* window, const float* fb, const float* dft,
                            float* power, float* mel, float* energy, int batch, int t, int n_fft, int hop,
                            int n_mels, int pad_left, int frames, void* stream) {
  OU_REQUIRE(x && window && fb && dft && power && mel && energy, "ou_mel_power: null pointer");
  OU_REQUIRE(batch > 0 && t > 0 && n_fft > 0 && n_fft % ou::MEL_BK == 0 && hop > 0 && n_mels > 0 && frames > 0,
             "ou_mel_power: bad shape (n_fft must be a multiple of 16)");
  const int n_freq = n_fft / 2 + 1;
  cudaStream_t st = (cudaStream_t)stream;
  dim3 grid(ou::ceil_div(2 * n_freq, ou::MEL_BN), ou::ceil_div(batch * frames, ou::MEL_BM));
  ou::dft_power_kernel<<<grid, 256, 0, st>>>(x, window, dft, power, batch, t, n_fft, hop, pad_left, frames);
  int rc = ou::check_launch("ou_mel_power(dft)");
  if (rc) return rc;
  ou::mel_project_kernel<<<dim3(frames, batch), 128, (size_t)(((n_freq + 3) & ~3) + 4) * sizeof(float), st>>>(power, fb, mel, energy,
                                                                                           n_freq, n_mels, frames);
  return ou::check_launch("ou_mel_power(project)");
}

extern "C" int ou_mel_finalize(const float* mel, const float* energy, float* mel_norm,
                               void* mel_blocked, int batch, int n_mels, int frames, void* stream) {
  OU_REQUIRE(mel && energy && (mel_norm || mel_blocked), "ou_mel_finalize: null pointer");
  OU_REQUIRE(batch > 0 && n_mels > 0 && frames > 0, "ou_mel_finalize: bad shape");
  OU_REQUIRE(mel_blocked == nullptr || n_mels % 16 == 0, "ou_mel_finalize: n_mels % 16 != 0");
  ou::mel_finalize_kernel<<<batch, 512, 0, (cudaStream_t)stream>>>(
      mel, energy, mel_norm, (act_t*)mel_blocked, n_mels, frames);
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
