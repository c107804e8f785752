// Post-path kernels around enhance(), sm_100a:
//   ou_resample_poly  polyphase windowed-sinc sample-rate conversion of the CLI
//                     (reference bin/enhance.py:77-80,188-190 -> torchaudio.functional.resample)
//   ou_lsd            log-spectral distance metric (reference metrics/lsd.py:26-147)
// Both are HBM-trivial element-wise / small-DFT kernels on CUDA cores.
#include "common.cuh"

namespace ou {

// out[b][n * up + p] = sum_j kern[p][j] * x[b][n * down + j - width]   (x zero outside [0, t_in))
// STAGE: the whole filter bank fits in shared memory (small rate ratios); otherwise it is read through L1
// (44.1 k -> 16 k: 160 phases x 475 taps = 304 KB)
template <bool STAGE>
__global__ void resample_poly_kernel(const float* __restrict__ x, const float* __restrict__ kern,
                                     float* __restrict__ out, int t_in, int t_out, int down, int up, int width,
                                     int klen) {
  extern __shared__ float ks_smem[];   // [up][klen]
  const float* ks = kern;
  if (STAGE) {
    for (int i = threadIdx.x; i < up * klen; i += blockDim.x) ks_smem[i] = kern[i];
    __syncthreads();
    ks = ks_smem;
  }
  const int b = blockIdx.y;
  const float* xb = x + (size_t)b * t_in;
  for (int o = blockIdx.x * blockDim.x + threadIdx.x; o < t_out; o += gridDim.x * blockDim.x) {
    const int n = o / up, p = o - n * up;
    const int i0 = n * down - width;
    const float* kp = ks + p * klen;
    const int j0 = i0 < 0 ? -i0 : 0;
    const int j1 = i0 + klen > t_in ? t_in - i0 : klen;
    float acc = 0.f;
    for (int j = j0; j < j1; j++) acc = fmaf(kp[j], __ldg(xb + i0 + j), acc);
    out[(size_t)b * t_out + o] = acc;
  }
}

// One CTA per (frame, clip): reflect-padded, windowed frames of both signals -> power spectra by a direct
// DFT against a shared twiddle table -> sum over bins of |log P_in - log P_tgt|^p.
__global__ void lsd_frames_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                  const float* __restrict__ window, const float* __restrict__ y_scale,
                                  float* __restrict__ partial, int t, int n_fft, int hop, int frames, float p,
                                  int db, float eps, float inv_wnorm) {
  extern __shared__ float sm[];
  float* fx = sm;                    // [n_fft] windowed frame of the input
  float* fy = fx + n_fft;            // [n_fft] windowed frame of the (scaled) target
  float2* tw = reinterpret_cast<float2*>(fy + n_fft);   // [n_fft] (cos, sin)(2 pi j / n_fft)
  __shared__ float red[32];
  const int m = blockIdx.x, b = blockIdx.y;
  const float sc = y_scale ? y_scale[b] : 1.f;
  const float* xb = x + (size_t)b * t;
  const float* yb = y + (size_t)b * t;
  for (int i = threadIdx.x; i < n_fft; i += blockDim.x) {
    int idx = m * hop + i - n_fft / 2;               // center = True, pad_mode = "reflect"
    if (idx < 0) idx = -idx;
    if (idx >= t) idx = 2 * (t - 1) - idx;
    idx = idx < 0 ? 0 : idx;
    const float w = window[i];
    fx[i] = w * xb[idx];
    fy[i] = w * sc * yb[idx];
    float s, c;
    sincospif(2.f * (float)i / (float)n_fft, &s, &c);
    tw[i] = make_float2(c, s);
  }
  __syncthreads();
  const int bins = n_fft / 2 + 1;
  float acc = 0.f;
  for (int k = threadIdx.x; k < bins; k += blockDim.x) {
    float xr = 0.f, xi = 0.f, yr = 0.f, yi = 0.f;
    int idx = 0;
    for (int i = 0; i < n_fft; i++) {
      const float2 w = tw[idx];
      const float a = fx[i], c = fy[i];
      xr = fmaf(a, w.x, xr), xi = fmaf(a, w.y, xi);
      yr = fmaf(c, w.x, yr), yi = fmaf(c, w.y, yi);
      idx += k;
      if (idx >= n_fft) idx -= n_fft;
    }
    const float px = (xr * xr + xi * xi) * inv_wnorm + eps;   // normalized = "window": / sum(w^2)
    const float py = (yr * yr + yi * yi) * inv_wnorm + eps;
    float d = db ? 10.f * (log10f(px) - log10f(py)) : (logf(px) - logf(py));
    d = fabsf(d);
    acc += p == 2.f ? d * d : (p == 1.f ? d : powf(d, p));
  }
  acc = warp_sum(acc);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = acc;
  __syncthreads();
  if (warp == 0) {
    float v = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    v = warp_sum(v);
    if (lane == 0) partial[(size_t)b * frames + m] = v;
  }
}

// One CTA per clip: fixed-order sum of the per-frame partials -> (sum / (bins * frames))^(1/p); with
// `dots`: the scale-invariant factor <x, y> / (<x, x> + eps) of a clip instead (fp64 accumulation).
__global__ void lsd_finish_kernel(const float* __restrict__ partial, float* __restrict__ out, int frames,
                                  int bins, float p) {
  __shared__ double red[32];
  const int b = blockIdx.x;
  double s = 0.0;
  for (int i = threadIdx.x; i < frames; i += blockDim.x) s += (double)partial[(size_t)b * frames + i];
  s = warp_sum(s);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) red[warp] = s;
  __syncthreads();
  if (warp == 0) {
    double v = lane < (blockDim.x >> 5) ? red[lane] : 0.0;
    v = warp_sum(v);
    if (lane == 0) out[b] = (float)pow(v / ((double)bins * frames), 1.0 / (double)p);
  }
}

__global__ void si_scale_kernel(const float* __restrict__ x, const float* __restrict__ y,
                                float* __restrict__ scale, int t, float eps) {
  __shared__ double rxy[32], rxx[32];
  const int b = blockIdx.x;
  double sxy = 0.0, sxx = 0.0;
  for (int i = threadIdx.x; i < t; i += blockDim.x) {
    const double a = x[(size_t)b * t + i], c = y[(size_t)b * t + i];
    sxy += a * c, sxx += a * a;
  }
  sxy = warp_sum(sxy), sxx = warp_sum(sxx);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (lane == 0) rxy[warp] = sxy, rxx[warp] = sxx;
  __syncthreads();
  if (warp == 0) {
    double a = lane < (blockDim.x >> 5) ? rxy[lane] : 0.0, c = lane < (blockDim.x >> 5) ? rxx[lane] : 0.0;
    a = warp_sum(a), c = warp_sum(c);
    if (lane == 0) scale[b] = (float)(a / (c + (double)eps));
  }
}

}  // namespace ou

extern "C" int ou_resample_poly(const float* x, const float* kern, float* out, int batch, int t_in, int t_out,
                                int down, int up, int width, int klen, void* stream) {
  OU_REQUIRE(x && kern && out, "ou_resample_poly: null pointer");
  OU_REQUIRE(batch > 0 && t_in > 0 && t_out > 0 && down > 0 && up > 0 && klen > 0 && width >= 0,
             "ou_resample_poly: bad shape");
  const size_t smem = (size_t)up * klen * sizeof(float);
  int gx = ou::ceil_div(t_out, 256);
  if (gx > 4 * ou::num_sms()) gx = 4 * ou::num_sms();
  if (smem <= 40 * 1024)
    ou::resample_poly_kernel<true><<<dim3(gx, batch), 256, smem, (cudaStream_t)stream>>>(x, kern, out, t_in, t_out,
                                                                                       down, up, width, klen);
  else
    ou::resample_poly_kernel<false><<<dim3(gx, batch), 256, 0, (cudaStream_t)stream>>>(x, kern, out, t_in, t_out,
                                                                                     down, up, width, klen);
  return ou::check_launch("ou_resample_poly");
}

extern "C" int ou_lsd(const float* input, const float* target, const float* window, float* partial, float* scale,
                      float* out, int batch, int t, int n_fft, int hop, int frames, float p, int db, float eps,
                      float window_sumsq, int scale_invariant, void* stream) {
  OU_REQUIRE(input && target && window && partial && out, "ou_lsd: null pointer");
  OU_REQUIRE(batch > 0 && t > n_fft / 2 && n_fft >= 2 && n_fft % 2 == 0 && hop > 0 && frames > 0 && p > 0.f &&
                 window_sumsq > 0.f,
             "ou_lsd: bad arguments (reflect padding needs t > n_fft / 2)");
  OU_REQUIRE(!scale_invariant || scale, "ou_lsd: scale_invariant needs the per-clip scale buffer");
  cudaStream_t st = (cudaStream_t)stream;
  if (scale_invariant) {
    ou::si_scale_kernel<<<batch, 512, 0, st>>>(input, target, scale, t, eps);
    int rc = ou::check_launch("ou_lsd(scale)");
    if (rc) return rc;
  }
  const size_t smem = (size_t)n_fft * (2 * sizeof(float) + sizeof(float2));
  static ou::SmemConfig cfg;
  int rc = ou::ensure_smem(ou::lsd_frames_kernel, smem, cfg, "ou_lsd");
  if (rc) return rc;
  ou::lsd_frames_kernel<<<dim3(frames, batch), 256, smem, st>>>(input, target, window,
                                                               scale_invariant ? scale : nullptr, partial, t, n_fft,
                                                               hop, frames, p, db, eps, 1.f / window_sumsq);
  if ((rc = ou::check_launch("ou_lsd(frames)"))) return rc;
  ou::lsd_finish_kernel<<<batch, 256, 0, st>>>(partial, out, frames, n_fft / 2 + 1, p);
  return ou::check_launch("ou_lsd(finish)");
}
