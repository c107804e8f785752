// AliasFreeSnake activation, optionally fused with the k-tap convolution to ONE channel that follows
// it in UniverseGAN.signal_decoupling_layer (see ou_b200.h: ou_alias_free_snake).
//
//   up   : y[2j + p] = sum_i ku[p][i] * x[j + i - wu]            (torchaudio Resample 1 -> 2, x = 0 outside)
//   act  : s = y + sin^2(alpha y) / (beta + 1e-9)                (Snake / SnakeBeta, log-scale params)
//   down : z[j]      = sum_i kd[i] * s[2j + i - wd]              (Resample 2 -> 1, s = 0 outside [0, 2T))
//   conv : out[t]    = bias + sum_c sum_q w[c][q] * z[c][t + q - k/2]   (z = 0 outside [0, T))
//
// Runs once per enhance() call on the warm-start / aux-signal path only: a plain fp32 CUDA-core
// kernel, one CTA per 64 output samples of one clip with the three intermediate tiles in shared
// memory (nothing but x is read from, nothing but the result written to, HBM).
#include "common.cuh"

namespace ou {

constexpr int SNK_NT = 64;        // output samples per CTA
constexpr int SNK_THREADS = 256;

struct SnakeArgs {
  const void* x;
  int x_blocked;
  const float* alpha;
  const float* beta;
  int logscale;
  const float* ku;
  int ku_len;
  const float* kd;
  int kd_len;
  const float* w;
  float bias;
  int k;
  float* out;
  int channels, t;
  int nx, ns, nz;   // tile widths (row strides are these + 1 when even, to spread banks)
};

__device__ __forceinline__ int floor_div2(int a) { return a >> 1; }   // arithmetic shift: floor

__global__ void __launch_bounds__(SNK_THREADS) alias_free_snake_kernel(const SnakeArgs a) {
  extern __shared__ float sm[];
  const int C = a.channels, T = a.t;
  const int b = blockIdx.y, t0 = blockIdx.x * SNK_NT;
  const int hc = a.w ? a.k / 2 : 0;
  const int wu = (a.ku_len - 1) / 2, wd = (a.kd_len - 2) / 2;
  // z positions [z0, z0 + nz), s positions [s0, s0 + ns) with s0 even, x positions [x0, x0 + nx)
  const int z0 = t0 - hc;
  const int s0 = 2 * floor_div2(2 * z0 - wd);
  const int x0 = floor_div2(s0) - wu;
  const int ldx = a.nx | 1, lds_ = a.ns | 1, ldz = a.nz | 1;
  float* xs = sm;
  float* ss = xs + C * ldx;
  float* zs = ss + C * lds_;
  float* par = zs + C * ldz;          // [alpha | 1 / (beta + 1e-9)] per channel, taps of both kernels
  float* ku = par + 2 * C;
  float* kd = ku + 2 * a.ku_len;

  const int tid = threadIdx.x;
  for (int c = tid; c < C; c += SNK_THREADS) {
    const float al = a.logscale ? expf(a.alpha[c]) : a.alpha[c];
    const float be = a.beta ? (a.logscale ? expf(a.beta[c]) : a.beta[c]) : al;
    par[c] = al;
    par[C + c] = 1.0f / (be + 0.000000001f);
  }
  for (int i = tid; i < 2 * a.ku_len; i += SNK_THREADS) ku[i] = a.ku[i];
  for (int i = tid; i < a.kd_len; i += SNK_THREADS) kd[i] = a.kd[i];

  // ---- x tile
  if (a.x_blocked) {
    const act_t* xb = (const act_t*)a.x;
    const int cb = cl_cb(C);
    for (int i = tid; i < C * a.nx; i += SNK_THREADS) {
      const int c = i % C, p = i / C;     // channel fastest: contiguous in the blocked layout
      const int t = x0 + p;
      xs[c * ldx + p] = (t >= 0 && t < T) ? act_to_f(xb[cl_off(b, c, t, C, T, cb)]) : 0.f;
    }
  } else {
    const float* xf = (const float*)a.x;
    for (int i = tid; i < C * a.nx; i += SNK_THREADS) {
      const int p = i % a.nx, c = i / a.nx;
      const int t = x0 + p;
      xs[c * ldx + p] = (t >= 0 && t < T) ? xf[((size_t)b * C + c) * T + t] : 0.f;
    }
  }
  __syncthreads();

  // ---- up-sample + snake
  for (int i = tid; i < C * a.ns; i += SNK_THREADS) {
    const int q = i % a.ns, c = i / a.ns;
    const int m = s0 + q;
    float s = 0.f;
    if (m >= 0 && m < 2 * T) {
      const int j = q >> 1, ph = q & 1;   // s0 is even: phase = q & 1, x offset = j (x0 = s0/2 - wu)
      const float* xr = xs + c * ldx + j;
      const float* kr = ku + ph * a.ku_len;
      float y = 0.f;
      for (int u = 0; u < a.ku_len; u++) y = fmaf(kr[u], xr[u], y);
      const float sn = sinf(y * par[c]);
      s = y + par[C + c] * (sn * sn);
    }
    ss[c * lds_ + q] = s;
  }
  __syncthreads();

  // ---- down-sample
  const int soff = 2 * z0 - wd - s0;      // 0 or 1
  for (int i = tid; i < C * a.nz; i += SNK_THREADS) {
    const int q = i % a.nz, c = i / a.nz;
    const int j = z0 + q;
    float z = 0.f;
    if (j >= 0 && j < T) {
      const float* sr = ss + c * lds_ + 2 * q + soff;
      for (int u = 0; u < a.kd_len; u++) z = fmaf(kd[u], sr[u], z);
    }
    zs[c * ldz + q] = z;
  }
  __syncthreads();

  if (a.w) {
    // ---- conv to one channel
    for (int q = tid; q < SNK_NT; q += SNK_THREADS) {
      const int t = t0 + q;
      if (t >= T) continue;
      float acc = a.bias;
      for (int c = 0; c < C; c++)
        for (int u = 0; u < a.k; u++) acc = fmaf(a.w[c * a.k + u], zs[c * ldz + q + u], acc);
      a.out[(size_t)b * T + t] = acc;
    }
  } else {
    for (int i = tid; i < C * SNK_NT; i += SNK_THREADS) {
      const int q = i % SNK_NT, c = i / SNK_NT;
      const int t = t0 + q;
      if (t < T) a.out[((size_t)b * C + c) * T + t] = zs[c * ldz + q];
    }
  }
}

}  // namespace ou

extern "C" int ou_alias_free_snake(const void* x, int x_blocked, const float* alpha, const float* beta,
                                   int logscale, const float* up_kernel, int up_len,
                                   const float* down_kernel, int down_len, const float* w, float bias,
                                   int k, float* out, int batch, int channels, int t, void* stream) {
  OU_REQUIRE(x && alpha && up_kernel && down_kernel && out, "ou_alias_free_snake: null pointer");
  OU_REQUIRE(batch > 0 && channels > 0 && t > 0, "ou_alias_free_snake: empty problem");
  OU_REQUIRE(!x_blocked || channels % 16 == 0, "ou_alias_free_snake: blocked input needs C %% 16 == 0");
  OU_REQUIRE(up_len >= 1 && (up_len & 1) && down_len >= 2 && !(down_len & 1) && up_len <= 63 && down_len <= 128,
             "ou_alias_free_snake: resampling kernels must be (2, odd) and (1, even) taps (torchaudio Resample 1<->2)");
  OU_REQUIRE(w == nullptr || (k >= 1 && (k & 1) && k <= 15), "ou_alias_free_snake: conv kernel size must be odd");
  ou::SnakeArgs a;
  a.x = x, a.x_blocked = x_blocked, a.alpha = alpha, a.beta = beta, a.logscale = logscale;
  a.ku = up_kernel, a.ku_len = up_len, a.kd = down_kernel, a.kd_len = down_len;
  a.w = w, a.bias = bias, a.k = w ? k : 1, a.out = out, a.channels = channels, a.t = t;
  const int hc = w ? k / 2 : 0;
  a.nz = ou::SNK_NT + 2 * hc;
  a.ns = 2 * (a.nz - 1) + down_len + 2;       // + 2: s0 is rounded down to an even position
  a.nx = a.ns / 2 + 1 + up_len;
  const size_t smem = sizeof(float) * ((size_t)channels * ((a.nx | 1) + (a.ns | 1) + (a.nz | 1)) + 2 * channels +
                                       2 * up_len + down_len);
  if (smem > 200 * 1024) {
    ou::set_error("ou_alias_free_snake: %d channels need %zu bytes of shared memory", channels, smem);
    return OU_ERR_UNSUPPORTED;
  }
  static ou::SmemConfig cfg;
  if (smem > 48 * 1024)
    if (int rc = ou::ensure_smem(ou::alias_free_snake_kernel, smem, cfg, "ou_alias_free_snake")) return rc;
  dim3 grid(ou::ceil_div(t, ou::SNK_NT), batch);
  ou::alias_free_snake_kernel<<<grid, ou::SNK_THREADS, smem, (cudaStream_t)stream>>>(a);
  return ou::check_launch("ou_alias_free_snake");
}
