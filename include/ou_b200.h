/*
 * ou_b200.h -- C ABI of libou_b200.so: hand-written sm_100a kernels for the enhance() hot path
 * of line/open-universe (UNIVERSE / UNIVERSE++ diffusion speech enhancement).
 *
 * The reference has NO native / FFI layer: its device work is eager ATen
 * and cuDNN calls made from Python.  This header is therefore the boundary a maintainer would
 * bind (ctypes stub in INTEGRATION.md); every entry point names the reference code it replaces
 * (paths relative to /root/reference/open_universe/).
 *
 * Conventions
 *   - every function returns 0 on success, a negative OU_ERR_* code otherwise; it never throws,
 *     never synchronises the device and never allocates: the caller owns all buffers and passes
 *     the CUDA stream (a cudaStream_t cast to void*; NULL = legacy default stream);
 *   - all pointers are DEVICE pointers unless stated otherwise;
 *   - ou_last_error() returns a thread-local human-readable message for the last failure.
 *
 * Activation layout ("blocked"): 16-bit "act" elements [B][C/CB][T][CB], CB = largest of {64, 32, 16} dividing C --
 *   channel c of time step t of clip b lives at ((b * (C/CB) + c/CB) * T + t) * CB + c%CB
 *   (channels-last within blocks of CB channels; C must be a multiple of 16).
 *   The element type of activations and packed conv weights is a build-time policy reported by
 *   ou_act_dtype(): IEEE fp16 by default (11-bit significand = the TF32 the reference's cuDNN path
 *   uses), act when built with -DOU_ACT_BF16.  "act" below means that type.
 * Signals are fp32 [B][T]; GRU pre-activations fp32 blocked [B][N/16][T][16].
 */
#ifndef OU_B200_H
#define OU_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define OU_ABI_VERSION 4

enum {
  OU_OK = 0,
  OU_ERR_INVALID = -1,   /* bad argument (shape / alignment / null pointer)          */
  OU_ERR_CUDA = -2,      /* a CUDA runtime call failed (launch, attribute, ...)      */
  OU_ERR_UNSUPPORTED = -3 /* valid request this build has no kernel for              */
};

int ou_abi_version(void);
/* Copies the last error message of the calling thread into buf (NUL-terminated). */
int ou_last_error(char* buf, size_t n);
/* Number of kernels launched by this library in the calling process since load (bench.py's
 * gpu_launches claim). */
int64_t ou_launch_count(void);
/* Storage / tensor-core operand type of this build: 0 = IEEE fp16 (default), 1 = bf16. */
int ou_act_dtype(void);
/* Number of ou_conv1d calls that the tcgen05 kernel rejected (input length not a multiple of the
 * stride: only reachable through the module-level entry points on odd lengths, never inside
 * enhance()) and that ran on the mma.sync kernel instead.  The first one is also logged to stderr. */
int64_t ou_conv_fallback_count(void);

/* ------------------------------------------------------------------------------------------
 * Fused implicit-GEMM Conv1d.   Replaces, in ONE launch:
 *   PReLU_Conv.forward                 networks/universe/blocks.py:205-227
 *     (right-pad, PReLU, binomial low-pass :119-130 folded into the weights, Conv1d /
 *      ConvTranspose1d, separate bias)
 *   the element-wise glue of ConvBlock.forward   blocks.py:355-412
 *     (residual / conditioning adds with 1/sqrt2, FiLM :53-59, the next layer's PReLU)
 *   LinearProj / signal_cond_proj 1x1 convs       score.py:165-170,189-194
 *   the GRU input projection x @ W_ih^T + b_ih    score.py:83-89 (torch.nn.GRU)
 *   the conditioner's st_convs                    condition.py:33-65
 *
 *   acc[j][n] = sum_{q<taps} sum_{c'<s*cin} W[n][q][c'] * PReLU_in(X)[ci][(j+tap_off+q)*s + r],
 *               c' = r*cin + ci, X zero outside [0, t_in)
 *   co = n % cout, p = n / cout, t = j*up + p   (t < t_out, j < rows)
 *   y = ((acc + bias[n] + add1[co][t]) * scale1 + add2[co][t]) * scale2
 *   y = gamma[b][co] * y + beta[b][co]                    (if gamma != NULL)
 *   y = PReLU(PReLU(y, prelu_out), prelu_out2)            (each if enabled)
 *   out (blocked act (B, cout, t_out))  or  out_f32_blk (fp32 [B][n/16][rows][16], no add/film/prelu)
 * ------------------------------------------------------------------------------------------ */
typedef struct ou_conv_params {
  const void* x;          /* blocked act (B, cin, t_in)                                  */
  const void* w;          /* packed act [taps][kpad/8][npad][8]; element (q, c', n_) = W[n_][q][c'],
                             zero padded; kpad % 32 == 0, npad % 32 == 0 (mma.sync / naive path) */
  const void* w_tc;       /* same weights packed for the tcgen05 path, or NULL: act
                             [taps][s*cin/CB][npad][CB], CB = channel block of the input layout
                             (K index c' = r*cin + ci cut into blocks of CB)                    */
  const float* bias;      /* fp32 [n] or NULL                                                   */
  const void* add1;       /* blocked act (B, cout, t_out) or NULL                         */
  const void* add2;       /* blocked act (B, cout, t_out) or NULL                         */
  const float* gamma;     /* fp32, element (b, co) at gamma[b*film_bstride + co], or NULL       */
  const float* beta;      /* fp32, same indexing                                                */
  void* out;              /* blocked act (B, cout, t_out) or NULL                         */
  float* out_f32_blk;     /* fp32 [B][n/16][rows][16] or NULL (exactly one of out / out_f32_blk) */
  int32_t batch, cin, t_in;
  int32_t s, taps, tap_off;
  int32_t n, cout, up;
  int32_t kpad, npad;     /* padded K (= s*cin rounded up to 32) and N of the packed weights    */
  int32_t rows, t_out;
  int32_t film_bstride;
  int32_t has_prelu_in, has_prelu_out, has_prelu_out2;
  float prelu_in, prelu_out, prelu_out2;
  float scale1, scale2;
  int32_t max_ctas;       /* cap on the persistent grid (0 = one CTA per SM): the host leaves SMs to a
                             kernel of another stream it wants to run concurrently (GRU overlap)  */
} ou_conv_params;

int ou_conv1d(const ou_conv_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * Fused ConvBlock trunk (conv1 -> FiLM -> conv2 -> conv3 -> residual) for narrow levels: ONE launch
 * that reads the block input (and the conditioning tensor) once and writes the block output once;
 * the two intermediate activations stay in shared memory.  Replaces the middle of
 * ConvBlock.forward, networks/universe/blocks.py:385-399:
 *   c1 = PReLU(FiLM((conv1_k5(PReLU(x, prelu_in)) + b1 [+ sc]) * scale1), prelu_mid1)   (act)
 *   c2 = PReLU(conv2_k3(c1) + b2, prelu_mid2)                                            (act)
 *   v  = (conv3_k3(c2) + b3 + x) * scale3  -> PReLU(prelu_out) -> PReLU(prelu_out2)  (each if enabled)
 * All convs are 'same' (zero padded) C -> C channels; FiLM = gamma[b][c] * y + beta[b][c] when gamma
 * is not NULL.  Rounding points are those of three consecutive ou_conv1d calls.
 * Supported: channels in {32, 64}, taps (5, 3, 3); anything else returns OU_ERR_UNSUPPORTED (the
 * caller runs the three ou_conv1d launches instead).
 * ------------------------------------------------------------------------------------------ */
typedef struct ou_trunk_params {
  const void* x;            /* blocked act (B, C, t): raw block input                          */
  const void* w1;           /* ou_conv_params.w_tc packing of conv1 / conv2 / conv3:            */
  const void* w2;           /*   act [taps][1][C][C]  (tap, out channel, in channel)            */
  const void* w3;
  const float* b1;          /* fp32 [C] each                                                     */
  const float* b2;
  const float* b3;
  const void* sc;           /* blocked act (B, C, t) or NULL                                    */
  const float* gamma;       /* fp32, element (b, c) at gamma[b*film_bstride + c], or NULL        */
  const float* beta;
  void* out;                /* blocked act (B, C, t)                                            */
  int32_t batch, channels, t;
  int32_t taps1, taps2, taps3;
  int32_t film_bstride;
  int32_t has_prelu_out, has_prelu_out2;
  float prelu_in, prelu_mid1, prelu_mid2, prelu_out, prelu_out2;
  float scale1, scale3;
  int32_t max_ctas;         /* cap on the persistent grid (0 = one CTA per SM), as in ou_conv_params */
  /* Optional tail (ABI v3, channels = 64 only): the NEXT decoder block's transposed rate-change conv
   * (blocks.py:360-376: PReLU -> ConvTranspose1d(C -> C/2, k = stride = 2) -> low-pass -> bias -> pad to
   * length -> (h + res)/sqrt2, folded to 3 taps at the low rate) runs on the block output while it is
   * still in shared memory; `out` is then not written (may be NULL):
   *   up_out[b][co][2 j + ph] = (sum_q W_up[ph*C/2 + co][q][:] . PReLU(v, up_prelu_in)[:, j + q - 1]
   *                              + up_bias[ph*C/2 + co] + up_skip[b][co][2 j + ph]) * up_scale          */
  const void* up_w;         /* ou_conv_params.w_tc packing [3][1][C][C] of the folded up conv, or NULL   */
  const float* up_bias;     /* fp32 [C] or NULL                                                         */
  const void* up_skip;      /* blocked act (B, C/2, up_t_out) or NULL                                   */
  void* up_out;             /* blocked act (B, C/2, up_t_out)                                           */
  int32_t up_t_out;         /* <= 2 t                                                                   */
  float up_scale, up_prelu_in;
  /* Optional output tail (ABI v3, channels = 32 only): the network's output conv (score.py:288,294: C -> 1,
   * k = 3, 'same') on the block output + EDM mix + reverse-SDE update, exactly ou_output_sde:
   *   net[t] = out_bias + sum_{k<3} sum_c out_w[k][c] * v[c][t + k - 1];  out_net[b][t] = net (if not NULL)
   *   out_xout[b][t] = ca*out_x[b][t] + cb*net + cc*out_noise[b][t], (ca, cb, cc) = out_coef[b] (if not NULL)
   * `out` is then not written (may be NULL).                                                            */
  const float* out_w;       /* fp32 [3][C] tap-major, or NULL                                           */
  float out_bias;
  const float* out_coef;    /* fp32 (B, 3) or NULL                                                      */
  const float* out_x;       /* fp32 (B, t)                                                              */
  const float* out_noise;   /* fp32 (B, t) or NULL                                                      */
  float* out_xout;          /* fp32 (B, t)                                                              */
  float* out_net;           /* fp32 (B, t) or NULL                                                      */
  /* Optional down tail (ABI v4, channels = 32 only): the block's own anti-aliased stride-2 rate-change conv
   * (blocks.py:205-227,400-404: PReLU -> binomial low-pass -> Conv1d(k = 2, stride 2), folded to 3 row taps of
   * 2 samples = 6 sample taps) on the block output while it is still in shared memory:
   *   dn_out[b][n][j] = dn_bias[n] + sum_{m<6} sum_c dn_w[m][n][c] * PReLU_dn(v)[c][2 j - 2 + m]
   * `out` (the block output, the decoder's skip connection) is still written.                            */
  const void* dn_w;         /* act [6][2C][C]: sample tap m, output channel n, input channel c, or NULL  */
  const float* dn_bias;     /* fp32 [2C]                                                                 */
  void* dn_out;             /* blocked act (B, 2C, dn_t_out)                                             */
  int32_t dn_t_out;         /* ceil(t / 2)                                                               */
  float dn_prelu_in;
  /* Row taps the up / down tail actually carries: 3 (anti-aliased, the low-pass folded in) or 1 (plain k = s
   * conv of UNIVERSE (original): up_w / dn_w keep the 3-tap frame with the conv in the middle tap, the outer
   * taps are neither read nor multiplied); 0 means 3.                                                     */
  int32_t up_taps, dn_taps;
} ou_trunk_params;

int ou_conv_trunk(const ou_trunk_params* p, void* stream);

/* Reference-quality fp32 CUDA-core version of the same contract (one thread per output): used by
 * the GPU tests to cross-check the tensor-core kernel on device.  Same arguments. */
int ou_conv1d_naive(const ou_conv_params* p, void* stream);

/* ------------------------------------------------------------------------------------------
 * First layer: k-tap 'same' Conv1d of the 1-channel signal, with the EDM input scale folded in.
 * Replaces ScoreNetwork.input_conv (score.py:239-241,285) + `w_in * x` of _edm_score_wrapper
 * (universe.py:197-203), and ConditionerNetwork.input_conv (condition.py:290-295,361).
 *   out[co][t] = bias[co] + sum_k w[co][k] * in_scale[b] * x[b][t + k - k/2]
 * x fp32 [B][t]; w fp32 [cout][k]; in_scale fp32 [B] or NULL; out blocked act (B, cout, t).
 * ------------------------------------------------------------------------------------------ */
int ou_input_conv(const float* x, const float* w, const float* bias, const float* in_scale,
                  void* out, int batch, int t, int cout, int k, void* stream);

/* ------------------------------------------------------------------------------------------
 * Last layer fused with the sampler update.  Replaces ScoreNetwork.output_conv + final pad
 * (score.py:288,294), the EDM mix `w_skip*x + w_out*net`, `score = (est-x)/sigma^2`
 * (universe.py:204-206) and the reverse-SDE update (universe.py:337-339, 342-343):
 *   net[b][t]  = bias + sum_{ci,k} w[ci][k] * src[ci][t + k - k/2]      (t < t_src, else 0)
 *   xout[b][t] = ca[b]*x[b][t] + cb[b]*net[b][t] + cc[b]*noise[b][t]    (t < t_sig)
 * src blocked act (B, cin, t_src) (already activated by the producer); w fp32 [cin][k];
 * coef fp32 [B][3] = (ca, cb, cc) or NULL; noise fp32 [B][t_sig] or NULL; net_out fp32 [B][t_sig]
 * or NULL; x / xout fp32 [B][t_sig] (may alias) or NULL when coef is NULL.
 * ------------------------------------------------------------------------------------------ */
int ou_output_sde(const void* src, const float* w, float bias, const float* coef, const float* x,
                  const float* noise, float* xout, float* net_out, int batch, int cin, int k,
                  int t_src, int t_sig, void* stream);

/* ------------------------------------------------------------------------------------------
 * Bidirectional GRU recurrence (one layer).  Replaces torch.nn.GRU at the score bottleneck
 * (score.py:82-89,116-117; every step) and in the conditioner (condition.py:173-179,213).
 *   gx    fp32 blocked [B][6H/16][T][16] (what ou_conv1d writes through out_f32_blk): x W_ih^T + b_ih,
 *         columns = forward (r,z,n) then backward (r,z,n)
 *   w_hh  fp32 [2][3H][H],  b_hh fp32 [2][3H]
 *   r = s(gx_r + W_hr h + b_hr); z = s(gx_z + W_hz h + b_hz); n = tanh(gx_n + r*(W_hn h + b_hn));
 *   h' = (1-z)*n + z*h; h0 = 0; backward direction runs t = T-1..0
 *   out blocked act (B, 2H, T) = ((h_fwd | h_bwd) + add) * scale   (add blocked or NULL)
 * ------------------------------------------------------------------------------------------ */
/* ou_gru_bidir_ex: cluster_ctas = 0 / 8: clusters of 8 CTAs per (direction, 8 clips) -- lowest latency;
 * 4: clusters of 4 (hidden <= 256; otherwise as 0): a step takes ~30 % longer but the recurrence holds half
 * the SMs -- for hosts that overlap it with other kernels (engine/runtime.py: PipelinedScoreRunner).
 * ou_gru_ctas: CTAs such a launch occupies (what to leave free when capping other grids with max_ctas). */
int ou_gru_bidir_ex(const float* gx, const float* w_hh, const float* b_hh, const void* add, float scale,
                    void* out, int batch, int t, int hidden, int cluster_ctas, void* stream);
int ou_gru_ctas(int hidden, int batch, int cluster_ctas);
int ou_gru_bidir(const float* gx, const float* w_hh, const float* b_hh, const void* add, float scale,
                 void* out, int batch, int t, int hidden, void* stream);

/* ------------------------------------------------------------------------------------------
 * Mel front-end.  Replaces MelAdapter.compute_mel_spec (condition.py:92-108): zero pad, frame,
 * periodic Hann, |rFFT|^2 (torchaudio MelSpectrogram power=2, center=False), mel filterbank and
 * the global energy normalisation.
 *   ou_mel_power:    x fp32 [B][t] -> mel fp32 [B][n_mels][frames] (un-normalised),
 *                    energy fp32 [B][frames] = sum_mel mel^2.  frame m = samples
 *                    [hop*m - pad_left, hop*m - pad_left + n_fft), zero outside [0, t).
 *                    window fp32 [n_fft]; fb fp32 [n_fft/2+1][n_mels];
 *                    dft fp32 [n_fft][2*(n_fft/2+1)]: columns (2k, 2k+1) = (cos, sin)(2*pi*i*k/n_fft)
 *                    (the power spectrum is one fp32 GEMM against it; n_fft % 16 == 0);
 *                    power fp32 [B*frames][n_fft/2+1] scratch, overwritten.
 *   ou_mel_finalize: per clip scale = 1 / max(sqrt(mean_frames energy), 1e-5); writes the
 *                    normalised mel as fp32 [B][n_mels][frames] (in place allowed, or NULL) and
 *                    blocked act (B, n_mels, frames) (or NULL).
 * ------------------------------------------------------------------------------------------ */
int ou_mel_power(const float* x, const float* window, const float* fb, const float* dft, float* power,
                 float* mel, float* energy, int batch, int t, int n_fft, int hop, int n_mels,
                 int pad_left, int frames, void* stream);
int ou_mel_finalize(const float* mel, const float* energy, float* mel_norm, void* mel_blocked,
                    int batch, int n_mels, int frames, void* stream);

/* ------------------------------------------------------------------------------------------
 * sigma embedding + small dense layers (step-invariant: computed for all steps before the loop).
 * Replaces SimpleTimeEmbedding.forward / SigmaBlock.forward (sigma_block.py:73-78, 50-57) and the
 * FiLM projections Linear(g) (score.py:58-66,108,159-164,207).
 *   ou_sigma_embed_simple: f = 0.5*sigmoid(w*ls + b); out[r][k] = sin(2*pi*f*k), out[r][half+k] =
 *                          cos(2*pi*f*k), k < half.           log10_sigma fp32 [rows], out [rows][2*half]
 *   ou_sigma_embed_rff:    out[r][k] = sin(2*pi*freq[k]*ls), out[r][n_rff+k] = cos(...)
 *   ou_linear_f32:         out[r][c] = act(bias[c] + sum_k w[c][k] * in[r][k]); act = PReLU(slope)
 *                          if has_prelu.   in [rows][k], w [n][k], out [rows][n] (ld_out >= n)
 * ------------------------------------------------------------------------------------------ */
int ou_sigma_embed_simple(const float* log10_sigma, float weight, float bias, float* out, int rows,
                          int half, void* stream);
int ou_sigma_embed_rff(const float* log10_sigma, const float* freq, float* out, int rows, int n_rff,
                       void* stream);
int ou_linear_f32(const float* in, const float* w, const float* bias, float* out, int rows, int k,
                  int n, int ld_out, int has_prelu, float slope, void* stream);

/* ------------------------------------------------------------------------------------------
 * Signal-level pre/post processing of Universe.enhance (universe.py:251-274, 346-357).
 *   ou_pad_normalize: centred zero pad (universe.py:219-223) + utils.normalize_batch with norm=2
 *                     (utils/norm.py:47-87): mean / unbiased std over the PADDED clip.
 *                     mix fp32 [B][t] -> out fp32 [B][t_pad], out[b][pad_left + i] = (mix - mean)*gain,
 *                     padding = (0 - mean)*gain; gain = level / max(std, 1e-5).
 *                     stats fp32 [B][2] receives (mean, 1/gain).  One CTA per clip.
 *   ou_unpad_limit:   unpad, right-pad to t, optional keep_rms rescale to mix_rms[b], peak limiter
 *                     x/max|x| where max|x| > 1 (universe.py:349-357).  x fp32 [B][t_pad] ->
 *                     out fp32 [B][t].  mix_rms fp32 [B] or NULL.
 * ------------------------------------------------------------------------------------------ */
int ou_pad_normalize(const float* mix, float* out, float* stats, int batch, int t, int t_pad,
                     int pad_left, float level, void* stream);
int ou_unpad_limit(const float* x, const float* mix_rms, float* out, int batch, int t_pad,
                   int pad_left, int t_valid, int t, void* stream);

/* Layout converters between the reference's (B, C, T) fp32 tensors and the blocked act layout
 * (module-level APIs: ScoreNetwork.forward's `cond` list, ConditionerNetwork's outputs). */
int ou_pack_blocked(const float* src, void* dst, int batch, int channels, int t, void* stream);
int ou_unpack_blocked(const void* src, float* dst, int batch, int channels, int t, void* stream);

/* FiLM on (B, C, T) fp32 tensors: out = y[:, :C] * x + y[:, C:]  (blocks.py:53-59, standalone). */
int ou_film_f32(const float* x, const float* y, float* out, int batch, int channels, int t,
                void* stream);

/* ------------------------------------------------------------------------------------------
 * AliasFreeSnake (networks/bigvgan/snake.py:127-157, alias_free_act.py:8-30) and, fused behind it,
 * the k-tap 'same' convolution to ONE channel of UniverseGAN.signal_decoupling_layer
 * (universe_gan.py:117-126; PReLU_Conv.forward with act_type snake / snakebeta, blocks.py:205-227):
 * the `aux_to_wav` step of enhance(use_aux_signal=True / warm_start=n) (universe.py:317-331).
 *   x            activations, act blocked [B][C/CB][T][CB] (x_blocked != 0) or fp32 [B][C][T]
 *   alpha, beta  fp32 [C] Snake parameters (beta NULL = Snake, else SnakeBeta); exp() applied when
 *                logscale != 0 (snake.py:52-62, 116-124)
 *   up_kernel    fp32 [2][up_len]   torchaudio Resample(1 -> 2) `kernel` buffer (up_len odd)
 *   down_kernel  fp32 [down_len]    torchaudio Resample(2 -> 1) `kernel` buffer (down_len even)
 *   w, bias, k   fp32 [C][k] conv weights (weight norm already folded), scalar bias, odd k; w NULL =
 *                activation only
 *   out          fp32 [B][T] (w != NULL) or fp32 [B][C][T] (w == NULL)
 * ------------------------------------------------------------------------------------------ */
int ou_alias_free_snake(const void* x, int x_blocked, const float* alpha, const float* beta,
                        int logscale, const float* up_kernel, int up_len, const float* down_kernel,
                        int down_len, const float* w, float bias, int k, float* out, int batch,
                        int channels, int t, void* stream);

/* Debug hook: when set to a device buffer of 4*64*4 int64, CTA 0 of the tcgen05 conv kernel stamps
 * clock64() per warp role / tile / event into it (tools/trace_conv.py).  NULL disables. */
int ou_debug_set_trace(void* device_buffer);

/* ------------------------------------------------------------------------------------------
 * Around the path.
 *
 * ou_resample_poly: polyphase windowed-sinc sample-rate conversion -- replaces the
 *   torchaudio.functional.resample calls of the reference CLI (bin/enhance.py:77-80,188-190).
 *   out[b][n*up + p] = sum_{j<klen} kern[p][j] * x[b][n*down + j - width], x zero outside [0, t_in);
 *   kern fp32 [up][klen] is built on the host with torchaudio's published formula
 *   (open_universe_b200/utils/resample.py); rates already divided by their gcd.
 * ou_lsd: log-spectral distance -- replaces metrics/lsd.py:26-147 (log_spectral_distance with
 *   center=True reflect padding, power = 2, normalized = "window", onesided):
 *   out[b] = ( mean_{bins, frames} | L(input) - L(s_b * target) |^p )^(1/p),  L = 10 log10(P + eps) if db
 *   else ln(P + eps); s_b = <input, target> / (<input, input> + eps) if scale_invariant else 1.
 *   partial: fp32 [batch][frames] scratch; scale: fp32 [batch] scratch (scale_invariant only).
 * ------------------------------------------------------------------------------------------ */
int ou_resample_poly(const float* x, const float* kern, float* out, int batch, int t_in, int t_out,
                     int down, int up, int width, int klen, void* stream);
int ou_lsd(const float* input, const float* target, const float* window, float* partial, float* scale,
           float* out, int batch, int t, int n_fft, int hop, int frames, float p, int db, float eps,
           float window_sumsq, int scale_invariant, void* stream);

/* ------------------------------------------------------------------------------------------
 * Plan level: ONE call per network evaluation.
 *
 * A plan is a recorded list of the launches of ScoreNetwork.forward (score.py:277-297) -- or of
 * ConditionerNetwork.forward (condition.py:346-377): SURVEY 8b's ou_condition_forward -- for a fixed
 * (batch, length) with every weight / activation pointer resolved; the host lowers the network once
 * (engine/program.py), records the ops in launch order with ou_plan_add_*, and replays them with
 * ou_plan_run.  What changes from one evaluation to the next comes in ou_step_args: the signal, the
 * FiLM row(s) of this noise level (gamma of an op = film + film_off, beta behind it), the EDM input
 * scale, the affine update coefficients of ou_output_sde, the noise and the outputs
 * (universe.py:197-209, 334-343).  ou_plan_run(first, count) replays a contiguous part of the list
 * (count < 0: to the end) -- the pipelined sampler runs encoder / recurrence / decoder of two
 * half-batch plans on two streams.  The plan owns no device memory; the caller keeps every buffer
 * alive.  Same return-code contract as the per-kernel entry points.
 * ------------------------------------------------------------------------------------------ */
typedef struct ou_plan ou_plan;
typedef struct ou_step_args {
  const float* x;         /* fp32 (B, 1, T) signal: input conv and EDM / SDE update               */
  const float* film;      /* fp32 FiLM row(s) of this evaluation, or NULL                          */
  int32_t film_bstride;   /* floats between the rows of consecutive clips (0: one row for all)     */
  const float* in_scale;  /* fp32 (B,) EDM input scale, or NULL                                    */
  const float* coef;      /* fp32 (B, 3) update coefficients (ca, cb, cc), or NULL                 */
  const float* noise;     /* fp32 (B, 1, T), or NULL                                               */
  float* xout;            /* fp32 (B, 1, T) updated signal (may alias x), or NULL                  */
  float* net_out;         /* fp32 (B, 1, T) raw network output, or NULL                            */
  const float* x_wav;     /* conditioner plans: fp32 (B, 1, T) waveform of the mel front-end, NULL = x */
} ou_step_args;

int ou_plan_create(ou_plan** plan);
int ou_plan_destroy(ou_plan* plan);
int ou_plan_size(const ou_plan* plan);
int ou_plan_add_conv(ou_plan* plan, const ou_conv_params* p, int32_t film_off /* -1: no FiLM */);
int ou_plan_add_trunk(ou_plan* plan, const ou_trunk_params* p, int32_t film_off);
int ou_plan_add_input_conv(ou_plan* plan, const float* w, const float* bias, void* out, int batch, int t,
                           int cout, int k, int use_in_scale);
int ou_plan_add_output_sde(ou_plan* plan, const void* src, const float* w, float bias, int batch, int cin,
                           int k, int t_src, int t_sig);
int ou_plan_add_gru(ou_plan* plan, const float* gx, const float* w_hh, const float* b_hh, const void* add,
                    float scale, void* out, int batch, int t, int hidden, int cluster_ctas /* see ou_gru_bidir_ex */);
/* ConditionerNetwork.forward (condition.py:346-377) is recorded the same way; its STFT-mel front-end
 * (condition.py:92-108) is one op = ou_mel_power + ou_mel_finalize on args->x_wav (or args->x). */
int ou_plan_add_mel(ou_plan* plan, const float* window, const float* fb, const float* dft, float* power, float* mel,
                    float* energy, void* mel_blocked, int batch, int t, int n_fft, int hop, int n_mels, int pad_left,
                    int frames);
int ou_plan_run(const ou_plan* plan, const ou_step_args* args, int first, int count, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* OU_B200_H */
