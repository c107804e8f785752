// Fused ConvBlock trunk for the narrow (C = 32 / 64) levels of the U-Nets, sm_100a:
//
//     v = (x + conv3(PReLU(conv2(PReLU(FiLM((conv1(PReLU(x)) [+ sc]) * s1)))))) * s3  [-> PReLUs]
//
// (reference: ConvBlock.forward, networks/universe/blocks.py:385-399, with PReLU_Conv.forward
// :205-227 for each conv and film :53-59).  Run as three separate ou_conv1d launches these layers
// are pure HBM traffic (10 tensor passes per block at 32-64 channels); here the two intermediate
// activations never leave the SM: ONE read of x (+ sc), ONE write of v.
//
// Work item = one window of 128*W accumulator rows of one clip (W = 64 / C), of which 128*W - 4 are
// valid outputs (halo of the k5 -> k3 -> k3 chain recomputed per window).  A persistent CTA (one per
// SM) keeps S items in flight in S "slots"; per slot
//     smem  X   : x rows [t0-4, t0+128W+4)  TMA-loaded (hardware swizzle, zero fill outside [0, T)),
//                 PReLU'd in place -> A operand of conv1; the raw rows are captured in registers
//                 first (the residual of the last stage)
//     smem  Cb  : sc rows (TMA) -> conv1 epilogue output c1 (in place) -> conv2 epilogue output c2
//                 (in place); always in the swizzled K-major layout = A operand of conv2 / conv3
//     TMEM      : 64 fp32 accumulator columns, reused by the three stages
// Warp roles: warp 0 loads the 11 weight taps once (TMA) and then issues every tcgen05.mma of the
// CTA, serving whichever slot has its operand ready (mbarrier.test_wait polling: the three stages
// of S slots interleave on the tensor pipe); warps 1.. form S warpgroups, one per slot, that run
// the slot's PReLU transform and its three epilogues (thread = accumulator row) and issue the
// slot's own TMA loads as soon as a tcgen05.commit has released the buffer.
#include <cuda.h>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ou {
namespace tc {
extern long long* g_trace;
}
namespace trunk {

using namespace ou::tc;

constexpr int S = 3;                          // item slots per CTA
// Warps per slot: 4 (one per TMEM lane quarter, a thread owns a whole accumulator row) or 8 (two column
// halves per quarter).  Measured on B200: 8 is 25 % SLOWER -- 25 warps leave only 72 registers per
// thread (spills), and shared memory is already the busiest unit: the N = 32 / 64 MMAs re-read their
// 4 KB A tile for every 16-wide k step (~260 KB of operand reads per item against ~115 KB of tile
// traffic).  ncu at C = 64: tensor pipe 33 % active, issue slots 36 %, i.e. what is left on the table
// is the latency of each slot's serial T0 -> MMA -> E1 -> MMA -> E2 -> MMA -> E3 chain; more slots would
// hide it but do not fit the 227 KB of shared memory next to the 11 resident weight taps.
constexpr int WPS = 4;
constexpr int NTHREADS = (1 + WPS * S) * 32;  // 416
constexpr int TAPS1 = 5, TAPS2 = 3, TAPS3 = 3, NTAPS = TAPS1 + TAPS2 + TAPS3;
constexpr int SLOT_COLS = 64;                 // TMEM columns per slot (= W * C)

struct TrunkArgs {
  long long* trace;   // debug (ou_debug_set_trace): CTA 0 stamps clock64(): [0..63][8] slot-0 items, then
                      // [64..127][8] MMA-warp issue log (event = slot * 3 + stage, round-robin over 64 rows)
  ou_trunk_params p;
  int items_per_clip, total_items;
  int x_box_rows, x_boxes;
  uint32_t idesc, desc_hi;
};

__device__ __forceinline__ bool mbar_test(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n"
      ".reg .pred P;\n"
      "mbarrier.test_wait.parity.shared::cta.b64 P, [%1], %2;\n"
      "selp.u32 %0, 1, 0, P;\n"
      "}\n"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}

template <int C>
struct Geo {
  static constexpr int W = 64 / C;                     // 128-row sub-tiles per item
  static constexpr int ROWB = C * 2;                   // bytes per activation row
  static constexpr int CH = C / 8;                     // 16-byte chunks per row
  static constexpr int K16 = C / 16;                   // k16 MMA steps per tap
  static constexpr int VALID = 128 * W - 4;            // valid output rows per item
  static constexpr int XPAD = 128 * W + 8;             // rows held per buffer (4 halo + 4 pad)
  static constexpr uint32_t BUF_BYTES = ((uint32_t)(XPAD * ROWB) + 1023u) & ~1023u;
  static constexpr uint32_t W_TAP_BYTES = (uint32_t)(C * ROWB);
  static constexpr uint32_t SWZ_MASK = ROWB == 128 ? 7u : 3u;
  static constexpr uint32_t W_BYTES = NTAPS * W_TAP_BYTES;
  static constexpr uint32_t TAIL_BYTES = 8u * (1 + 5 * S) + 16u + 4u * (S * 2 * C + 2 * C);
  static constexpr size_t SMEM = 1024 + W_BYTES + (size_t)S * 2 * BUF_BYTES + TAIL_BYTES;
};

// byte offset `off` inside a 1024-aligned buffer -> address in the hardware swizzle (16-byte chunk
// index XORed with the 128-byte line index, exactly what TMA writes and the MMA descriptors read)
template <int C>
__device__ __forceinline__ uint32_t swz(uint32_t base, uint32_t off) {
  return base + (off ^ (((off >> 7) & Geo<C>::SWZ_MASK) << 4));
}

struct Smem {
  uint32_t w;                    // NTAPS weight tiles
  uint32_t x[S], cb[S];
  uint32_t w_full;
  uint32_t x_full[S], sc_full[S], xp_ready[S], c_ready[S], acc_full[S];
  uint32_t tmem_slot;
  uint32_t coef1[S];             // fp32 [c0 | c1][C] of the slot's current clip (conv1 epilogue)
  uint32_t bias2, coef3;         // fp32 [C]: conv2 bias; s3 * conv3 bias
};

// ---------------------------------------------------------------------------------- MMA issue
template <int C, int TAPS>
__device__ __forceinline__ void issue_stage(uint32_t a_buf, uint32_t w_buf, uint32_t d_tmem,
                                            uint32_t acc_bar, uint32_t idesc, uint64_t hi64) {
  using G = Geo<C>;
  if (elect_one()) {
#pragma unroll
    for (int sub = 0; sub < G::W; sub++) {
#pragma unroll
      for (int q = 0; q < TAPS; q++) {
        const uint32_t a_lo = (((a_buf + (uint32_t)((sub * 128 + q) * G::ROWB)) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t b_lo = (((w_buf + (uint32_t)q * G::W_TAP_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
        for (int kk = 0; kk < G::K16; kk++)
          umma_f16(d_tmem + (uint32_t)(sub * C), hi64 | (a_lo + 2 * kk), hi64 | (b_lo + 2 * kk), idesc,
                   (q | kk) != 0 ? 1u : 0u);
      }
    }
    umma_commit(acc_bar);
  }
  __syncwarp();
}

template <int C>
__device__ __forceinline__ void mma_role(const TrunkArgs& a, const Smem& sm, uint32_t tmem_base,
                                         int n_items, const CUtensorMap* tm_w1, const CUtensorMap* tm_w2,
                                         const CUtensorMap* tm_w3, int lane) {
  using G = Geo<C>;
  if (lane == 0) {
    mbar_arrive_expect_tx(sm.w_full, G::W_BYTES);
    for (int q = 0; q < TAPS1; q++) tma_load_3d(sm.w + q * G::W_TAP_BYTES, tm_w1, 0, 0, q, sm.w_full);
    for (int q = 0; q < TAPS2; q++)
      tma_load_3d(sm.w + (TAPS1 + q) * G::W_TAP_BYTES, tm_w2, 0, 0, q, sm.w_full);
    for (int q = 0; q < TAPS3; q++)
      tma_load_3d(sm.w + (TAPS1 + TAPS2 + q) * G::W_TAP_BYTES, tm_w3, 0, 0, q, sm.w_full);
  }
  __syncwarp();
  mbar_wait(sm.w_full, 0);
  tc_fence_after();

  const uint64_t hi64 = (uint64_t)a.desc_hi << 32;
  const uint32_t idesc = a.idesc;
  int stage[S], left[S];
  uint32_t ph_xp[S], ph_c[S];
#pragma unroll
  for (int s = 0; s < S; s++) {
    stage[s] = 0;
    left[s] = n_items > s ? (n_items - s + S - 1) / S : 0;
    ph_xp[s] = 0, ph_c[s] = 0;
  }
  int remaining = 3 * n_items;
  int issued = 0;
  while (remaining > 0) {
    bool any = false;
#pragma unroll
    for (int s = 0; s < S; s++) {
      if (left[s] == 0) continue;
      const uint32_t d_tmem = tmem_base + (uint32_t)(s * SLOT_COLS);
      if (stage[s] == 0) {
        if (!mbar_test(sm.xp_ready[s], ph_xp[s])) continue;
        tc_fence_after();
        issue_stage<C, TAPS1>(sm.x[s], sm.w, d_tmem, sm.acc_full[s], idesc, hi64);
        ph_xp[s] ^= 1;
        stage[s] = 1;
      } else {
        if (!mbar_test(sm.c_ready[s], ph_c[s])) continue;
        tc_fence_after();
        if (stage[s] == 1) {
          issue_stage<C, TAPS2>(sm.cb[s], sm.w + TAPS1 * G::W_TAP_BYTES, d_tmem, sm.acc_full[s], idesc, hi64);
          stage[s] = 2;
        } else {
          issue_stage<C, TAPS3>(sm.cb[s], sm.w + (TAPS1 + TAPS2) * G::W_TAP_BYTES, d_tmem, sm.acc_full[s],
                                idesc, hi64);
          stage[s] = 0;
          left[s]--;
        }
        ph_c[s] ^= 1;
      }
      if (a.trace != nullptr && blockIdx.x == 0 && lane == 0 && issued >= 72 && issued < 72 + 64 * 8) {
        a.trace[512 + (issued - 72)] = clock64() * 16 + s * 4 + (stage[s] + 2) % 3;
      }
      issued++;
      remaining--;
      any = true;
    }
    if (!any) __nanosleep(20);
  }
}

// ---------------------------------------------------------------------------------- slot warpgroup
template <int NPRELU>
__device__ __forceinline__ float out_act(float y, float s1, float s2) {
  if (NPRELU > 0) y = prelu_f(y, s1);
  if (NPRELU > 1) y = prelu_f(y, s2);
  return y;
}

template <int C, bool HAS_SC, int NPRELU>
__device__ __forceinline__ void slot_role(const TrunkArgs& a, const Smem& sm, uint32_t tmem_base, int slot,
                                          int n_items, const CUtensorMap* tm_x, const CUtensorMap* tm_sc,
                                          int warp, int lane) {
  using G = Geo<C>;
  const ou_trunk_params& p = a.p;
  const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
  const int half = ((warp - 1) % WPS) >> 2;     // which part of the channels this thread handles
  const int row = quarter * 32 + lane;          // accumulator row inside a 128-row sub-tile
  const int wg_tid = (int)threadIdx.x - 32 - slot * (WPS * 32);
  const bool leader = row == 0 && half == 0;
  constexpr int HC = C / (WPS / 4);             // channels per thread
  constexpr int CHH = G::CH / (WPS / 4);        // 16-byte chunks per thread and row
  const int col0 = half * HC, ch0 = half * CHH;
  const int T = p.t;
  // per-slot addresses resolved once (the struct is indexed dynamically only here)
  const uint32_t X = sm.x[slot], Cb = sm.cb[slot];
  const uint32_t coef1 = sm.coef1[slot];
  const uint32_t bar_x = sm.x_full[slot], bar_sc = sm.sc_full[slot], bar_xp = sm.xp_ready[slot];
  const uint32_t bar_c = sm.c_ready[slot], bar_acc = sm.acc_full[slot];
  const uint32_t bias2 = sm.bias2, coef3 = sm.coef3;
  const float slope_in = p.prelu_in, slope_m1 = p.prelu_mid1, slope_m2 = p.prelu_mid2;
  const float s3 = p.scale3;
  uint32_t a_in_hi, a_in_lo;
  split_slope(slope_in, a_in_hi, a_in_lo);
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * SLOT_COLS);
  act_t* outp = (act_t*)p.out;

  auto item_pos = [&](int n, int& b, int& t0) {
    const int item = (int)blockIdx.x + (int)gridDim.x * n;
    b = item / a.items_per_clip;
    t0 = (item - b * a.items_per_clip) * G::VALID;
  };
  auto load_x = [&](int n) {
    int b, t0;
    item_pos(n, b, t0);
    mbar_arrive_expect_tx(bar_x, (uint32_t)(a.x_boxes * a.x_box_rows * G::ROWB));
    for (int k = 0; k < a.x_boxes; k++)
      tma_load_3d(X + (uint32_t)(k * a.x_box_rows * G::ROWB), tm_x, 0, t0 - 4 + k * a.x_box_rows, b,
                  bar_x);
  };
  auto load_sc = [&](int n) {
    int b, t0;
    item_pos(n, b, t0);
    mbar_arrive_expect_tx(bar_sc, (uint32_t)(G::W * 128 * G::ROWB));
    for (int sub = 0; sub < G::W; sub++)
      tma_load_3d(Cb + (uint32_t)(sub * 128 * G::ROWB), tm_sc, 0, t0 - 2 + sub * 128, b, bar_sc);
  };

  // rows [128W, 128W+8) of Cb are read by the last taps of conv2 / conv3 (results discarded): keep
  // them finite
  if (wg_tid < 8 * G::CH) sts_u4(Cb + (uint32_t)(128 * G::W * G::ROWB) + 16u * wg_tid, make_uint4(0, 0, 0, 0));
  fence_proxy_async();
  if (leader && slot < n_items) {
    load_x(slot);
    if (HAS_SC) load_sc(slot);
  }

  uint32_t ph_x = 0, ph_sc = 0, ph_acc = 0;
  int last_b = -1;
#define TRUNK_STAMP(ev)                                                                     \
  if (a.trace != nullptr && blockIdx.x == 0 && slot == 0 && leader && n / S >= 8 && n / S < 72) \
    a.trace[(n / S - 8) * 8 + (ev)] = clock64();
  for (int n = slot; n < n_items; n += S) {
    int b, t0;
    item_pos(n, b, t0);
    const bool has_next = n + S < n_items;

    // ---- conv1 epilogue coefficients of this clip: y = c0 * (acc + sc) + c1
    if (b != last_b && (last_b < 0 || (p.gamma != nullptr && p.film_bstride != 0))) {
      if (wg_tid < C) {
        float g = 1.f, be = 0.f;
        if (p.gamma != nullptr) {
          g = p.gamma[(size_t)b * p.film_bstride + wg_tid];
          be = p.beta[(size_t)b * p.film_bstride + wg_tid];
        }
        const float c0 = g * p.scale1;
        sts_f1(coef1 + 4u * wg_tid, c0);
        sts_f1(coef1 + 4u * (C + wg_tid), fmaf(c0, p.b1[wg_tid], be));
      }
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(WPS * 32) : "memory");
    }
    last_b = b;

    // ---- stage 0: capture the raw residual rows, PReLU the tile in place.  All loads of a phase are
    // issued before the first dependent instruction (the accessors are volatile asm and keep their
    // order): a slot's warpgroup is a serial chain, every exposed shared-memory / TMEM latency
    // is paid in full.
    uint4 res[G::W][CHH];
    TRUNK_STAMP(0)
    mbar_wait(bar_x, ph_x);
    ph_x ^= 1;
    TRUNK_STAMP(1)
    {
      uint4 halo[CHH];
#pragma unroll
      for (int sub = 0; sub < G::W; sub++)
#pragma unroll
        for (int c = 0; c < CHH; c++)
          res[sub][c] = lds_u4(swz<C>(X, (uint32_t)((sub * 128 + row + 4) * G::ROWB + (ch0 + c) * 16)));
      if (row < 4) {
#pragma unroll
        for (int c = 0; c < CHH; c++) halo[c] = lds_u4(swz<C>(X, (uint32_t)(row * G::ROWB + (ch0 + c) * 16)));
      }
#pragma unroll
      for (int sub = 0; sub < G::W; sub++)
#pragma unroll
        for (int c = 0; c < CHH; c++)
          sts_u4(swz<C>(X, (uint32_t)((sub * 128 + row + 4) * G::ROWB + (ch0 + c) * 16)),
                 prelu_act8(res[sub][c], a_in_hi, a_in_lo));
      if (row < 4) {
#pragma unroll
        for (int c = 0; c < CHH; c++)
          sts_u4(swz<C>(X, (uint32_t)(row * G::ROWB + (ch0 + c) * 16)), prelu_act8(halo[c], a_in_hi, a_in_lo));
      }
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_xp);
    TRUNK_STAMP(2)

    // ---- stage 1: c1 = PReLU(FiLM((conv1 + b1 + sc) * s1)) -> Cb (bf16, swizzled)
    mbar_wait(bar_acc, ph_acc);
    ph_acc ^= 1;
    tc_fence_after();
    TRUNK_STAMP(3)
    if (leader && has_next) load_x(n + S);      // conv1's MMAs are done with X
    if (HAS_SC) {
      mbar_wait(bar_sc, ph_sc);
      ph_sc ^= 1;
    }
    // The accumulator is read in NQ chunks of 16 columns, chunk q + 1 in flight (tcgen05.ld) while chunk
    // q is processed; with the per-chunk shared loads issued before the wait.
    constexpr int QS = HC / 16;   // chunks per sub-tile and thread
    constexpr int NQ = G::W * QS;
    auto q_taddr = [&](int q) { return taddr + (uint32_t)((q / QS) * C + col0 + (q % QS) * 16); };
    uint32_t rbuf[2][16];
    tmem_ld16(q_taddr(0), rbuf[0]);
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int sub = q / QS, cc = q % QS;
      const int i = sub * 128 + row;
      const int t = t0 - 2 + i;
      const bool inside = t >= 0 && t < T;
      float4 k0[4], k1[4];
      uint4 scv[2];
#pragma unroll
      for (int j = 0; j < 4; j++) {
        k0[j] = lds_f4(coef1 + 4u * (col0 + cc * 16 + j * 4));
        k1[j] = lds_f4(coef1 + 4u * (C + col0 + cc * 16 + j * 4));
      }
      if (HAS_SC) {
        scv[0] = lds_u4(swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc) * 16)));
        scv[1] = lds_u4(swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc + 1) * 16)));
      }
      tmem_ld_wait();
      if (q + 1 < NQ) tmem_ld16(q_taddr(q + 1), rbuf[(q + 1) & 1]);
      const uint32_t(&r)[16] = rbuf[q & 1];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const uint32_t* sw = reinterpret_cast<const uint32_t*>(&scv[h]);
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const float4 c0 = k0[h * 2 + k], c1 = k1[h * 2 + k];
          const int e = h * 8 + k * 4;
          float a0 = __uint_as_float(r[e]), a1 = __uint_as_float(r[e + 1]);
          float a2 = __uint_as_float(r[e + 2]), a3 = __uint_as_float(r[e + 3]);
          if (HAS_SC) {
            const float2 fa = act2_to_f2(sw[2 * k]), fb = act2_to_f2(sw[2 * k + 1]);
            a0 += fa.x, a1 += fa.y, a2 += fb.x, a3 += fb.y;
          }
          a0 = prelu_f(fmaf(c0.x, a0, c1.x), slope_m1), a1 = prelu_f(fmaf(c0.y, a1, c1.y), slope_m1);
          a2 = prelu_f(fmaf(c0.z, a2, c1.z), slope_m1), a3 = prelu_f(fmaf(c0.w, a3, c1.w), slope_m1);
          ow[2 * k] = inside ? f2_to_act2(a0, a1) : 0u;
          ow[2 * k + 1] = inside ? f2_to_act2(a2, a3) : 0u;
        }
        sts_u4(swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc + h) * 16)), o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_c);
    TRUNK_STAMP(4)

    // ---- stage 2: c2 = PReLU(conv2 + b2) -> Cb in place
    mbar_wait(bar_acc, ph_acc);
    ph_acc ^= 1;
    tc_fence_after();
    TRUNK_STAMP(5)
    tmem_ld16(q_taddr(0), rbuf[0]);
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int sub = q / QS, cc = q % QS;
      const int i = sub * 128 + row;
      const int t = t0 - 1 + i;
      const bool inside = t >= 0 && t < T && i < 128 * G::W - 2;
      float4 bb[4];
#pragma unroll
      for (int j = 0; j < 4; j++) bb[j] = lds_f4(bias2 + 4u * (col0 + cc * 16 + j * 4));
      tmem_ld_wait();
      if (q + 1 < NQ) tmem_ld16(q_taddr(q + 1), rbuf[(q + 1) & 1]);
      const uint32_t(&r)[16] = rbuf[q & 1];
#pragma unroll
      for (int h = 0; h < 2; h++) {
        uint4 o;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const float4 b4 = bb[h * 2 + k];
          const int e = h * 8 + k * 4;
          const float a0 = prelu_f(__uint_as_float(r[e]) + b4.x, slope_m2);
          const float a1 = prelu_f(__uint_as_float(r[e + 1]) + b4.y, slope_m2);
          const float a2 = prelu_f(__uint_as_float(r[e + 2]) + b4.z, slope_m2);
          const float a3 = prelu_f(__uint_as_float(r[e + 3]) + b4.w, slope_m2);
          ow[2 * k] = inside ? f2_to_act2(a0, a1) : 0u;
          ow[2 * k + 1] = inside ? f2_to_act2(a2, a3) : 0u;
        }
        sts_u4(swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc + h) * 16)), o);
      }
    }
    fence_proxy_async();
    tc_fence_before();
    __syncwarp();
    if (lane == 0) mbar_arrive(bar_c);
    TRUNK_STAMP(6)

    // ---- stage 3: v = (conv3 + b3 + x) * s3 -> PReLUs -> global
    mbar_wait(bar_acc, ph_acc);
    ph_acc ^= 1;
    tc_fence_after();
    TRUNK_STAMP(7)
    if (HAS_SC && leader && has_next) load_sc(n + S);   // conv3's MMAs are done with Cb
    tmem_ld16(q_taddr(0), rbuf[0]);
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int sub = q / QS, cc = q % QS;
      const int i = sub * 128 + row;
      const int t = t0 + i;
      const bool valid = i < G::VALID && t < T;
      act_t* dst = outp + ((size_t)b * T + t) * C;
      float4 kk[4];
#pragma unroll
      for (int j = 0; j < 4; j++) kk[j] = lds_f4(coef3 + 4u * (col0 + cc * 16 + j * 4));
      tmem_ld_wait();
      if (q + 1 < NQ) tmem_ld16(q_taddr(q + 1), rbuf[(q + 1) & 1]);
      const uint32_t(&r)[16] = rbuf[q & 1];
      U8 o;
#pragma unroll
      for (int h = 0; h < 2; h++) {
        const uint32_t* rw = reinterpret_cast<const uint32_t*>(&res[sub][2 * cc + h]);
#pragma unroll
        for (int k = 0; k < 2; k++) {
          const float4 c1 = kk[h * 2 + k];
          const int e = h * 8 + k * 4;
          const float2 fa = act2_to_f2(rw[2 * k]), fb = act2_to_f2(rw[2 * k + 1]);
          const float a0 = fmaf(s3, __uint_as_float(r[e]) + fa.x, c1.x);
          const float a1 = fmaf(s3, __uint_as_float(r[e + 1]) + fa.y, c1.y);
          const float a2 = fmaf(s3, __uint_as_float(r[e + 2]) + fb.x, c1.z);
          const float a3 = fmaf(s3, __uint_as_float(r[e + 3]) + fb.y, c1.w);
          o.w[h * 4 + 2 * k] = f2_to_act2(out_act<NPRELU>(a0, p.prelu_out, p.prelu_out2),
                                         out_act<NPRELU>(a1, p.prelu_out, p.prelu_out2));
          o.w[h * 4 + 2 * k + 1] = f2_to_act2(out_act<NPRELU>(a2, p.prelu_out, p.prelu_out2),
                                             out_act<NPRELU>(a3, p.prelu_out, p.prelu_out2));
        }
      }
      if (valid) stg_v8(dst + col0 + cc * 16, o);
    }
    tc_fence_before();   // TMEM reads ordered before the next item's MMAs (via xp_ready)
  }
}

// ---------------------------------------------------------------------------------- kernel
template <int C, bool HAS_SC>
__global__ void __launch_bounds__(NTHREADS, 1)
trunk_kernel(const TrunkArgs a, const __grid_constant__ CUtensorMap tm_x,
             const __grid_constant__ CUtensorMap tm_sc, const __grid_constant__ CUtensorMap tm_w1,
             const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_w3) {
  using G = Geo<C>;
  extern __shared__ uint8_t smem_raw[];
  const ou_trunk_params& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  Smem sm;
  uint32_t at = (smem_u32(smem_raw) + 1023u) & ~1023u;
  sm.w = at, at += G::W_BYTES;
#pragma unroll
  for (int s = 0; s < S; s++) {
    sm.x[s] = at, at += G::BUF_BYTES;
    sm.cb[s] = at, at += G::BUF_BYTES;
  }
  sm.w_full = at, at += 8;
#pragma unroll
  for (int s = 0; s < S; s++) {
    sm.x_full[s] = at, sm.sc_full[s] = at + 8, sm.xp_ready[s] = at + 16, sm.c_ready[s] = at + 24;
    sm.acc_full[s] = at + 32;
    at += 40;
  }
  sm.tmem_slot = at, at += 16;
#pragma unroll
  for (int s = 0; s < S; s++) sm.coef1[s] = at, at += 8u * C;
  sm.bias2 = at, at += 4u * C;
  sm.coef3 = at;

  if (threadIdx.x == 0) {
    mbar_init(sm.w_full, 1);
    for (int s = 0; s < S; s++) {
      mbar_init(sm.x_full[s], 1);
      mbar_init(sm.sc_full[s], 1);
      mbar_init(sm.xp_ready[s], WPS);
      mbar_init(sm.c_ready[s], WPS);
      mbar_init(sm.acc_full[s], 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < C) {
    sts_f1(sm.bias2 + 4u * threadIdx.x, p.b2[threadIdx.x]);
    sts_f1(sm.coef3 + 4u * threadIdx.x, p.scale3 * p.b3[threadIdx.x]);
  }
  if (warp == 0) tmem_alloc(sm.tmem_slot, 256);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot));

  // items blockIdx.x, blockIdx.x + gridDim.x, ... ; the n-th of them runs in slot n % S
  const int n_items = (a.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (warp == 0) {
    mma_role<C>(a, sm, tmem_base, n_items, &tm_w1, &tm_w2, &tm_w3, lane);
  } else {
    const int slot = (warp - 1) / WPS;
    const int nprelu = p.has_prelu_out2 ? 2 : (p.has_prelu_out ? 1 : 0);
    if (nprelu == 0)
      slot_role<C, HAS_SC, 0>(a, sm, tmem_base, slot, n_items, &tm_x, &tm_sc, warp, lane);
    else if (nprelu == 1)
      slot_role<C, HAS_SC, 1>(a, sm, tmem_base, slot, n_items, &tm_x, &tm_sc, warp, lane);
    else
      slot_role<C, HAS_SC, 2>(a, sm, tmem_base, slot, n_items, &tm_x, &tm_sc, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 256);
}

// ---------------------------------------------------------------------------------- host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int init_once() {
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
      set_error("ou_conv_trunk: cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return OU_ERR_CUDA;
    }
    g_encode = (EncodeTiledFn)fn;
  }
  return OU_OK;
}

static int encode3(CUtensorMap* tm, const void* base, int c, uint64_t d1, uint64_t d2, uint32_t box_rows,
                   CUtensorMapL2promotion promo, const char* what) {
  const cuuint64_t rowb = (cuuint64_t)c * 2;
  cuuint64_t dims[3] = {(cuuint64_t)c, d1, d2};
  cuuint64_t strides[2] = {rowb, rowb * d1};
  cuuint32_t box[3] = {(cuuint32_t)c, box_rows, 1};
  cuuint32_t estr[3] = {1, 1, 1};
  CUresult r = g_encode(tm, OU_TMA_ACT, 3, const_cast<void*>(base), dims, strides, box,
                        estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        rowb == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, promo,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("ou_conv_trunk: cuTensorMapEncodeTiled(%s) failed with %d", what, (int)r);
    return OU_ERR_CUDA;
  }
  return OU_OK;
}

template <int C>
static int launch_c(const ou_trunk_params* p, cudaStream_t st) {
  using G = Geo<C>;
  TrunkArgs a;
  a.trace = ou::tc::g_trace;
  a.p = *p;
  a.items_per_clip = ceil_div(p->t, G::VALID);
  a.total_items = a.items_per_clip * p->batch;
  // X rows per item: 128W + 8, fetched as equal boxes of <= 256 rows whose byte size keeps every
  // box start on a swizzle-pattern boundary (multiple of 8 rows)
  a.x_boxes = 1;
  while ((G::XPAD / a.x_boxes) > 256 || G::XPAD % a.x_boxes || (G::XPAD / a.x_boxes) % 8) a.x_boxes++;
  a.x_box_rows = G::XPAD / a.x_boxes;
  a.idesc = (1u << 4) | ((uint32_t)OU_ACT_IS_BF16 << 7) | ((uint32_t)OU_ACT_IS_BF16 << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t sbo = 8u * G::ROWB;
  const uint32_t layout = G::ROWB == 128 ? 2u : 4u;
  a.desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << (46 - 32)) | (layout << (61 - 32));

  CUtensorMap tm_x, tm_sc, tm_w1, tm_w2, tm_w3;
  int rc;
  if ((rc = encode3(&tm_x, p->x, C, (uint64_t)p->t, (uint64_t)p->batch, (uint32_t)a.x_box_rows,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "x")))
    return rc;
  if ((rc = encode3(&tm_sc, p->sc ? p->sc : p->x, C, (uint64_t)p->t, (uint64_t)p->batch, 128,
                    CU_TENSOR_MAP_L2_PROMOTION_L2_128B, "sc")))
    return rc;
  if ((rc = encode3(&tm_w1, p->w1, C, (uint64_t)C, TAPS1, C, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w1"))) return rc;
  if ((rc = encode3(&tm_w2, p->w2, C, (uint64_t)C, TAPS2, C, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w2"))) return rc;
  if ((rc = encode3(&tm_w3, p->w3, C, (uint64_t)C, TAPS3, C, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w3"))) return rc;

  auto kern = p->sc ? trunk_kernel<C, true> : trunk_kernel<C, false>;
  static SmemConfig cfg[2];
  if ((rc = ensure_smem(kern, (size_t)G::SMEM, cfg[p->sc ? 1 : 0], "ou_conv_trunk"))) return rc;
  const int n_sms = num_sms();
  int grid = n_sms < a.total_items ? n_sms : a.total_items;
  kern<<<grid, NTHREADS, G::SMEM, st>>>(a, tm_x, tm_sc, tm_w1, tm_w2, tm_w3);
  return check_launch("ou_conv_trunk");
}

}  // namespace trunk
}  // namespace ou

extern "C" int ou_conv_trunk(const ou_trunk_params* p, void* stream) {
  OU_REQUIRE(p != nullptr, "ou_conv_trunk: null params");
  OU_REQUIRE(p->x && p->w1 && p->w2 && p->w3 && p->b1 && p->b2 && p->b3 && p->out,
             "ou_conv_trunk: null pointer");
  OU_REQUIRE(p->batch > 0 && p->t > 0, "ou_conv_trunk: empty problem");
  OU_REQUIRE((p->gamma == nullptr) == (p->beta == nullptr), "ou_conv_trunk: gamma / beta must come together");
  if ((p->channels != 32 && p->channels != 64) || p->taps1 != ou::trunk::TAPS1 || p->taps2 != ou::trunk::TAPS2 ||
      p->taps3 != ou::trunk::TAPS3) {
    ou::set_error("ou_conv_trunk: only C in {32, 64} with a k5-k3-k3 chain has a fused kernel (got C=%d, k%d-k%d-k%d)",
                  p->channels, p->taps1, p->taps2, p->taps3);
    return OU_ERR_UNSUPPORTED;
  }
  int rc = ou::trunk::init_once();
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  return p->channels == 64 ? ou::trunk::launch_c<64>(p, st) : ou::trunk::launch_c<32>(p, st);
}
