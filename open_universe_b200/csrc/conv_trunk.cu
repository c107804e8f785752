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
// What bounds the kernel (measured, tools/ubench/mma_bench.cu): a tcgen05.mma of M = 128, K = 16 costs
// 67 cycles for ANY N <= 128 -- the A operand (128 rows x 32 B) is fetched from shared memory at 64 B per
// cycle -- so an item costs 44-48 MMAs x 67 cycles on the tensor pipe whatever the channel count: 177 /
// 187 us per launch at cfg-2 sizes for C = 64 / 32.  Epilogue arithmetic is packed fp32 (FFMA2 / FMUL2).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ou {
namespace tc {
extern long long* g_trace;
}
namespace trunk {

using namespace ou::tc;

constexpr int MAXS = 6;                       // most item slots per CTA any configuration uses
// Warps per slot: 4 (one per TMEM lane quarter, a thread owns a whole accumulator row) or 8 (two column
// halves per quarter).  Measured on B200: 8 is 25 % SLOWER -- 25 warps leave only 72 registers per
// thread (spills), and shared memory is already the busiest unit: the N = 32 / 64 MMAs re-read their
// 4 KB A tile for every 16-wide k step (~260 KB of operand reads per item against ~115 KB of tile
// traffic).  ncu at C = 64: tensor pipe 33 % active, issue slots 36 %, i.e. what is left on the table
// is the latency of each slot's serial T0 -> MMA -> E1 -> MMA -> E2 -> MMA -> E3 chain; more slots would
// hide it but do not fit the 227 KB of shared memory next to the 11 resident weight taps.
// Both are template parameters of the kernel (S slots x WPS warps); launch_c picks the configuration per
// channel count (OU_TRUNK_CFG32 / OU_TRUNK_CFG64 = "S,WPS" override it for A/B runs).
constexpr int TAPS1 = 5, TAPS2 = 3, TAPS3 = 3, NTAPS = TAPS1 + TAPS2 + TAPS3;
constexpr int SLOT_COLS = 64;                 // TMEM columns per slot (= W * C)

struct TrunkArgs {
  long long* trace;   // debug (ou_debug_set_trace): CTA 0 stamps clock64(): [0..31][16] slot-0 items (leader thread:
                      // start, x landed, T0 done, handed over, MMA1 issued, acc1, E1 done, handed over, MMA2 issued, ...)
  ou_trunk_params p;
  int items_per_clip, total_items;
  int x_box_rows, x_boxes;
  uint32_t idesc, desc_hi;
  uint32_t idesc_dn;    // down tail: N = 2C
};

// TAIL: 0 none; 1 the next block's x2 up conv (C = 64); 2 the network's output conv + EDM / SDE update (C = 32);
// 3 the block's own stride-2 anti-aliased down conv (C = 32)
template <int C, int S = 3, int TAIL = 0>
struct Geo {
  static constexpr bool UP = TAIL == 1, OUTC = TAIL == 2, DOWN = TAIL == 3;
  static constexpr int W = 64 / C;                     // 128-row sub-tiles per item
  static constexpr int ROWB = C * 2;                   // bytes per activation row
  static constexpr int CH = C / 8;                     // 16-byte chunks per row
  static constexpr int K16 = C / 16;                   // k16 MMA steps per tap
  static constexpr int VALID = 128 * W - 4;            // valid output rows per item
  static constexpr int XPAD = 128 * W + 8;             // rows held per buffer (4 halo + 4 pad)
  static constexpr uint32_t BUF_BYTES = ((uint32_t)(XPAD * ROWB) + 1023u) & ~1023u;
  static constexpr uint32_t W_TAP_BYTES = (uint32_t)(C * ROWB);
  static constexpr uint32_t SWZ_MASK = ROWB == 128 ? 7u : 3u;
  static constexpr int TAPS_UP = UP ? 3 : 0;                   // the fused up conv of the next block (tail)
  static constexpr int HALO = DOWN ? 2 : (TAIL ? 1 : 0);       // block-output rows an item recomputes on each side
  static constexpr int ITEM_VALID = VALID - 2 * HALO;          // rows an item contributes to the final output
  // down tail: 6 sample taps of [2C output channels][C] next to the block's own taps; the block output is kept
  // de-interleaved in Cb: even time steps from row 0, odd ones from row DN_ODD (each 126 rows + 2 of tap reach)
  static constexpr int DN_TAPS = 6, DN_ODD = 64 * W + 4;
  static constexpr uint32_t DN_TAP_BYTES = 2u * W_TAP_BYTES;
  static constexpr uint32_t W_BYTES = (NTAPS + TAPS_UP) * W_TAP_BYTES + (DOWN ? DN_TAPS * DN_TAP_BYTES : 0u);
  static constexpr uint32_t TAIL_BYTES = 8u * (1 + 5 * S) + 32u + 4u * (S * 2 * C + (OUTC ? 5 : (DOWN ? 4 : 3)) * C);   // barriers, TMEM slot, coefficients
  // no alignment slack: the dynamic shared window of a kernel without static shared memory starts 1024-byte
  // aligned (checked at run time) -- with it, 4 slots do not fit next to the 90 KB of weights at C = 64
  static constexpr size_t SMEM = W_BYTES + (size_t)S * 2 * BUF_BYTES + TAIL_BYTES;
};

// byte offset `off` inside a 1024-aligned buffer -> address in the hardware swizzle (16-byte chunk
// index XORed with the 128-byte line index, exactly what TMA writes and the MMA descriptors read)
template <int C>
__device__ __forceinline__ uint32_t swz(uint32_t base, uint32_t off) {
  return base + (off ^ (((off >> 7) & Geo<C>::SWZ_MASK) << 4));
}

struct Smem {
  uint32_t w;                    // NTAPS weight tiles
  uint32_t x[MAXS], cb[MAXS];
  uint32_t w_full;
  uint32_t x_full[MAXS], sc_full[MAXS], acc_full[MAXS];
  uint32_t tmem_slot;
  uint32_t coef1[MAXS];           // fp32 [c0 | c1][C] of the slot's current clip (conv1 epilogue)
  uint32_t bias2, coef3;         // fp32 [C]: conv2 bias; s3 * conv3 bias
  uint32_t coef_up;              // fp32 [C]: up_scale * up conv bias (tail)
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

// ---------------------------------------------------------------------------------- slot warpgroup
template <int C, int S, int WPS, bool HAS_SC, int NPRELU, bool FASTP, int TAIL>
__device__ __forceinline__ void slot_role(const TrunkArgs& a, const Smem& sm, uint32_t tmem_base, int slot,
                                          int n_items, const CUtensorMap* tm_x, const CUtensorMap* tm_sc,
                                          int warp, int lane) {
  using G = Geo<C, S, TAIL>;
  constexpr bool UP = TAIL == 1, OUTC = TAIL == 2, DOWN = TAIL == 3;
  const ou_trunk_params& p = a.p;
  const int quarter = warp & 3;                 // TMEM lane quarter this warp may read
  const int wis = warp % WPS;                   // warp index inside the slot's warpgroup
  const int half = wis >> 2;                    // which part of the channels this thread handles
  const int row = quarter * 32 + lane;          // accumulator row inside a 128-row sub-tile
  const int wg_tid = (int)threadIdx.x - slot * (WPS * 32);
  const bool issuer_warp = wis == 0;            // issues the slot's own tcgen05.mma (one elected lane)
  const uint64_t hi64 = (uint64_t)a.desc_hi << 32;
  const uint32_t idesc = a.idesc;
  const uint32_t d_tmem = tmem_base + (uint32_t)(slot * SLOT_COLS);
  const bool leader = row == 0 && half == 0;
  constexpr int HC = C / (WPS / 4);             // channels per thread
  constexpr int CHH = G::CH / (WPS / 4);        // 16-byte chunks per thread and row
  const int col0 = half * HC, ch0 = half * CHH;
  const int T = p.t;
  // per-slot addresses resolved once (the struct is indexed dynamically only here)
  const uint32_t X = sm.x[slot], Cb = sm.cb[slot];
  const uint32_t coef1 = sm.coef1[slot];
  const uint32_t bar_x = sm.x_full[slot], bar_sc = sm.sc_full[slot], bar_acc = sm.acc_full[slot];
  // hand-over of an operand tile from the warpgroup (generic-proxy stores, TMEM reads) to the slot's own
  // MMA issue: every thread fences, the warpgroup meets on its named barrier, the issuer warp goes on
  auto wg_handover = [&]() {
    fence_proxy_async();
    tc_fence_before();
    asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(WPS * 32) : "memory");
  };
  const uint32_t bias2 = sm.bias2, coef3 = sm.coef3;
  const float slope_in = p.prelu_in, slope_m1 = p.prelu_mid1, slope_m2 = p.prelu_mid2;
  const float s3 = p.scale3;
  uint32_t a_in_hi, a_in_lo;
  split_slope(slope_in, a_in_hi, a_in_lo);
  uint32_t a_dn_hi = 0, a_dn_lo = 0;
  if (DOWN) split_slope(p.dn_prelu_in, a_dn_hi, a_dn_lo);
  const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(slot * SLOT_COLS);
  act_t* outp = (act_t*)p.out;

  auto item_pos = [&](int n, int& b, int& t0) {
    const int item = (int)blockIdx.x + (int)gridDim.x * n;
    b = item / a.items_per_clip;
    // with the up tail, output row r of an item is low-rate index t0 + 1 + r (one row of halo on each side)
    t0 = (item - b * a.items_per_clip) * G::ITEM_VALID - G::HALO;
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

  uint32_t ph_x = 0, ph_acc = 0, ph_sc = 0;
  if (issuer_warp && slot < n_items) {   // the 11 weight taps (TMA, issued by thread 0) have landed
    mbar_wait(sm.w_full, 0);
    tc_fence_after();
  }
  int last_b = -1;
#define TRUNK_STAMP(ev)                                                                     \
  if (a.trace != nullptr && blockIdx.x == 0 && slot == 0 && leader && n / S >= 8 && n / S < 40) \
    a.trace[(n / S - 8) * 16 + (ev)] = clock64();
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
#pragma unroll
      for (int sub = 0; sub < G::W; sub++)
#pragma unroll
        for (int c = 0; c < CHH; c++)
          res[sub][c] = lds_u4(swz<C>(X, (uint32_t)((sub * 128 + row + 4) * G::ROWB + (ch0 + c) * 16)));
#pragma unroll
      for (int sub = 0; sub < G::W; sub++)
#pragma unroll
        for (int c = 0; c < CHH; c++)
          sts_u4(swz<C>(X, (uint32_t)((sub * 128 + row + 4) * G::ROWB + (ch0 + c) * 16)),
                 prelu_act8(res[sub][c], a_in_hi, a_in_lo));
      // the 4 halo rows in front of the tile, by the first 4 threads, chunk by chunk (their registers are
      // not worth keeping live next to the residual rows: that spilled in the variants with a tail)
      if (row < 4) {
#pragma unroll
        for (int c = 0; c < CHH; c++) {
          const uint32_t ha = swz<C>(X, (uint32_t)(row * G::ROWB + (ch0 + c) * 16));
          sts_u4(ha, prelu_act8(lds_u4(ha), a_in_hi, a_in_lo));
        }
      }
    }
    TRUNK_STAMP(2)
    wg_handover();
    TRUNK_STAMP(3)
    if (issuer_warp) {
      tc_fence_after();
      issue_stage<C, TAPS1>(X, sm.w, d_tmem, bar_acc, idesc, hi64);
    }
    TRUNK_STAMP(4)

    // ---- stage 1: c1 = PReLU(FiLM((conv1 + b1 + sc) * s1)) -> Cb (16-bit, swizzled)
    mbar_wait(bar_acc, ph_acc);
    ph_acc ^= 1;
    tc_fence_after();
    TRUNK_STAMP(5)
    if (leader && has_next) load_x(n + S);      // conv1's MMAs are done with X
    if (HAS_SC) {
      mbar_wait(bar_sc, ph_sc);
      ph_sc ^= 1;
    }
    // The accumulator is read in NQ chunks of 16 columns, chunk q + 1 in flight (tcgen05.ld) while chunk
    // q is processed; with the per-chunk shared loads issued before the wait.  Rows outside the clip
    // must read as zeros to the next conv ("same" padding of the intermediate activation): a per-row
    // branch (taken only in the two items at the ends of a clip), not a select per element.
    constexpr int QS = HC / 16;   // chunks per sub-tile and thread
    constexpr int NQ = G::W * QS;
    auto q_taddr = [&](int q) { return taddr + (uint32_t)((q / QS) * C + col0 + (q % QS) * 16); };
    const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
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
      const uint32_t dst = swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc) * 16));
      const uint32_t dst1 = swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc + 1) * 16));
      if (HAS_SC) scv[0] = lds_u4(dst), scv[1] = lds_u4(dst1);
      tmem_ld_wait();
      if (q + 1 < NQ) tmem_ld16(q_taddr(q + 1), rbuf[(q + 1) & 1]);
      const uint32_t(&r)[16] = rbuf[q & 1];
      if (inside) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          const uint32_t* sw = reinterpret_cast<const uint32_t*>(&scv[h]);
          uint4 o;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const float4 c0 = k0[h * 2 + k], c1 = k1[h * 2 + k];
            const int e = h * 8 + k * 4;
            float2 a01 = u2_as_f2(r[e], r[e + 1]), a23 = u2_as_f2(r[e + 2], r[e + 3]);
            if (HAS_SC) a01 = fadd2(a01, act2_to_f2(sw[2 * k])), a23 = fadd2(a23, act2_to_f2(sw[2 * k + 1]));
            const float2 v01 = prelu2<FASTP>(ffma2(a01, make_float2(c0.x, c0.y), make_float2(c1.x, c1.y)), slope_m1);
            const float2 v23 = prelu2<FASTP>(ffma2(a23, make_float2(c0.z, c0.w), make_float2(c1.z, c1.w)), slope_m1);
            ow[2 * k] = f2_to_act2(v01.x, v01.y);
            ow[2 * k + 1] = f2_to_act2(v23.x, v23.y);
          }
          sts_u4(h == 0 ? dst : dst1, o);
        }
      } else {
        sts_u4(dst, zero4);
        sts_u4(dst1, zero4);
      }
    }
    TRUNK_STAMP(6)
    wg_handover();
    TRUNK_STAMP(7)
    if (issuer_warp) {
      tc_fence_after();
      issue_stage<C, TAPS2>(Cb, sm.w + TAPS1 * G::W_TAP_BYTES, d_tmem, bar_acc, idesc, hi64);
    }
    TRUNK_STAMP(8)

    // ---- stage 2: c2 = PReLU(conv2 + b2) -> Cb in place
    mbar_wait(bar_acc, ph_acc);
    ph_acc ^= 1;
    tc_fence_after();
    TRUNK_STAMP(9)
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
      const uint32_t dst = swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc) * 16));
      const uint32_t dst1 = swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc + 1) * 16));
      if (inside) {
#pragma unroll
        for (int h = 0; h < 2; h++) {
          uint4 o;
          uint32_t* ow = reinterpret_cast<uint32_t*>(&o);
#pragma unroll
          for (int k = 0; k < 2; k++) {
            const float4 b4 = bb[h * 2 + k];
            const int e = h * 8 + k * 4;
            const float2 v01 = prelu2<FASTP>(fadd2(u2_as_f2(r[e], r[e + 1]), make_float2(b4.x, b4.y)), slope_m2);
            const float2 v23 = prelu2<FASTP>(fadd2(u2_as_f2(r[e + 2], r[e + 3]), make_float2(b4.z, b4.w)), slope_m2);
            ow[2 * k] = f2_to_act2(v01.x, v01.y);
            ow[2 * k + 1] = f2_to_act2(v23.x, v23.y);
          }
          sts_u4(h == 0 ? dst : dst1, o);
        }
      } else {
        sts_u4(dst, zero4);
        sts_u4(dst1, zero4);
      }
    }
    TRUNK_STAMP(10)
    wg_handover();
    TRUNK_STAMP(11)
    if (issuer_warp) {
      tc_fence_after();
      issue_stage<C, TAPS3>(Cb, sm.w + (TAPS1 + TAPS2) * G::W_TAP_BYTES, d_tmem, bar_acc, idesc, hi64);
    }
    TRUNK_STAMP(12)

    // ---- stage 3: v = (conv3 + b3 + x) * s3 -> PReLUs -> global
    mbar_wait(bar_acc, ph_acc);
    ph_acc ^= 1;
    tc_fence_after();
    TRUNK_STAMP(13)
    if (!TAIL && HAS_SC && leader && has_next) load_sc(n + S);   // conv3's MMAs are done with Cb
    tmem_ld16(q_taddr(0), rbuf[0]);
    const float2 s3v = make_float2(s3, s3);
#pragma unroll
    for (int q = 0; q < NQ; q++) {
      const int sub = q / QS, cc = q % QS;
      const int i = sub * 128 + row;
      const int t = t0 + i;
      const bool valid = i < G::VALID && t < T && t >= 0;
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
          float2 v01 = ffma2(fadd2(u2_as_f2(r[e], r[e + 1]), act2_to_f2(rw[2 * k])), s3v,
                             make_float2(c1.x, c1.y));
          float2 v23 = ffma2(fadd2(u2_as_f2(r[e + 2], r[e + 3]), act2_to_f2(rw[2 * k + 1])), s3v,
                             make_float2(c1.z, c1.w));
          if (NPRELU > 0) v01 = prelu2<FASTP>(v01, p.prelu_out), v23 = prelu2<FASTP>(v23, p.prelu_out);
          if (NPRELU > 1) v01 = prelu2<FASTP>(v01, p.prelu_out2), v23 = prelu2<FASTP>(v23, p.prelu_out2);
          if (UP) {
            // the block output is rounded to the storage type exactly as if it went through HBM, then the
            // up conv's input PReLU is applied (same rounding points as the separate launches)
            const float2 r01 = act2_to_f2(f2_to_act2(v01.x, v01.y)), r23 = act2_to_f2(f2_to_act2(v23.x, v23.y));
            v01 = prelu2<false>(r01, p.up_prelu_in), v23 = prelu2<false>(r23, p.up_prelu_in);
          }
          o.w[h * 4 + 2 * k] = f2_to_act2(v01.x, v01.y);
          o.w[h * 4 + 2 * k + 1] = f2_to_act2(v23.x, v23.y);
        }
      }
      if (DOWN) {
        // the block output goes to HBM (the decoder's skip connection; each row by the one item that owns it) AND,
        // through the down conv's input PReLU, into Cb as the tail's operand: even time steps from row 0, odd
        // ones from row DN_ODD, zero outside the clip (the low-pass's "same" padding)
        if (valid && i >= G::HALO && i < G::HALO + G::ITEM_VALID) stg_v8(dst + col0 + cc * 16, o);
        const int drow = (i >> 1) + ((i & 1) ? G::DN_ODD : 0);
        const uint32_t d0 = swz<C>(Cb, (uint32_t)(drow * G::ROWB + (ch0 + 2 * cc) * 16));
        const uint32_t d1 = swz<C>(Cb, (uint32_t)(drow * G::ROWB + (ch0 + 2 * cc + 1) * 16));
        sts_u4(d0, valid ? prelu_act8(make_uint4(o.w[0], o.w[1], o.w[2], o.w[3]), a_dn_hi, a_dn_lo) : zero4);
        sts_u4(d1, valid ? prelu_act8(make_uint4(o.w[4], o.w[5], o.w[6], o.w[7]), a_dn_hi, a_dn_lo) : zero4);
      } else if (TAIL) {
        // operand rows of the tail conv (zero outside the clip: its "same" padding)
        const uint32_t d0 = swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc) * 16));
        const uint32_t d1 = swz<C>(Cb, (uint32_t)(i * G::ROWB + (ch0 + 2 * cc + 1) * 16));
        sts_u4(d0, valid ? make_uint4(o.w[0], o.w[1], o.w[2], o.w[3]) : zero4);
        sts_u4(d1, valid ? make_uint4(o.w[4], o.w[5], o.w[6], o.w[7]) : zero4);
      } else if (valid) {
        stg_v8(dst + col0 + cc * 16, o);
      }
    }
    if (UP) {
      // ---- stage 4 (tail): the next block's transposed up conv + skip add on the tile just computed:
      // out[2 (t0 + 1 + r) + ph][co] = ((sum_q W_q[(ph, co)] . v[t0 + r + q]) + b_up + skip) * up_scale
      wg_handover();
      if (issuer_warp) {
        tc_fence_after();
        if (p.up_taps == 1) {   // plain k = s up conv: the middle tap of the 3-tap frame only
          if (elect_one()) {
            const uint32_t a_lo = (((Cb + (uint32_t)G::ROWB) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t b_lo = (((sm.w + (uint32_t)(NTAPS + 1) * G::W_TAP_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
            for (int kk = 0; kk < G::K16; kk++)
              umma_f16(d_tmem, hi64 | (a_lo + 2 * kk), hi64 | (b_lo + 2 * kk), idesc, kk != 0 ? 1u : 0u);
            umma_commit(bar_acc);
          }
          __syncwarp();
        } else {
          issue_stage<C, 3>(Cb, sm.w + (TAPS1 + TAPS2 + TAPS3) * G::W_TAP_BYTES, d_tmem, bar_acc, idesc, hi64);
        }
      }
      static_assert(!UP || (G::W == 1 && WPS == 4), "the up tail is built for C = 64 (one sub-tile, thread = row)");
      const int low = t0 + 1 + row;                    // low-rate index of this thread's output row
      const bool row_ok = row < G::ITEM_VALID && low >= 0 && low < T;
      // the thread's 64 output values are 2 consecutive time steps x 32 channels = 128 contiguous bytes of
      // the [t][C/2] output (and skip) layout: chunk q = elements [16 q, 16 q + 16), phase q / 2
      const size_t ubase = ((size_t)b * p.up_t_out + 2 * low) * (C / 2);
      const bool ok0 = row_ok && 2 * low < p.up_t_out, ok1 = row_ok && 2 * low + 1 < p.up_t_out;
      act_t* up_out = (act_t*)p.up_out + ubase;
      const act_t* up_skip = p.up_skip ? (const act_t*)p.up_skip + ubase : nullptr;
      U8 sk[4];
      if (up_skip != nullptr) {
#pragma unroll
        for (int q = 0; q < 4; q++)
          if (q < 2 ? ok0 : ok1) sk[q] = ldg_nc_v8(up_skip + 16 * q);
      }
      mbar_wait(bar_acc, ph_acc);
      ph_acc ^= 1;
      tc_fence_after();
      if (HAS_SC && leader && has_next) load_sc(n + S);   // the tail's MMAs are done with Cb
      tmem_ld16(q_taddr(0), rbuf[0]);
      const float2 usv = make_float2(p.up_scale, p.up_scale);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const bool ok = q < 2 ? ok0 : ok1;
        float4 kk[4];
#pragma unroll
        for (int j = 0; j < 4; j++) kk[j] = lds_f4(sm.coef_up + 4u * (q * 16 + j * 4));
        tmem_ld_wait();
        if (q + 1 < 4) tmem_ld16(q_taddr(q + 1), rbuf[(q + 1) & 1]);
        const uint32_t(&r)[16] = rbuf[q & 1];
        if (ok) {
          U8 o;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            float2 a01 = u2_as_f2(r[4 * k], r[4 * k + 1]), a23 = u2_as_f2(r[4 * k + 2], r[4 * k + 3]);
            if (up_skip != nullptr) {
              a01 = fadd2(a01, act2_to_f2(sk[q].w[2 * k]));
              a23 = fadd2(a23, act2_to_f2(sk[q].w[2 * k + 1]));
            }
            a01 = ffma2(a01, usv, make_float2(kk[k].x, kk[k].y));
            a23 = ffma2(a23, usv, make_float2(kk[k].z, kk[k].w));
            o.w[2 * k] = f2_to_act2(a01.x, a01.y);
            o.w[2 * k + 1] = f2_to_act2(a23.x, a23.y);
          }
          stg_v8(up_out + 16 * q, o);
        }
      }
      TRUNK_STAMP(15)
    }
    if (DOWN) {
      // ---- stage 4 (tail): the block's stride-2 down conv on the de-interleaved tile: output row jj of the item
      // is low-rate step (t0 + 2) / 2 + jj and reads block-output rows 2 jj .. 2 jj + 5 (sample taps m = 0..5):
      // even m from the even tile at row jj + m / 2, odd m from the odd tile at row jj + (m - 1) / 2
      static_assert(!DOWN || (C == 32 && WPS == 4 && SLOT_COLS >= 2 * C), "the down tail is built for C = 32");
      wg_handover();
      if (issuer_warp) {
        tc_fence_after();
        if (elect_one()) {
          const uint32_t wdn = sm.w + (uint32_t)NTAPS * G::W_TAP_BYTES;
          const int m_lo = p.dn_taps == 1 ? 2 : 0, m_hi = p.dn_taps == 1 ? 4 : G::DN_TAPS;   // k = s conv: middle row tap
#pragma unroll
          for (int m = 0; m < G::DN_TAPS; m++) {
            if (m < m_lo || m >= m_hi) continue;
            const int arow = (m >> 1) + ((m & 1) ? G::DN_ODD : 0);
            const uint32_t a_lo = (((Cb + (uint32_t)(arow * G::ROWB)) >> 4) & 0x3FFFu) | (1u << 16);
            const uint32_t b_lo = (((wdn + (uint32_t)m * G::DN_TAP_BYTES) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
            for (int kk = 0; kk < G::K16; kk++)
              umma_f16(d_tmem, hi64 | (a_lo + 2 * kk), hi64 | (b_lo + 2 * kk), a.idesc_dn, (m != m_lo || kk != 0) ? 1u : 0u);
          }
          umma_commit(bar_acc);
        }
        __syncwarp();
      }
      const int j = ((t0 + G::HALO) >> 1) + row;       // low-rate index of this thread's output row
      const bool ok = row < G::ITEM_VALID / 2 && j < p.dn_t_out;
      act_t* dn_out = (act_t*)p.dn_out + ((size_t)b * p.dn_t_out + j) * (2 * C);
      mbar_wait(bar_acc, ph_acc);
      ph_acc ^= 1;
      tc_fence_after();
      if (HAS_SC && leader && has_next) load_sc(n + S);   // the tail's MMAs are done with Cb
      tmem_ld16(taddr, rbuf[0]);
#pragma unroll
      for (int q = 0; q < 2 * C / 16; q++) {
        float4 kk[4];
#pragma unroll
        for (int jv = 0; jv < 4; jv++) kk[jv] = lds_f4(sm.coef_up + 4u * (q * 16 + jv * 4));
        tmem_ld_wait();
        if (q + 1 < 2 * C / 16) tmem_ld16(taddr + (uint32_t)((q + 1) * 16), rbuf[(q + 1) & 1]);
        const uint32_t(&r)[16] = rbuf[q & 1];
        if (ok) {
          U8 o;
#pragma unroll
          for (int k = 0; k < 4; k++) {
            const float2 a01 = fadd2(u2_as_f2(r[4 * k], r[4 * k + 1]), make_float2(kk[k].x, kk[k].y));
            const float2 a23 = fadd2(u2_as_f2(r[4 * k + 2], r[4 * k + 3]), make_float2(kk[k].z, kk[k].w));
            o.w[2 * k] = f2_to_act2(a01.x, a01.y);
            o.w[2 * k + 1] = f2_to_act2(a23.x, a23.y);
          }
          stg_v8(dn_out + 16 * q, o);
        }
      }
      TRUNK_STAMP(15)
    }
    if (OUTC) {
      // ---- tail: the network's output conv (C -> 1, k = 3) on the block output just written to Cb, fused with
      // the EDM mix and the reverse-SDE update (ou_output_sde): thread = output time step t0 + 1 + r
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(WPS * 32) : "memory");
      static_assert(!OUTC || (C == 32 && WPS == 4), "the output tail is built for C = 32");
      const float ca = p.out_coef ? p.out_coef[b * 3] : 0.f, cbv = p.out_coef ? p.out_coef[b * 3 + 1] : 0.f;
      const float ccv = p.out_coef ? p.out_coef[b * 3 + 2] : 0.f;
      // a thread owns rows `row` and `row + 128` of the item: each weight vector is fetched once for both
      float2 acc[G::W];
      float xv[G::W], zv[G::W];
#pragma unroll
      for (int sub = 0; sub < G::W; sub++) {
        const int r = sub * 128 + row, tt = t0 + 1 + r;
        acc[sub] = make_float2(0.f, 0.f), xv[sub] = zv[sub] = 0.f;
        if (r < G::ITEM_VALID && tt >= 0 && tt < T && p.out_coef) {
          xv[sub] = __ldg(p.out_x + (size_t)b * T + tt);
          if (p.out_noise) zv[sub] = __ldg(p.out_noise + (size_t)b * T + tt);
        }
      }
#pragma unroll 1
      for (int k = 0; k < 3; k++) {
#pragma unroll
        for (int c = 0; c < G::CH; c++) {
          const float4 w0 = lds_f4(sm.coef_up + 4u * (k * C + c * 8)), w1 = lds_f4(sm.coef_up + 4u * (k * C + c * 8 + 4));
#pragma unroll
          for (int sub = 0; sub < G::W; sub++) {
            const uint4 v = lds_u4(swz<C>(Cb, (uint32_t)((sub * 128 + row + k) * G::ROWB + c * 16)));
            acc[sub] = ffma2(make_float2(w0.x, w0.y), act2_to_f2(v.x), acc[sub]);
            acc[sub] = ffma2(make_float2(w0.z, w0.w), act2_to_f2(v.y), acc[sub]);
            acc[sub] = ffma2(make_float2(w1.x, w1.y), act2_to_f2(v.z), acc[sub]);
            acc[sub] = ffma2(make_float2(w1.z, w1.w), act2_to_f2(v.w), acc[sub]);
          }
        }
      }
#pragma unroll
      for (int sub = 0; sub < G::W; sub++) {
        const int r = sub * 128 + row, tt = t0 + 1 + r;
        const float net = p.out_bias + (acc[sub].x + acc[sub].y);
        if (r < G::ITEM_VALID && tt >= 0 && tt < T) {
          const size_t o = (size_t)b * T + tt;
          if (p.out_net) p.out_net[o] = net;
          if (p.out_coef) p.out_xout[o] = fmaf(ccv, zv[sub], fmaf(cbv, net, ca * xv[sub]));
        }
      }
      // every thread is done reading Cb: the next item's conditioning rows may land
      asm volatile("bar.sync %0, %1;" ::"r"(1 + slot), "n"(WPS * 32) : "memory");
      if (HAS_SC && leader && has_next) load_sc(n + S);
    }
    // (the TMEM reads above are ordered before the next item's MMAs by wg_handover() after its T0)
    TRUNK_STAMP(14)
  }
}

// ---------------------------------------------------------------------------------- kernel
template <int C, bool HAS_SC, int S, int WPS, bool FASTP, int TAIL>
__global__ void __launch_bounds__(WPS * S * 32, 1)
trunk_kernel(const TrunkArgs a, const __grid_constant__ CUtensorMap tm_x,
             const __grid_constant__ CUtensorMap tm_sc, const __grid_constant__ CUtensorMap tm_w1,
             const __grid_constant__ CUtensorMap tm_w2, const __grid_constant__ CUtensorMap tm_w3,
             const __grid_constant__ CUtensorMap tm_wup) {
  using G = Geo<C, S, TAIL>;
  constexpr bool UP = TAIL == 1, OUTC = TAIL == 2, DOWN = TAIL == 3;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const ou_trunk_params& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  Smem sm;
  uint32_t at = smem_u32(smem_raw);
  if ((at & 1023u) != 0u) __trap();   // the swizzled tiles need 1024-byte aligned bases
  sm.w = at, at += G::W_BYTES;
#pragma unroll
  for (int s = 0; s < S; s++) {
    sm.x[s] = at, at += G::BUF_BYTES;
    sm.cb[s] = at, at += G::BUF_BYTES;
  }
  sm.w_full = at, at += 8;
#pragma unroll
  for (int s = 0; s < S; s++) {
    sm.x_full[s] = at, sm.sc_full[s] = at + 8;
    sm.acc_full[s] = at + 32;
    at += 40;
  }
  sm.tmem_slot = at, at += 16;
  at = (at + 15u) & ~15u;       // the coefficient vectors are read as float4
#pragma unroll
  for (int s = 0; s < S; s++) sm.coef1[s] = at, at += 8u * C;
  sm.bias2 = at, at += 4u * C;
  sm.coef3 = at, at += 4u * C;
  sm.coef_up = at;

  if (threadIdx.x == 0) {
    mbar_init(sm.w_full, 1);
    for (int s = 0; s < S; s++) {
      mbar_init(sm.x_full[s], 1);
      mbar_init(sm.sc_full[s], 1);
      mbar_init(sm.acc_full[s], 1);
    }
    fence_barrier_init();
  }
  if (threadIdx.x < C) {
    sts_f1(sm.bias2 + 4u * threadIdx.x, p.b2[threadIdx.x]);
    sts_f1(sm.coef3 + 4u * threadIdx.x, p.scale3 * p.b3[threadIdx.x]);
    if (UP) sts_f1(sm.coef_up + 4u * threadIdx.x, p.up_scale * (p.up_bias ? p.up_bias[threadIdx.x] : 0.f));
  }
  if (DOWN && threadIdx.x < 2 * C) sts_f1(sm.coef_up + 4u * threadIdx.x, p.dn_bias ? p.dn_bias[threadIdx.x] : 0.f);
  // output tail: the C x 3 fp32 weights of the output conv, tap-major [k][C]
  if (OUTC && threadIdx.x < 3 * C) sts_f1(sm.coef_up + 4u * threadIdx.x, p.out_w[threadIdx.x]);
  constexpr uint32_t TMEM_COLS = S * SLOT_COLS <= 256 ? 256u : 512u;
  if (warp == 0) tmem_alloc(sm.tmem_slot, TMEM_COLS);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  uint32_t tmem_base;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(tmem_base) : "r"(sm.tmem_slot));

  // items blockIdx.x, blockIdx.x + gridDim.x, ... ; the n-th of them runs in slot n % S
  const int n_items = (a.total_items - (int)blockIdx.x + (int)gridDim.x - 1) / (int)gridDim.x;

  if (threadIdx.x == 0) {
    mbar_arrive_expect_tx(sm.w_full, G::W_BYTES);
    for (int q = 0; q < TAPS1; q++) tma_load_3d(sm.w + q * G::W_TAP_BYTES, &tm_w1, 0, 0, q, sm.w_full);
    for (int q = 0; q < TAPS2; q++)
      tma_load_3d(sm.w + (TAPS1 + q) * G::W_TAP_BYTES, &tm_w2, 0, 0, q, sm.w_full);
    for (int q = 0; q < TAPS3; q++)
      tma_load_3d(sm.w + (TAPS1 + TAPS2 + q) * G::W_TAP_BYTES, &tm_w3, 0, 0, q, sm.w_full);
    for (int q = 0; q < G::TAPS_UP; q++)
      tma_load_3d(sm.w + (NTAPS + q) * G::W_TAP_BYTES, &tm_wup, 0, 0, q, sm.w_full);
    if (DOWN)
      for (int m = 0; m < G::DN_TAPS; m++)
        tma_load_3d(sm.w + NTAPS * G::W_TAP_BYTES + m * G::DN_TAP_BYTES, &tm_wup, 0, 0, m, sm.w_full);
  }
  __syncwarp();
  {
    const int slot = warp / WPS;
    const int nprelu = p.has_prelu_out2 ? 2 : (p.has_prelu_out ? 1 : 0);
    if (nprelu == 0)
      slot_role<C, S, WPS, HAS_SC, 0, FASTP, TAIL>(a, sm, tmem_base, slot, n_items, &tm_x, &tm_sc, warp, lane);
    else if (nprelu == 1)
      slot_role<C, S, WPS, HAS_SC, 1, FASTP, TAIL>(a, sm, tmem_base, slot, n_items, &tm_x, &tm_sc, warp, lane);
    else
      slot_role<C, S, WPS, HAS_SC, 2, FASTP, TAIL>(a, sm, tmem_base, slot, n_items, &tm_x, &tm_sc, warp, lane);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
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

template <int C, int S, int WPS, int TAIL = 0>
static int launch_cfg(const ou_trunk_params* p, cudaStream_t st) {
  using G = Geo<C, S, TAIL>;
  constexpr bool UP = TAIL == 1, DOWN = TAIL == 3;
  static_assert(G::SMEM <= 232448, "trunk_kernel: shared-memory budget exceeds the 227 KB a CTA may opt into");
  TrunkArgs a;
  a.trace = ou::tc::g_trace;
  a.p = *p;
  a.items_per_clip = ceil_div(p->t, G::ITEM_VALID);
  a.total_items = a.items_per_clip * p->batch;
  // X rows per item: 128W + 8, fetched as equal boxes of <= 256 rows whose byte size keeps every
  // box start on a swizzle-pattern boundary (multiple of 8 rows)
  a.x_boxes = 1;
  while ((G::XPAD / a.x_boxes) > 256 || G::XPAD % a.x_boxes || (G::XPAD / a.x_boxes) % 8) a.x_boxes++;
  a.x_box_rows = G::XPAD / a.x_boxes;
  a.idesc = (1u << 4) | ((uint32_t)OU_ACT_IS_BF16 << 7) | ((uint32_t)OU_ACT_IS_BF16 << 10) | ((uint32_t)(C >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  a.idesc_dn = (1u << 4) | ((uint32_t)OU_ACT_IS_BF16 << 7) | ((uint32_t)OU_ACT_IS_BF16 << 10) | ((uint32_t)((2 * C) >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
  const uint32_t sbo = 8u * G::ROWB;
  const uint32_t layout = G::ROWB == 128 ? 2u : 4u;
  a.desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << (46 - 32)) | (layout << (61 - 32));

  CUtensorMap tm_x, tm_sc, tm_w1, tm_w2, tm_w3, tm_wup;
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
  if (DOWN) {
    if ((rc = encode3(&tm_wup, p->dn_w, C, (uint64_t)(2 * C), G::DN_TAPS, 2 * C, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w_dn")))
      return rc;
  } else if ((rc = encode3(&tm_wup, UP ? p->up_w : p->w3, C, (uint64_t)C, 3, C, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, "w_up")))
    return rc;

  // FASTP: every epilogue PReLU slope in [0, 1] -> PReLU(x) = max(x, a x)
  auto in01 = [](float a) { return a >= 0.f && a <= 1.f; };
  const bool fastp = in01(p->prelu_mid1) && in01(p->prelu_mid2) && (!p->has_prelu_out || in01(p->prelu_out)) &&
                     (!p->has_prelu_out2 || in01(p->prelu_out2));
  auto kern = p->sc ? (fastp ? trunk_kernel<C, true, S, WPS, true, TAIL> : trunk_kernel<C, true, S, WPS, false, TAIL>)
                    : (fastp ? trunk_kernel<C, false, S, WPS, true, TAIL> : trunk_kernel<C, false, S, WPS, false, TAIL>);
  static SmemConfig cfg[4];
  if ((rc = ensure_smem(kern, (size_t)G::SMEM, cfg[(p->sc ? 1 : 0) + (fastp ? 2 : 0)], "ou_conv_trunk"))) return rc;
  const int n_sms = p->max_ctas > 0 && p->max_ctas < num_sms() ? p->max_ctas : num_sms();
  int grid = n_sms < a.total_items ? n_sms : a.total_items;
  kern<<<grid, WPS * S * 32, G::SMEM, st>>>(a, tm_x, tm_sc, tm_w1, tm_w2, tm_w3, tm_wup);
  return check_launch("ou_conv_trunk");
}

// Slot configuration: S = 3 slots x WPS = 4 warps.  Measured dead ends (profiles/README.md, round 2): 5 x 4 at
// C = 32 (21 warps -> 80 registers, spills: 27 % slower), 2 x 8 and 4 x 4 (17 warps put 5 on one scheduler
// -> 96 registers); the kernel is bound by instruction issue, more warps in flight do not help.
template <int C>
static int launch_c(const ou_trunk_params* p, cudaStream_t st) {
  if constexpr (C == 64) {
    if (p->up_w != nullptr) return launch_cfg<C, 3, 4, 1>(p, st);   // fused up-conv tail
    // 4 slots fit exactly (no alignment slack) next to the 90 KB of weights; since the halo rows are
    // transformed chunk by chunk the kernel does not spill at 128 registers: 234 / 262 us against 250 / 265 us
    // with 3 slots (enc / dec, cfg-2 sizes); OU_TRUNK_S64=3 for A/B runs
    static const int s64 = [] { const char* e = getenv("OU_TRUNK_S64"); return e ? atoi(e) : 4; }();
    if (s64 == 4) return launch_cfg<C, 4, 4>(p, st);
  }
  if constexpr (C == 32) {
    if (p->out_w != nullptr) return launch_cfg<C, 4, 4, 2>(p, st);   // fused output conv + SDE update
    if (p->dn_w != nullptr) return launch_cfg<C, 4, 4, 3>(p, st);    // fused stride-2 down conv
    // 4 slots fit next to the 22 KB of weights at C = 32 (16 warps -> 128 registers): measured 213 / 237 us
    // against 233 / 245 us with 3 slots (enc / dec, cfg-2 sizes); OU_TRUNK_S32=3 for A/B runs
    static const int s32 = [] { const char* e = getenv("OU_TRUNK_S32"); return e ? atoi(e) : 4; }();
    if (s32 == 4) return launch_cfg<C, 4, 4>(p, st);
  }
  return launch_cfg<C, 3, 4>(p, st);
}

}  // namespace trunk
}  // namespace ou

extern "C" int ou_conv_trunk(const ou_trunk_params* p, void* stream) {
  OU_REQUIRE(p != nullptr, "ou_conv_trunk: null params");
  OU_REQUIRE(p->x && p->w1 && p->w2 && p->w3 && p->b1 && p->b2 && p->b3 && (p->out || p->up_w || p->out_w),
             "ou_conv_trunk: null pointer");
  OU_REQUIRE(p->out_w == nullptr || (p->channels == 32 && p->up_w == nullptr && (p->out_net || p->out_coef) &&
                                     (!p->out_coef || (p->out_x && p->out_xout))),
             "ou_conv_trunk: the fused output conv needs C = 32, no up tail, and net_out or (coef, x, xout)");
  OU_REQUIRE(p->up_w == nullptr || (p->channels == 64 && p->up_out != nullptr && p->up_t_out > 0 &&
                                    p->up_t_out <= 2 * p->t && !p->has_prelu_out && !p->has_prelu_out2),
             "ou_conv_trunk: the fused up-conv tail needs C = 64, an output buffer of at most 2 t samples and no "
             "output PReLU on the block");
  OU_REQUIRE(p->dn_w == nullptr || (p->channels == 32 && p->up_w == nullptr && p->out_w == nullptr && p->out != nullptr &&
                                    p->dn_out != nullptr && p->dn_t_out == (p->t + 1) / 2 && !p->has_prelu_out &&
                                    !p->has_prelu_out2),
             "ou_conv_trunk: the fused down conv needs C = 32, no other tail, the block output buffer, an output of "
             "ceil(t / 2) steps and no output PReLU on the block");
  OU_REQUIRE((p->up_taps == 0 || p->up_taps == 1 || p->up_taps == 3) && (p->dn_taps == 0 || p->dn_taps == 1 || p->dn_taps == 3),
             "ou_conv_trunk: up_taps / dn_taps must be 1 or 3 (0 = 3)");
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
