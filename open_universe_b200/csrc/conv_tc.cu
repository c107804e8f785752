// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d for sm_100a -- the production path of
// ou_conv1d for stride-1 input geometry (conv1/conv2/conv3, transposed up convs, 1x1s, GRU input
// projections: > 80 % of the FLOPs of a score step).  Contract: include/ou_b200.h.
//
// Persistent, warp-specialised CTA (one per SM, 448 threads):
//   warp 0       producer   : TMA tensor loads (cp.async.bulk.tensor, UTMALDG) global -> smem rings
//                             in the hardware 128B/64B/32B swizzle; rows outside the sequence ("same"
//                             padding, tile overrun) are zero-filled by the TMA unit
//   warp 1       MMA issuer : tcgen05.mma.cta_group::1.kind::f16 (bf16 x bf16 -> fp32 in TMEM) issued
//                             by one lane from straight-line code (taps / K steps are template
//                             parameters: the issue loop is pure scalar latency); tcgen05.commit frees
//                             smem stages and publishes accumulators; owns the TMEM allocation
//   warps 2-5    transform  : only for layers with a fused input PReLU: element-wise pass over each
//                             landed A stage, fence.proxy.async, hand-over to the MMA warp
//   warps 6-13   epilogue   : tcgen05.ld accumulator rows -> registers; every per-column constant
//                             (bias, FiLM gamma/beta, residual scales) is folded into three smem
//                             coefficient vectors so a column costs  y = c0*(acc+add1) + c2*add2 + c1;
//                             residual vectors are prefetched before the accumulator is ready;
//                             32-byte stores in the blocked layout (16-bit, or fp32 [B][N/16][rows][16])
// TMEM holds two accumulator buffers (2 x BN columns): the epilogue of tile i overlaps the MMAs of
// tile i+1.
//
// Operand layout: the blocked activation layout [B][C/CB][T][CB] makes one time step of one
// channel block a contiguous CB*2-byte row, i.e. a K-major operand row; a TMA box of (CB channels x
// 128+taps-1 time steps) lands as the canonical swizzled K-major tile.  A conv tap is a shift by
// whole rows: the A descriptor's start address advances by q rows (the hardware swizzle is a
// function of the shared-memory address bits, so base_offset stays 0 -- verified on B200 against
// the fp32 reference kernel for the 128B, 64B and 32B modes), and ONE staged tile feeds all taps.
// Weights are pre-packed per (tap, K block) as [npad][CB] K-major tiles; a CTA's N-slice stays
// resident in shared memory across all its M tiles when it fits (C <= 128), otherwise it streams
// from L2 through a ring.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"
#include "tc_ptx.cuh"

namespace ou {
namespace tc {

constexpr int BM = 128;
constexpr int XF_WARP0 = 2, EPI_WARP0 = 6, N_EPI_WARPS = 12;
constexpr int NTHREADS = (EPI_WARP0 + N_EPI_WARPS) * 32;   // 576
constexpr int MAX_A_STAGES = 8, MAX_B_STAGES = 16;

struct TcArgs {
  ou_conv_params p;
  int cb;              // channels per K block (= channel block of the input layout: 64 / 32 / 16)
  int row_bytes;       // cb * 2
  int n_kblocks;       // s * cin / cb  (K block kb <-> input phase r = kb / cin_blocks, channel block kb % cin_blocks)
  int cin_blocks;      // cin / cb
  int arows;           // rows per A stage: BM + taps - 1
  int m_sub;           // 128-row sub-tiles per scheduling unit (2 for narrow layers: halves the
                       // per-tile synchronisation cost that dominates when a tile is only a few MMAs)
  uint32_t a_sub_bytes;  // 1024-aligned bytes of one sub-tile inside an A stage
  int bn;              // N per CTA tile
  int n_ntiles;        // npad / bn
  int m_tiles;         // scheduling units per clip: ceil(rows / (BM * m_sub))
  int total_m_tiles;   // m_tiles * batch
  int ctas_per_ntile;  // gridDim.x / n_ntiles
  int a_stages, b_stages;
  int resident;        // whole weight slice stays in smem (b_stages == taps * n_kblocks)
  uint32_t a_stage_bytes, b_stage_bytes;   // 1024-aligned strides
  uint32_t a_tx_bytes, b_tx_bytes;         // bytes one TMA box delivers
  uint32_t idesc;
  uint32_t tmem_cols;
  uint32_t desc_hi;    // descriptor bits [32,64): SBO, version, layout type (base_offset = 0)
  long long* trace;    // debug: [role 0..3][tile 0..63][event 0..3] clock64 stamps of CTA 0 (or NULL)
};

// Everything the warp roles share, resolved once per CTA (shared-memory addresses are 32-bit).
struct Ctx {
  uint32_t smA, smB;                         // stage rings
  uint32_t full_a, ready_a, empty_a;         // barrier arrays (8 bytes per stage)
  uint32_t full_b, empty_b;
  uint32_t tmem_full, tmem_empty;            // [2]
  uint32_t coef;                             // fp32 [2 buffers][c0 | c1 | c2][bn]
  const float* coef_ptr;                     // the same array as a generic pointer (plain C++ loads)
  const int* coltab_ptr;                     // int2 [16]: per 16-column group (output offset or -1, phase)
  uint32_t coltab;
  uint32_t tmem_base;
  int nt, n0, mt0, mt_stride;
};

__device__ __forceinline__ void trace_ev(const TcArgs& a, int role, int tile_i, int ev) {
  if (a.trace != nullptr && blockIdx.x == 0 && tile_i < 64)
    a.trace[(role * 64 + tile_i) * 4 + ev] = clock64();
}

__device__ __forceinline__ void epi_bar_sync(int n_threads) {
  asm volatile("bar.sync 1, %0;" ::"r"(n_threads) : "memory");
}

// ================================================================================ producer
__device__ __forceinline__ void producer_role(const TcArgs& a, const Ctx& c, const CUtensorMap* tm_a,
                                              const CUtensorMap* tm_w) {
  const ou_conv_params& p = a.p;
  Ring ra, rb;
  bool b_loaded = false;
  const int taps = p.taps;
  int ti = 0;
  for (int mt = c.mt0; mt < a.total_m_tiles; mt += c.mt_stride, ti++) {
    const int b = mt / a.m_tiles;
    const int m0 = (mt - b * a.m_tiles) * BM * a.m_sub;
    for (int kb = 0; kb < a.n_kblocks; kb++) {
      mbar_wait(c.empty_a + 8u * ra.stage, ra.phase ^ 1);
      if (kb == 0) trace_ev(a, 0, ti, 0);
      mbar_arrive_expect_tx(c.full_a + 8u * ra.stage, a.a_tx_bytes * a.m_sub);
      // rows j in [m0 + tap_off, +arows) <-> time steps j*s + r of clip b, channel block kbi;
      // the TMA unit zero-fills rows outside [0, T/s)
      const int r = kb / a.cin_blocks, kbi = kb - r * a.cin_blocks;
      for (int sub = 0; sub < a.m_sub; sub++)
        tma_load_5d(c.smA + ra.stage * a.a_stage_bytes + sub * a.a_sub_bytes, tm_a, 0, r,
                    m0 + sub * BM + p.tap_off, kbi, b, c.full_a + 8u * ra.stage);
      ra.advance(a.a_stages);
      if (!(a.resident && b_loaded)) {
        for (int q = 0; q < taps; q++) {
          mbar_wait(c.empty_b + 8u * rb.stage, rb.phase ^ 1);
          mbar_arrive_expect_tx(c.full_b + 8u * rb.stage, a.b_tx_bytes);
          tma_load_3d(c.smB + rb.stage * a.b_stage_bytes, tm_w, 0, c.n0, q * a.n_kblocks + kb,
                      c.full_b + 8u * rb.stage);
          rb.advance(a.b_stages);
        }
      }
    }
    trace_ev(a, 0, ti, 1);
    b_loaded = true;
  }
}

// ================================================================================ MMA issuer
// Executed by the whole warp (uniform control flow, cheap uniform-datapath address math); only the
// tcgen05 instructions themselves are issued by lane 0.
template <int TAPS, int K16>
__device__ __forceinline__ void mma_role(const TcArgs& a, const Ctx& c, bool use_xf, int lane) {
  Ring ra, rb;
  int acc = 0;
  uint32_t acc_phase = 0;
  bool b_waited = false;
  const uint32_t a_stride = a.a_stage_bytes, b_stride = a.b_stage_bytes;
  const uint32_t row_units = (uint32_t)a.row_bytes >> 4;   // tap shift in 16-byte descriptor units
  const uint64_t hi64 = (uint64_t)a.desc_hi << 32;
  const uint32_t idesc = a.idesc;
  const bool resident = a.resident != 0;
  const int n_kblocks = a.n_kblocks, a_stages = a.a_stages, b_stages = a.b_stages;
  const uint32_t a_ready = use_xf ? c.ready_a : c.full_a;
  const bool issuer = lane == 0;   // tracing only
  int ti = 0;
  for (int mt = c.mt0; mt < a.total_m_tiles; mt += c.mt_stride, ti++) {
    mbar_wait(c.tmem_empty + 8u * acc, acc_phase ^ 1);
    tc_fence_after();
    if (issuer) trace_ev(a, 1, ti, 0);
    const uint32_t d_tmem = c.tmem_base + (uint32_t)(acc * a.bn * a.m_sub);
    // sub-tiles of the unit that hold rows of the clip at all (the last unit of a clip may be short: its empty
    // sub-tiles are neither multiplied nor stored)
    const int rows_left = a.p.rows - (mt % a.m_tiles) * BM * a.m_sub;
    const int m_sub = rows_left >= BM * a.m_sub ? a.m_sub : (rows_left + BM - 1) / BM;
    const uint32_t sub_units = a.a_sub_bytes >> 4;
    for (int kb = 0; kb < n_kblocks; kb++) {
      mbar_wait(a_ready + 8u * ra.stage, ra.phase);
      tc_fence_after();
      if (issuer && kb == 0) trace_ev(a, 1, ti, 1);
      const uint32_t a_lo0 = (((c.smA + ra.stage * a_stride) >> 4) & 0x3FFFu) | (1u << 16);
#pragma unroll
      for (int q = 0; q < TAPS; q++) {
        int bstage;
        if (resident) {
          bstage = kb * TAPS + q;
          if (!b_waited) {
            mbar_wait(c.full_b + 8u * bstage, 0);
            tc_fence_after();
          }
        } else {
          bstage = rb.stage;
          mbar_wait(c.full_b + 8u * bstage, rb.phase);
          tc_fence_after();
        }
        const uint32_t b_lo = (((c.smB + bstage * b_stride) >> 4) & 0x3FFFu) | (1u << 16);
        const uint32_t a_lo = a_lo0 + q * row_units;
        if (elect_one()) {
          for (int sub = 0; sub < m_sub; sub++) {
#pragma unroll
            for (int kk = 0; kk < K16; kk++) {
              // k16 step inside the swizzled row: +32 bytes = +2 descriptor address units
              umma_f16(d_tmem + (uint32_t)(sub * a.bn), hi64 | (a_lo + sub * sub_units + 2 * kk),
                       hi64 | (b_lo + 2 * kk), idesc, (kb | q | kk) != 0 ? 1u : 0u);
            }
          }
          if (!resident) umma_commit(c.empty_b + 8u * bstage);
          if (q == TAPS - 1) umma_commit(c.empty_a + 8u * ra.stage);
          if (q == TAPS - 1 && kb == n_kblocks - 1) umma_commit(c.tmem_full + 8u * acc);
        }
        __syncwarp();
        if (!resident) rb.advance(b_stages);
      }
      ra.advance(a_stages);
    }
    if (issuer) trace_ev(a, 1, ti, 3);
    b_waited = true;
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
}

// Weights resident in shared memory (C <= 128: the layers whose MMAs are only 32-64 clocks long, so that
// the ISSUE rate of this warp is what bounds the kernel): one wait for the whole weight slice before the
// first tile, then per K block a single elected straight-line burst of TAPS x MSUB x K16 MMAs whose
// descriptors differ by compile-time multiples of loop-invariant strides -- one barrier wait, one
// elect and two commits per K block instead of per tap.
template <int TAPS, int K16, int MSUB>
__device__ __forceinline__ void mma_role_resident(const TcArgs& a, const Ctx& c, bool use_xf, int lane) {
  Ring ra;
  int acc = 0;
  uint32_t acc_phase = 0;
  const uint32_t a_stride16 = a.a_stage_bytes >> 4, b_stride16 = a.b_stage_bytes >> 4;
  const uint32_t row_units = (uint32_t)a.row_bytes >> 4;   // tap shift in 16-byte descriptor units
  const uint32_t sub_units = a.a_sub_bytes >> 4;
  const uint32_t hi = a.desc_hi, idesc = a.idesc;
  const uint32_t a_lo_base = ((c.smA >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t b_lo_base = ((c.smB >> 4) & 0x3FFFu) | (1u << 16);
  const uint32_t bn = (uint32_t)a.bn;
  const int n_kblocks = a.n_kblocks, a_stages = a.a_stages;
  const uint32_t a_ready = use_xf ? c.ready_a : c.full_a;
  const bool issuer = lane == 0;   // tracing only
  // The weight taps of K block kb arrive right behind its first A stage (producer order A(0), B(0, *),
  // A(1), B(1, *), ...): wait for them K block by K block during the first tile only.  Waiting for the
  // whole slice up front would deadlock when a layer has more K blocks than A stages (the producer
  // cannot reach B(kb >= a_stages, *) before this warp has released an A stage).
  bool b_waited = false;
  int ti = 0;
  for (int mt = c.mt0; mt < a.total_m_tiles; mt += c.mt_stride, ti++) {
    mbar_wait(c.tmem_empty + 8u * acc, acc_phase ^ 1);
    tc_fence_after();
    if (issuer) trace_ev(a, 1, ti, 0);
    const uint32_t d_tmem = c.tmem_base + (uint32_t)acc * bn * MSUB;
    for (int kb = 0; kb < n_kblocks; kb++) {
      mbar_wait(a_ready + 8u * ra.stage, ra.phase);
      tc_fence_after();
      if (issuer && kb == 0) trace_ev(a, 1, ti, 1);
      if (!b_waited) {
#pragma unroll
        for (int q = 0; q < TAPS; q++) mbar_wait(c.full_b + 8u * (uint32_t)(kb * TAPS + q), 0);
        tc_fence_after();
      }
      const uint32_t a_lo0 = a_lo_base + (uint32_t)ra.stage * a_stride16;
      const uint32_t b_lo0 = b_lo_base + (uint32_t)(kb * TAPS) * b_stride16;
      const uint32_t first = kb != 0 ? 1u : 0u;
      if (elect_one()) {
#pragma unroll
        for (int q = 0; q < TAPS; q++) {
#pragma unroll
          for (int sub = 0; sub < MSUB; sub++) {
#pragma unroll
            for (int kk = 0; kk < K16; kk++) {
              // k16 step inside the swizzled row: +32 bytes = +2 descriptor address units
              umma_f16_lohi(d_tmem + (uint32_t)sub * bn, a_lo0 + q * row_units + sub * sub_units + 2 * kk,
                            b_lo0 + q * b_stride16 + 2 * kk, hi, idesc, (q | kk) != 0 ? 1u : first);
            }
          }
        }
        umma_commit(c.empty_a + 8u * ra.stage);
        if (kb == n_kblocks - 1) umma_commit(c.tmem_full + 8u * acc);
      }
      __syncwarp();
      ra.advance(a_stages);
    }
    if (issuer) trace_ev(a, 1, ti, 3);
    b_waited = true;
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
}

// ================================================================================ transform
__device__ __forceinline__ void transform_role(const TcArgs& a, const Ctx& c, int xt, int lane) {
  Ring ra;
  uint32_t slope, slope_lo;   // packed bf16x2 (hi, lo) halves of the slope
  split_slope(a.p.prelu_in, slope, slope_lo);
  const int nvec = (int)(a.a_tx_bytes >> 4);
  int ti = 0;
  for (int mt = c.mt0; mt < a.total_m_tiles; mt += c.mt_stride, ti++) {
    for (int kb = 0; kb < a.n_kblocks; kb++) {
      mbar_wait(c.full_a + 8u * ra.stage, ra.phase);
      if (xt == 0 && kb == 0) trace_ev(a, 2, ti, 0);
      for (int sub = 0; sub < a.m_sub; sub++) {
      const uint32_t base = c.smA + ra.stage * a.a_stage_bytes + sub * a.a_sub_bytes;
      int i = xt;
      for (; i + 3 * 128 < nvec; i += 4 * 128) {   // 4 independent vectors in flight
        uint4 v0 = lds_u4(base + (uint32_t)i * 16), v1 = lds_u4(base + (uint32_t)(i + 128) * 16);
        uint4 v2 = lds_u4(base + (uint32_t)(i + 256) * 16), v3 = lds_u4(base + (uint32_t)(i + 384) * 16);
        sts_u4(base + (uint32_t)i * 16, prelu_act8(v0, slope, slope_lo));
        sts_u4(base + (uint32_t)(i + 128) * 16, prelu_act8(v1, slope, slope_lo));
        sts_u4(base + (uint32_t)(i + 256) * 16, prelu_act8(v2, slope, slope_lo));
        sts_u4(base + (uint32_t)(i + 384) * 16, prelu_act8(v3, slope, slope_lo));
      }
      for (; i < nvec; i += 128) sts_u4(base + (uint32_t)i * 16, prelu_act8(lds_u4(base + (uint32_t)i * 16), slope, slope_lo));
      }
      fence_proxy_async();
      __syncwarp();
      if (lane == 0) mbar_arrive(c.ready_a + 8u * ra.stage);
      if (xt == 0 && kb == a.n_kblocks - 1) trace_ev(a, 2, ti, 2);
      ra.advance(a.a_stages);
    }
  }
}

// ================================================================================ epilogue
// NADD: number of residual inputs (0, 1: add1, 2: add1 + add2); NPRELU: output PReLUs (0, 1, 2).
// Eight warps (6-13): warp pair (w, w+4) shares TMEM lane quarter w%4.  Layers WITHOUT a fused input PReLU
// leave the four transform warps (2-5) idle: they join as a third epilogue warp per quarter (the
// epilogue, not the tensor pipe, bounds those layers at C <= 256).  A tile is cut into units of
// (sub-tile, 16 columns); the `nw` warps of a quarter take them round-robin (`widx` = 0..nw-1).
// FILM: per-clip gamma / beta present; without it the scale factors are per-launch scalars and only the
// folded bias vector c1 is read per column.
template <int NADD, int NPRELU, bool F32TM, bool FILM>
__device__ __forceinline__ void epilogue_role(const TcArgs& a, const Ctx& c, int warp, int lane, int widx,
                                              int nw) {
  const ou_conv_params& p = a.p;
  const int quarter = warp & 3;
  const int row = quarter * 32 + lane;
  // ordinal among the epilogue threads (coefficient set-up): main warps 0..255, helper warps 256..383
  const int et = warp >= EPI_WARP0 ? (int)threadIdx.x - EPI_WARP0 * 32
                                   : N_EPI_WARPS * 32 + (int)threadIdx.x - XF_WARP0 * 32;
  const int n_epi_threads = nw * 4 * 32;
  const int bn = a.bn;
  const int cout = p.cout, up = p.up, t_out = p.t_out, n_total = p.n;
  const int cbo = cl_cb(cout);
  const int cbo_shift = cbo == 64 ? 6 : (cbo == 32 ? 5 : 4);
  const size_t blk_stride = (size_t)t_out * cbo;
  const int n0 = c.n0;
  const act_t* add1 = (const act_t*)p.add1;
  const act_t* add2 = (const act_t*)p.add2;
  act_t* outp = (act_t*)p.out;
  const float s1 = p.scale1, s2 = p.scale2;
  const float slope1 = p.prelu_out, slope2 = p.prelu_out2;
  const bool has_film = FILM;
  const float c0s = s1 * s2;     // column scale without FiLM

  // column part of the output addressing, once per kernel (the first coefficient barrier publishes it):
  // columns [n0 + 16 g, +16) = channels co.. of depth-to-space phase ph -> channel block, position in it
  if (et < (bn >> 4)) {
    const int n = n0 + 16 * et;
    int off = -1, ph = 0;
    if (n < n_total && !F32TM) {
      ph = n / cout;
      const int co = n - ph * cout;
      off = (int)((size_t)(co >> cbo_shift) * blk_stride + (size_t)ph * cbo + (co & (cbo - 1)));
    }
    asm volatile("st.shared.v2.s32 [%0], {%1, %2};" ::"r"(c.coltab + 8u * et), "r"(off), "r"(ph) : "memory");
  }

  int acc = 0;
  uint32_t acc_phase = 0;
  int ti = 0;
  for (int mt = c.mt0; mt < a.total_m_tiles; mt += c.mt_stride, ti++) {
    const int b = mt / a.m_tiles;
    const int m0 = (mt - b * a.m_tiles) * BM * a.m_sub;
    // per-column coefficients: y = c0*(acc + add1) + c2*add2 + c1   (see header comment)
    const uint32_t coef = c.coef + (uint32_t)(acc * 3 * bn) * 4u;
    const float* coefp = c.coef_ptr + acc * 3 * bn;
    if (ti < 2 || has_film) {
      for (int i = et; i < bn; i += n_epi_threads) {
        const int n = n0 + i;
        float g = 1.f, be = 0.f, bias = 0.f;
        if (n < n_total) {
          const int co = n % cout;
          if (has_film) {
            g = p.gamma[(size_t)b * p.film_bstride + co];
            be = p.beta[(size_t)b * p.film_bstride + co];
          }
          if (p.bias) bias = p.bias[n];
        }
        const float c0 = F32TM ? 1.f : g * s1 * s2;
        sts_f1(coef + 4u * i, c0);
        sts_f1(coef + 4u * (bn + i), F32TM ? bias : fmaf(c0, bias, be));
        sts_f1(coef + 4u * (2 * bn + i), g * s2);
      }
      epi_bar_sync(n_epi_threads);   // coefficients visible to all epilogue warps
    }
    if (row == 0 && widx == 0) trace_ev(a, 3, ti, 0);

    const size_t clip_base = (size_t)b * cout * t_out;
    const int m_sub = a.m_sub;
    // Work units of this thread: (sub-tile, 16 columns) = one 32-byte output vector (16 consecutive
    // channels of one time step); unit u = widx + k * nw is this warp's k-th.  The column part of the
    // output address (channel block, depth-to-space phase) comes from the per-kernel table `coltab`
    // (int2 per 16-column group: element offset inside the clip or -1, phase); the row part is
    // (m0 + sub * 128 + row) * up * cbo.  Residual vectors are prefetched D units ahead (the first D while
    // the MMAs of this tile still run) so that their HBM latency never sits on the critical path.
    constexpr int D = NADD == 1 ? 4 : (NADD == 2 ? 2 : 1);
    const int unit_shift = bn == 256 ? 4 : (bn == 128 ? 3 : (bn == 64 ? 2 : 1));   // log2(bn / 16)
    const int n_units = m_sub << unit_shift;
    const int nk = widx < n_units ? (n_units - widx + nw - 1) / nw : 0;
    const int row_stride = up * cbo;                      // elements per GEMM row in the output layout
    const long row_off0 = (long)(m0 + row) * row_stride;  // sub-tile 0
    // element offset of unit u's 16-channel vector inside the output tensor, or -1
    auto unit_offset = [&](int u) -> long {
      const int sub = u >> unit_shift, g = u - (sub << unit_shift);
      const int2 ct = *reinterpret_cast<const int2*>(c.coltab_ptr + 2 * g);
      const int jj = m0 + sub * BM + row;
      if (ct.x < 0 || jj >= p.rows || jj * up + ct.y >= t_out) return -1;
      return (long)clip_base + row_off0 + (long)sub * (BM * row_stride) + ct.x;
    };
    U8 pre1[D], pre2[D];
    long poff[D];
    auto prefetch = [&](int d, int k) {
      poff[d] = k < nk ? unit_offset(widx + k * nw) : -1;
      if (NADD > 0 && poff[d] >= 0) {
        pre1[d] = ldg_nc_v8(add1 + poff[d]);
        if (NADD > 1) pre2[d] = ldg_nc_v8(add2 + poff[d]);
      }
    };
#pragma unroll
    for (int d = 0; d < D; d++) prefetch(d, d);

    mbar_wait(c.tmem_full + 8u * acc, acc_phase);
    tc_fence_after();
    if (row == 0 && widx == 0) trace_ev(a, 3, ti, 1);
    if (nk == 0) {   // more warps than units (N = 32, one sub-tile): nothing to read, release at once
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(c.tmem_empty + 8u * acc);
    }
    const uint32_t taddr0 = c.tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * m_sub * bn);
    for (int base = 0; base < nk; base += D) {
#pragma unroll
      for (int d = 0; d < D; d++) {
        const int k = base + d;
        if (k >= nk) break;
        const int u = widx + k * nw;
        const int sub = u >> unit_shift;
        const int col = (u - (sub << unit_shift)) << 4;
        uint32_t r[16];
        tmem_ld16(taddr0 + (uint32_t)(sub * bn + col), r);
        tmem_ld_wait();
        if (k == nk - 1) {
          // accumulators fully read by this warp: hand the TMEM buffer back before the store phase
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(c.tmem_empty + 8u * acc);
          if (row == 0 && widx == 0) trace_ev(a, 3, ti, 2);
        }
        if (F32TM) {
          const int j = m0 + sub * BM + row;
          if (j < p.rows && n0 + col < n_total) {
            // blocked fp32 [B][N/16][rows][16]: the 16 columns of a row are one 64-byte line, consecutive rows
            // (lanes) consecutive lines -- two full-sector 32-byte stores per lane
            float* dst = p.out_f32_blk + f32blk_off(b, n0 + col, j, n_total, p.rows);
#pragma unroll
            for (int h = 0; h < 2; h++) {
              U8 o;
#pragma unroll
              for (int i = 0; i < 2; i++) {
                const float4 x1 = lds_f4(coef + 4u * (bn + col + 8 * h + 4 * i));
                o.w[4 * i] = __float_as_uint(__uint_as_float(r[8 * h + 4 * i]) + x1.x);
                o.w[4 * i + 1] = __float_as_uint(__uint_as_float(r[8 * h + 4 * i + 1]) + x1.y);
                o.w[4 * i + 2] = __float_as_uint(__uint_as_float(r[8 * h + 4 * i + 2]) + x1.z);
                o.w[4 * i + 3] = __float_as_uint(__uint_as_float(r[8 * h + 4 * i + 3]) + x1.w);
              }
              stg_v8(dst + 8 * h, o);
            }
          }
          continue;
        }
        const long off = poff[d];
        if (off >= 0) {
          // coefficient vectors as plain shared-memory loads: the compiler is free to hoist them over the
          // arithmetic of the previous group (volatile asm accessors would serialise every load).
          // Arithmetic on packed fp32 pairs (FFMA2 / FADD2 / FMUL2).
          const float4* k0 = reinterpret_cast<const float4*>(coefp + col);
          const float4* k1 = reinterpret_cast<const float4*>(coefp + bn + col);
          const float4* k2 = reinterpret_cast<const float4*>(coefp + 2 * bn + col);
          U8 o;
#pragma unroll
          for (int i = 0; i < 4; i++) {
            const float4 x1 = k1[i];
            float4 x0 = make_float4(c0s, c0s, c0s, c0s);
            if (FILM) x0 = k0[i];
            float2 a01 = u2_as_f2(r[4 * i], r[4 * i + 1]), a23 = u2_as_f2(r[4 * i + 2], r[4 * i + 3]);
            if (NADD > 0) {
              a01 = fadd2(a01, act2_to_f2(pre1[d].w[2 * i]));
              a23 = fadd2(a23, act2_to_f2(pre1[d].w[2 * i + 1]));
            }
            a01 = ffma2(make_float2(x0.x, x0.y), a01, make_float2(x1.x, x1.y));
            a23 = ffma2(make_float2(x0.z, x0.w), a23, make_float2(x1.z, x1.w));
            if (NADD > 1) {
              float4 x2 = make_float4(s2, s2, s2, s2);
              if (FILM) x2 = k2[i];
              a01 = ffma2(make_float2(x2.x, x2.y), act2_to_f2(pre2[d].w[2 * i]), a01);
              a23 = ffma2(make_float2(x2.z, x2.w), act2_to_f2(pre2[d].w[2 * i + 1]), a23);
            }
            if (NPRELU > 0) a01 = prelu2<false>(a01, slope1), a23 = prelu2<false>(a23, slope1);
            if (NPRELU > 1) a01 = prelu2<false>(a01, slope2), a23 = prelu2<false>(a23, slope2);
            o.w[2 * i] = f2_to_act2(a01.x, a01.y);
            o.w[2 * i + 1] = f2_to_act2(a23.x, a23.y);
          }
          stg_v8(outp + off, o);
        }
        // the slot's residual registers are free again: fetch the vectors of unit k + D into them
        prefetch(d, k + D);
      }
    }
    if (row == 0 && widx == 0) trace_ev(a, 3, ti, 3);
    acc ^= 1;
    if (acc == 0) acc_phase ^= 1;
  }
}

// ================================================================================ kernel
__global__ void __launch_bounds__(NTHREADS, 1)   // 14 warps: the register file gives 128 per thread
conv1d_tc_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tm_a,
                 const __grid_constant__ CUtensorMap tm_w) {
  extern __shared__ uint8_t smem_raw[];
  const ou_conv_params& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // shared memory carve-up in the 32-bit shared window (stages 1024-byte aligned for the swizzle)
  Ctx c;
  const uint32_t raw = smem_u32(smem_raw);
  c.smA = (raw + 1023u) & ~1023u;
  c.smB = c.smA + (uint32_t)a.a_stages * a.a_stage_bytes;
  uint32_t tail = c.smB + (uint32_t)a.b_stages * a.b_stage_bytes;
  c.full_a = tail;
  c.ready_a = c.full_a + 8u * MAX_A_STAGES;
  c.empty_a = c.ready_a + 8u * MAX_A_STAGES;
  c.full_b = c.empty_a + 8u * MAX_A_STAGES;
  c.empty_b = c.full_b + 8u * MAX_B_STAGES;
  c.tmem_full = c.empty_b + 8u * MAX_B_STAGES;
  c.tmem_empty = c.tmem_full + 16u;
  const uint32_t tmem_slot = c.tmem_empty + 16u;
  c.coef = tmem_slot + 16u;
  c.coef_ptr = reinterpret_cast<const float*>(smem_raw + (c.coef - raw));
  c.coltab = c.coef + (uint32_t)(6 * a.bn) * 4u;
  c.coltab_ptr = reinterpret_cast<const int*>(smem_raw + (c.coltab - raw));

  const bool use_xf = p.has_prelu_in != 0;

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.a_stages; i++) {
      mbar_init(c.full_a + 8u * i, 1);
      mbar_init(c.ready_a + 8u * i, 4);
      mbar_init(c.empty_a + 8u * i, 1);
    }
    for (int i = 0; i < a.b_stages; i++) {
      mbar_init(c.full_b + 8u * i, 1);
      mbar_init(c.empty_b + 8u * i, 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(c.tmem_full + 8u * i, 1);
      mbar_init(c.tmem_empty + 8u * i, use_xf ? N_EPI_WARPS : N_EPI_WARPS + 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(tmem_slot, a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(c.tmem_base) : "r"(tmem_slot));

  // static tile schedule: this CTA owns N tile `nt` and M tiles mt0, mt0 + stride, ...
  c.nt = blockIdx.x % a.n_ntiles;
  c.mt0 = blockIdx.x / a.n_ntiles;
  c.mt_stride = a.ctas_per_ntile;
  c.n0 = c.nt * a.bn;

  if (warp == 0) {
    if (lane == 0) producer_role(a, c, &tm_a, &tm_w);
  } else if (warp == 1) {
    const int k16 = a.cb / 16;
    if (a.resident) {
      switch ((p.taps * 8 + k16) * 4 + (a.m_sub == 4 ? 2 : a.m_sub - 1)) {
        case (1 * 8 + 1) * 4 + 0: mma_role_resident<1, 1, 1>(a, c, use_xf, lane); break;
        case (1 * 8 + 2) * 4 + 0: mma_role_resident<1, 2, 1>(a, c, use_xf, lane); break;
        case (1 * 8 + 4) * 4 + 0: mma_role_resident<1, 4, 1>(a, c, use_xf, lane); break;
        case (3 * 8 + 1) * 4 + 0: mma_role_resident<3, 1, 1>(a, c, use_xf, lane); break;
        case (3 * 8 + 2) * 4 + 0: mma_role_resident<3, 2, 1>(a, c, use_xf, lane); break;
        case (3 * 8 + 4) * 4 + 0: mma_role_resident<3, 4, 1>(a, c, use_xf, lane); break;
        case (5 * 8 + 1) * 4 + 0: mma_role_resident<5, 1, 1>(a, c, use_xf, lane); break;
        case (5 * 8 + 2) * 4 + 0: mma_role_resident<5, 2, 1>(a, c, use_xf, lane); break;
        case (5 * 8 + 4) * 4 + 0: mma_role_resident<5, 4, 1>(a, c, use_xf, lane); break;
        case (1 * 8 + 1) * 4 + 1: mma_role_resident<1, 1, 2>(a, c, use_xf, lane); break;
        case (1 * 8 + 2) * 4 + 1: mma_role_resident<1, 2, 2>(a, c, use_xf, lane); break;
        case (1 * 8 + 4) * 4 + 1: mma_role_resident<1, 4, 2>(a, c, use_xf, lane); break;
        case (3 * 8 + 1) * 4 + 1: mma_role_resident<3, 1, 2>(a, c, use_xf, lane); break;
        case (3 * 8 + 2) * 4 + 1: mma_role_resident<3, 2, 2>(a, c, use_xf, lane); break;
        case (3 * 8 + 4) * 4 + 1: mma_role_resident<3, 4, 2>(a, c, use_xf, lane); break;
        case (5 * 8 + 1) * 4 + 1: mma_role_resident<5, 1, 2>(a, c, use_xf, lane); break;
        case (5 * 8 + 2) * 4 + 1: mma_role_resident<5, 2, 2>(a, c, use_xf, lane); break;
        case (5 * 8 + 4) * 4 + 1: mma_role_resident<5, 4, 2>(a, c, use_xf, lane); break;
        case (1 * 8 + 1) * 4 + 2: mma_role_resident<1, 1, 4>(a, c, use_xf, lane); break;
        case (1 * 8 + 2) * 4 + 2: mma_role_resident<1, 2, 4>(a, c, use_xf, lane); break;
        case (1 * 8 + 4) * 4 + 2: mma_role_resident<1, 4, 4>(a, c, use_xf, lane); break;
        case (3 * 8 + 1) * 4 + 2: mma_role_resident<3, 1, 4>(a, c, use_xf, lane); break;
        case (3 * 8 + 2) * 4 + 2: mma_role_resident<3, 2, 4>(a, c, use_xf, lane); break;
        case (3 * 8 + 4) * 4 + 2: mma_role_resident<3, 4, 4>(a, c, use_xf, lane); break;
        case (5 * 8 + 1) * 4 + 2: mma_role_resident<5, 1, 4>(a, c, use_xf, lane); break;
        case (5 * 8 + 2) * 4 + 2: mma_role_resident<5, 2, 4>(a, c, use_xf, lane); break;
        case (5 * 8 + 4) * 4 + 2: mma_role_resident<5, 4, 4>(a, c, use_xf, lane); break;
        default: break;   // rejected on the host
      }
    } else
    switch (p.taps * 8 + k16) {
      case 1 * 8 + 1: mma_role<1, 1>(a, c, use_xf, lane); break;
      case 1 * 8 + 2: mma_role<1, 2>(a, c, use_xf, lane); break;
      case 1 * 8 + 4: mma_role<1, 4>(a, c, use_xf, lane); break;
      case 3 * 8 + 1: mma_role<3, 1>(a, c, use_xf, lane); break;
      case 3 * 8 + 2: mma_role<3, 2>(a, c, use_xf, lane); break;
      case 3 * 8 + 4: mma_role<3, 4>(a, c, use_xf, lane); break;
      case 5 * 8 + 1: mma_role<5, 1>(a, c, use_xf, lane); break;
      case 5 * 8 + 2: mma_role<5, 2>(a, c, use_xf, lane); break;
      case 5 * 8 + 4: mma_role<5, 4>(a, c, use_xf, lane); break;
      default: break;   // rejected on the host
    }
  } else if (warp < EPI_WARP0 && use_xf) {
    transform_role(a, c, threadIdx.x - XF_WARP0 * 32, lane);
  } else {
    // epilogue: warps 6-13 always; warps 2-5 as a third warp per TMEM lane quarter when they have no
    // input PReLU to apply
    const int nw = use_xf ? N_EPI_WARPS / 4 : N_EPI_WARPS / 4 + 1;
    const int widx = warp >= EPI_WARP0 ? (warp - EPI_WARP0) >> 2 : N_EPI_WARPS / 4;
    const int nadd = p.add2 ? 2 : (p.add1 ? 1 : 0);
    const int nprelu = p.has_prelu_out2 ? 2 : (p.has_prelu_out ? 1 : 0);
    if (p.out_f32_blk) {
      epilogue_role<0, 0, true, false>(a, c, warp, lane, widx, nw);
    } else if (p.gamma != nullptr) {
      switch (nadd * 3 + nprelu) {
        case 0: epilogue_role<0, 0, false, true>(a, c, warp, lane, widx, nw); break;
        case 1: epilogue_role<0, 1, false, true>(a, c, warp, lane, widx, nw); break;
        case 2: epilogue_role<0, 2, false, true>(a, c, warp, lane, widx, nw); break;
        case 3: epilogue_role<1, 0, false, true>(a, c, warp, lane, widx, nw); break;
        case 4: epilogue_role<1, 1, false, true>(a, c, warp, lane, widx, nw); break;
        case 5: epilogue_role<1, 2, false, true>(a, c, warp, lane, widx, nw); break;
        case 6: epilogue_role<2, 0, false, true>(a, c, warp, lane, widx, nw); break;
        case 7: epilogue_role<2, 1, false, true>(a, c, warp, lane, widx, nw); break;
        default: epilogue_role<2, 2, false, true>(a, c, warp, lane, widx, nw); break;
      }
    } else {
      switch (nadd * 3 + nprelu) {
        case 0: epilogue_role<0, 0, false, false>(a, c, warp, lane, widx, nw); break;
        case 1: epilogue_role<0, 1, false, false>(a, c, warp, lane, widx, nw); break;
        case 2: epilogue_role<0, 2, false, false>(a, c, warp, lane, widx, nw); break;
        case 3: epilogue_role<1, 0, false, false>(a, c, warp, lane, widx, nw); break;
        case 4: epilogue_role<1, 1, false, false>(a, c, warp, lane, widx, nw); break;
        case 5: epilogue_role<1, 2, false, false>(a, c, warp, lane, widx, nw); break;
        case 6: epilogue_role<2, 0, false, false>(a, c, warp, lane, widx, nw); break;
        case 7: epilogue_role<2, 1, false, false>(a, c, warp, lane, widx, nw); break;
        default: epilogue_role<2, 2, false, false>(a, c, warp, lane, widx, nw); break;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(c.tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side
long long* g_trace = nullptr;

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn g_encode = nullptr;

static int get_encode() {
  if (g_encode) return OU_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
  if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || fn == nullptr) {
    set_error("ou_conv1d(tc): cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
    return OU_ERR_CUDA;
  }
  g_encode = (EncodeTiledFn)fn;
  return OU_OK;
}

int plan(const ou_conv_params* p, TcArgs* a) {
  // strided input (down convs, st_convs) is read through a (CB, s, T/s, ...) view of the tensor:
  // needs T divisible by s (always true inside enhance(); odd module-level lengths fall back)
  if (p->w_tc == nullptr || p->t_in % p->s != 0) return OU_ERR_UNSUPPORTED;
  if (p->taps != 1 && p->taps != 3 && p->taps != 5) return OU_ERR_UNSUPPORTED;
  int bn = 0;
  static const int bn_max = [] { const char* e = getenv("OU_TC_BN_MAX"); return e ? atoi(e) : 256; }();
  for (int cand : {256, 128, 64, 32})
    if (cand <= bn_max && p->npad % cand == 0) {
      bn = cand;
      break;
    }
  if (!bn || p->cin % 16 || p->cout % 16) return OU_ERR_UNSUPPORTED;
  if (p->add2 && !p->add1) return OU_ERR_UNSUPPORTED;
  const int g_num_sms = p->max_ctas > 0 && p->max_ctas < num_sms() ? p->max_ctas : num_sms();
  a->p = *p;
  a->cb = cl_cb(p->cin);
  a->row_bytes = a->cb * 2;
  a->cin_blocks = p->cin / a->cb;
  a->n_kblocks = p->s * a->cin_blocks;
  a->bn = bn;
  a->n_ntiles = p->npad / bn;
  a->arows = BM + p->taps - 1;
  a->a_tx_bytes = (uint32_t)(a->arows * a->row_bytes);
  a->b_tx_bytes = (uint32_t)(bn * a->row_bytes);
  a->a_sub_bytes = (a->a_tx_bytes + 1023u) & ~1023u;
  a->b_stage_bytes = (a->b_tx_bytes + 1023u) & ~1023u;
  const int budget = 232448 - 2048 - 6 * bn * 4 - 128;   // 227 KB minus alignment slack, barriers, coefficients, column table
  const int nb_all = p->taps * a->n_kblocks;
  // Sub-tiles of 128 rows per scheduling unit.  A unit costs a fixed round of barrier hand-overs
  // between the five warp roles (~2 000 clk measured) whatever its size, which dominates layers whose
  // 128-row tile is only a handful of 32-64 clk MMAs: take the largest of 4 / 2 / 1 that fits the
  // 512 TMEM columns (two accumulator buffers) and the shared-memory budget and leaves every SM a
  // few units (measured at C = 128, k5: streamed weights with 2 sub-tiles beat resident weights
  // with 1).  OU_TC_MSUB caps it (A/B runs).
  static const int msub_cap = [] { const char* e = getenv("OU_TC_MSUB"); return e ? atoi(e) : 4; }();
  a->m_sub = 1;
  for (int ms : {4, 2}) {
    if (ms > msub_cap || 2 * bn * ms > 512 || p->rows <= ms * BM) continue;
    // three A stages next to the resident weights, or next to a ring of at least two weight stages
    if ((budget - 3 * (int)a->a_sub_bytes * ms) / (int)a->b_stage_bytes < 2) continue;
    if ((long)ceil_div(p->rows, BM * ms) * p->batch < 4L * g_num_sms) continue;
    a->m_sub = ms;
    break;
  }
  a->m_tiles = ceil_div(p->rows, BM * a->m_sub);
  a->total_m_tiles = a->m_tiles * p->batch;
  a->a_stage_bytes = a->a_sub_bytes * a->m_sub;
  if (nb_all <= MAX_B_STAGES && nb_all * (int)a->b_stage_bytes + 3 * (int)a->a_stage_bytes <= budget) {
    a->resident = 1;
    a->b_stages = nb_all;
  } else {
    a->resident = 0;
    int bs = (budget - 3 * (int)a->a_stage_bytes) / (int)a->b_stage_bytes;
    if (bs < 2) return OU_ERR_UNSUPPORTED;
    a->b_stages = bs > 6 ? 6 : bs;
  }
  int as = (budget - a->b_stages * (int)a->b_stage_bytes) / (int)a->a_stage_bytes;
  if (as < 2) return OU_ERR_UNSUPPORTED;
  a->a_stages = as > MAX_A_STAGES ? MAX_A_STAGES : as;
  // instruction descriptor: D=f32, A=B=bf16, K-major both, N = bn, M = 128
  a->idesc = (1u << 4) | ((uint32_t)OU_ACT_IS_BF16 << 7) | ((uint32_t)OU_ACT_IS_BF16 << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * bn * a->m_sub)) cols <<= 1;
  a->tmem_cols = cols;
  const uint32_t sbo = 8u * a->row_bytes;                       // 8 rows
  const uint32_t layout = a->row_bytes == 128 ? 2u : (a->row_bytes == 64 ? 4u : 6u);
  a->desc_hi = ((sbo >> 4) & 0x3FFFu) | (1u << (46 - 32)) | (layout << (61 - 32));
  return OU_OK;
}

static CUtensorMapSwizzle swizzle_for(int row_bytes) {
  return row_bytes == 128 ? CU_TENSOR_MAP_SWIZZLE_128B
                          : (row_bytes == 64 ? CU_TENSOR_MAP_SWIZZLE_64B : CU_TENSOR_MAP_SWIZZLE_32B);
}

int launch(const ou_conv_params* p, cudaStream_t st) {
  TcArgs a;
  int rc = plan(p, &a);
  if (rc) return rc;
  rc = get_encode();
  if (rc) return rc;
  CUtensorMap tm_a, tm_w;
  {
    // activations: [B][C/CB][T][CB] bf16 viewed as (CB, s, T/s, C/CB, B): time t = j*s + r
    const cuuint64_t rowb = (cuuint64_t)a.row_bytes;
    cuuint64_t dims[5] = {(cuuint64_t)a.cb, (cuuint64_t)p->s, (cuuint64_t)(p->t_in / p->s),
                          (cuuint64_t)a.cin_blocks, (cuuint64_t)p->batch};
    cuuint64_t strides[4] = {rowb, rowb * p->s, rowb * p->t_in, rowb * p->t_in * a.cin_blocks};
    cuuint32_t box[5] = {(cuuint32_t)a.cb, 1, (cuuint32_t)a.arows, 1, 1};
    cuuint32_t estr[5] = {1, 1, 1, 1, 1};
    CUresult r = g_encode(&tm_a, OU_TMA_ACT, 5, const_cast<void*>(p->x), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(a.row_bytes),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("ou_conv1d(tc): cuTensorMapEncodeTiled(A) failed with %d", (int)r);
      return OU_ERR_CUDA;
    }
  }
  {
    // weights: [taps * cin/CB][npad][CB] bf16 -> dims (CB, npad, taps * cin/CB)
    cuuint64_t dims[3] = {(cuuint64_t)a.cb, (cuuint64_t)p->npad, (cuuint64_t)(p->taps * a.n_kblocks)};
    cuuint64_t strides[2] = {(cuuint64_t)a.row_bytes, (cuuint64_t)p->npad * a.row_bytes};
    cuuint32_t box[3] = {(cuuint32_t)a.cb, (cuuint32_t)a.bn, 1};
    cuuint32_t estr[3] = {1, 1, 1};
    CUresult r = g_encode(&tm_w, OU_TMA_ACT, 3, const_cast<void*>(p->w_tc), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(a.row_bytes),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("ou_conv1d(tc): cuTensorMapEncodeTiled(W) failed with %d", (int)r);
      return OU_ERR_CUDA;
    }
  }
  const size_t smem = 1024 + (size_t)a.a_stages * a.a_stage_bytes + (size_t)a.b_stages * a.b_stage_bytes +
                      (3 * MAX_A_STAGES + 2 * MAX_B_STAGES + 4) * sizeof(uint64_t) + 16 +
                      (size_t)6 * a.bn * sizeof(float) + 128;
  static SmemConfig cfg;
  if ((rc = ensure_smem(conv1d_tc_kernel, smem, cfg, "ou_conv1d(tc)"))) return rc;
  int per = (p->max_ctas > 0 && p->max_ctas < num_sms() ? p->max_ctas : num_sms()) / a.n_ntiles;
  if (per < 1) per = 1;
  if (per > a.total_m_tiles) per = a.total_m_tiles;
  a.ctas_per_ntile = per;
  a.trace = g_trace;
  dim3 grid(per * a.n_ntiles);
  conv1d_tc_kernel<<<grid, NTHREADS, smem, st>>>(a, tm_a, tm_w);
  return check_launch("ou_conv1d(tc)");
}

}  // namespace tc
}  // namespace ou

extern "C" int ou_debug_set_trace(void* device_buffer) {
  ou::tc::g_trace = (long long*)device_buffer;
  return OU_OK;
}
