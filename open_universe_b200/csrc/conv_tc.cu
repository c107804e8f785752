// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d for sm_100a -- the production path of
// ou_conv1d for stride-1 input geometry (conv1/conv2/conv3, transposed up convs, 1x1s, GRU
// input projections: > 80 % of the FLOPs of a score step).  Contract: include/ou_b200.h.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      producer   : cp.async.bulk (TMA bulk engine, UBLKCP) global -> smem rings,
//                            completion on mbarriers (complete_tx)
//   warp 1      MMA issuer : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                            (bf16 x bf16 -> fp32 in TMEM), tcgen05.commit frees smem stages and
//                            publishes accumulators; also owns the TMEM allocation
//   warps 2-5   transform  : in-smem prologue on each landed A stage: zero-fill of rows outside the
//                            sequence ("same" padding / tile overrun) and the fused input PReLU;
//                            fence.proxy.async, then hand the stage to the MMA warp
//   warps 6-9   epilogue   : tcgen05.ld accumulator rows -> registers, bias / adds / FiLM / PReLUs,
//                            16-byte bf16 stores in the blocked layout (or fp32 time-major)
// TMEM holds two accumulator buffers (2 x BN columns) so the epilogue of tile i overlaps the MMAs
// of tile i+1.  Operands use the K-major NO-SWIZZLE canonical layout: 8-row x 16-byte core
// matrices, which is exactly a run of consecutive time steps of one 8-channel group in the blocked
// activation layout -- so a conv tap is a +16-byte shift of the A descriptor's start address and
// one staged A tile (with taps-1 halo rows) feeds every tap.
// Weights of a CTA's N-slice stay resident in shared memory across all its M tiles when they fit
// (C <= 128); otherwise they stream through a ring from L2.
#include "common.cuh"

namespace ou {
namespace tc {

constexpr int BM = 128;
constexpr int NTHREADS = 320;
constexpr int XF_WARP0 = 2, EPI_WARP0 = 6;
constexpr int MAX_STAGES = 16;

struct TcArgs {
  ou_conv_params p;
  int cin_chunks;      // cin / 8
  int kb_chunks;       // 16-byte channel chunks per K block (KB / 8)
  int n_kblocks;       // kpad / KB
  int arows;           // BM + taps - 1
  int bn;              // N per CTA tile
  int n_ntiles;        // npad / bn
  int m_tiles;         // ceil(rows / BM) per clip
  int total_m_tiles;   // m_tiles * batch
  int ctas_per_ntile;  // gridDim.x / n_ntiles
  int a_stages, b_stages;
  int resident;        // whole weight slice stays in smem (b_stages == n_kblocks)
  uint32_t a_stage_bytes, b_stage_bytes;
  uint32_t idesc;
  uint32_t tmem_cols;
  long long* trace;    // debug: [role 0..3][tile 0..63][event 0..3] clock64 stamps of CTA 0 (or NULL)
};

__device__ __forceinline__ void trace_ev(const TcArgs& a, int role, int tile_i, int ev) {
  if (a.trace != nullptr && blockIdx.x == 0 && tile_i < 64)
    a.trace[(role * 64 + tile_i) * 4 + ev] = clock64();
}

// ------------------------------------------------------------------------------------ PTX helpers
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra DONE;\n"
      "bra WAIT_LOOP;\n"
      "DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tmem_alloc(uint32_t smem_dst, uint32_t ncols) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_dst),
               "r"(ncols)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols)
               : "memory");
}
__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc,
                                         uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,"
      "%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
        "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]),
        "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]),
        "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]),
        "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// K-major, no-swizzle shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
//   [0,14) start>>4 | [16,30) LBO>>4 (stride between the two 8-element K chunks of one MMA)
//   [32,46) SBO>>4 (stride between 8-row groups) | [46,48) version=1 | [61,64) layout=0
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
         ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | (1ull << 46);
}

struct Ring {
  int stage = 0;
  uint32_t phase = 0;
  __device__ __forceinline__ void advance(int n) {
    if (++stage == n) {
      stage = 0;
      phase ^= 1;
    }
  }
};

// Epilogue for 8 consecutive GEMM columns [n, n+8) of one row (one 16-byte output vector).
__device__ __forceinline__ void epilogue_vec8(const ou_conv_params& p, int b, int j, int n,
                                              const uint32_t* acc) {
  float v[8];
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] = __uint_as_float(acc[i]);
  if (p.bias) {
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(p.bias + n));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(p.bias + n + 4));
    v[0] += b0.x, v[1] += b0.y, v[2] += b0.z, v[3] += b0.w;
    v[4] += b1.x, v[5] += b1.y, v[6] += b1.z, v[7] += b1.w;
  }
  if (p.out_f32_tm) {
    float4* dst = reinterpret_cast<float4*>(p.out_f32_tm + ((size_t)b * p.rows + j) * p.n + n);
    dst[0] = make_float4(v[0], v[1], v[2], v[3]);
    dst[1] = make_float4(v[4], v[5], v[6], v[7]);
    return;
  }
  const int ph = n / p.cout;
  const int co = n - ph * p.cout;
  const int t = j * p.up + ph;
  if (t >= p.t_out) return;
  const size_t off = (((size_t)b * (p.cout >> 3) + (co >> 3)) * p.t_out + t) * 8;
  if (p.add1) {
    const uint4 a = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)p.add1 + off);
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 f = bf2_to_f2(pa[i]);
      v[2 * i] += f.x, v[2 * i + 1] += f.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] *= p.scale1;
  if (p.add2) {
    const uint4 a = *reinterpret_cast<const uint4*>((const __nv_bfloat16*)p.add2 + off);
    const uint32_t* pa = reinterpret_cast<const uint32_t*>(&a);
#pragma unroll
    for (int i = 0; i < 4; i++) {
      const float2 f = bf2_to_f2(pa[i]);
      v[2 * i] += f.x, v[2 * i + 1] += f.y;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; i++) v[i] *= p.scale2;
  if (p.gamma) {
    const float* g = p.gamma + (size_t)b * p.film_bstride + co;
    const float* be = p.beta + (size_t)b * p.film_bstride + co;
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(g)), g1 = __ldg(reinterpret_cast<const float4*>(g + 4));
    const float4 e0 = __ldg(reinterpret_cast<const float4*>(be)), e1 = __ldg(reinterpret_cast<const float4*>(be + 4));
    v[0] = fmaf(g0.x, v[0], e0.x), v[1] = fmaf(g0.y, v[1], e0.y), v[2] = fmaf(g0.z, v[2], e0.z);
    v[3] = fmaf(g0.w, v[3], e0.w), v[4] = fmaf(g1.x, v[4], e1.x), v[5] = fmaf(g1.y, v[5], e1.y);
    v[6] = fmaf(g1.z, v[6], e1.z), v[7] = fmaf(g1.w, v[7], e1.w);
  }
  if (p.has_prelu_out) {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = prelu_f(v[i], p.prelu_out);
  }
  if (p.has_prelu_out2) {
#pragma unroll
    for (int i = 0; i < 8; i++) v[i] = prelu_f(v[i], p.prelu_out2);
  }
  const uint4 o = make_uint4(f2_to_bf2(v[0], v[1]), f2_to_bf2(v[2], v[3]), f2_to_bf2(v[4], v[5]),
                             f2_to_bf2(v[6], v[7]));
  *reinterpret_cast<uint4*>((__nv_bfloat16*)p.out + off) = o;
}

__global__ void __launch_bounds__(NTHREADS, 1) conv1d_tc_kernel(const TcArgs a) {
  extern __shared__ __align__(128) uint8_t smem[];
  const ou_conv_params& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // shared memory carve-up
  uint8_t* smA = smem;
  uint8_t* smB = smA + (size_t)a.a_stages * a.a_stage_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smB + (size_t)a.b_stages * a.b_stage_bytes);
  uint64_t* full_a = bars;                       // [a_stages]
  uint64_t* ready_a = full_a + MAX_STAGES;       // [a_stages]
  uint64_t* empty_a = ready_a + MAX_STAGES;      // [a_stages]
  uint64_t* full_b = empty_a + MAX_STAGES;       // [b_stages]
  uint64_t* empty_b = full_b + MAX_STAGES;       // [b_stages]
  uint64_t* tmem_full = empty_b + MAX_STAGES;    // [2]
  uint64_t* tmem_empty = tmem_full + 2;          // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  if (threadIdx.x == 0) {
    for (int i = 0; i < a.a_stages; i++) {
      mbar_init(smem_u32(&full_a[i]), 1);
      mbar_init(smem_u32(&ready_a[i]), 4);
      mbar_init(smem_u32(&empty_a[i]), 1);
    }
    for (int i = 0; i < a.b_stages; i++) {
      mbar_init(smem_u32(&full_b[i]), 1);
      mbar_init(smem_u32(&empty_b[i]), 1);
    }
    for (int i = 0; i < 2; i++) {
      mbar_init(smem_u32(&tmem_full[i]), 1);
      mbar_init(smem_u32(&tmem_empty[i]), 4);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(tmem_slot), a.tmem_cols);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  // static tile schedule: this CTA owns N tile `nt` and M tiles mt0, mt0 + stride, ...
  const int nt = blockIdx.x % a.n_ntiles;
  const int mt0 = blockIdx.x / a.n_ntiles;
  const int mt_stride = a.ctas_per_ntile;
  const int n0 = nt * a.bn;
  const int taps = p.taps;

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      const __nv_bfloat16* wg = (const __nv_bfloat16*)p.w;
      Ring ra, rb;
      bool b_loaded = false;
      int ti = 0;
      for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
        const int b = mt / a.m_tiles;
        const int m0 = (mt - b * a.m_tiles) * BM;
        const __nv_bfloat16* xg = (const __nv_bfloat16*)p.x + (size_t)b * a.cin_chunks * p.t_in * 8;
        // valid input rows of this tile: rr in [r_lo, r_hi)  <->  0 <= m0 + rr + tap_off < t_in
        const int j0 = m0 + p.tap_off;
        const int r_lo = j0 < 0 ? -j0 : 0;
        int r_hi = p.t_in - j0;
        r_hi = r_hi > a.arows ? a.arows : r_hi;
        const int nrows = r_hi > r_lo ? r_hi - r_lo : 0;
        for (int kb = 0; kb < a.n_kblocks; kb++) {
          if (!(a.resident && b_loaded)) {
            mbar_wait(smem_u32(&empty_b[rb.stage]), rb.phase ^ 1);
            const uint32_t bar = smem_u32(&full_b[rb.stage]);
            mbar_arrive_expect_tx(bar, a.b_stage_bytes);
            const uint32_t dst0 = smem_u32(smB + (size_t)rb.stage * a.b_stage_bytes);
            for (int q = 0; q < taps; q++)
              for (int c = 0; c < a.kb_chunks; c++) {
                const size_t src =
                    (((size_t)q * (p.kpad >> 3) + kb * a.kb_chunks + c) * p.npad + n0) * 8;
                bulk_g2s(dst0 + (uint32_t)((q * a.kb_chunks + c) * a.bn * 16), wg + src,
                         (uint32_t)a.bn * 16, bar);
              }
            rb.advance(a.b_stages);
          }
          mbar_wait(smem_u32(&empty_a[ra.stage]), ra.phase ^ 1);
          if (kb == 0) trace_ev(a, 0, ti, 0);
          const uint32_t bar = smem_u32(&full_a[ra.stage]);
          int nvalid = a.cin_chunks - kb * a.kb_chunks;
          nvalid = nvalid > a.kb_chunks ? a.kb_chunks : (nvalid < 0 ? 0 : nvalid);
          mbar_arrive_expect_tx(bar, (uint32_t)(nvalid * nrows * 16));
          if (nrows > 0) {
            const uint32_t dst0 = smem_u32(smA + (size_t)ra.stage * a.a_stage_bytes);
            for (int c = 0; c < nvalid; c++) {
              const int cg = kb * a.kb_chunks + c;
              bulk_g2s(dst0 + (uint32_t)((c * a.arows + r_lo) * 16),
                       xg + ((size_t)cg * p.t_in + (j0 + r_lo)) * 8, (uint32_t)nrows * 16, bar);
            }
          }
          ra.advance(a.a_stages);
          if (kb == a.n_kblocks - 1) trace_ev(a, 0, ti, 1);
        }
        b_loaded = true;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    Ring ra, rb;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool b_waited = false;
    int ti = 0;
    for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      tc_fence_after();
      if (lane == 0) trace_ev(a, 1, ti, 0);
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.bn);
      for (int kb = 0; kb < a.n_kblocks; kb++) {
        int bstage;
        if (a.resident) {
          bstage = kb;
          if (!b_waited) mbar_wait(smem_u32(&full_b[kb]), 0);
        } else {
          bstage = rb.stage;
          mbar_wait(smem_u32(&full_b[rb.stage]), rb.phase);
        }
        mbar_wait(smem_u32(&ready_a[ra.stage]), ra.phase);
        tc_fence_after();
        if (lane == 0) {
          if (kb == 0) trace_ev(a, 1, ti, 1);
          if (kb == a.n_kblocks - 1) trace_ev(a, 1, ti, 2);
          const uint32_t a_base = smem_u32(smA + (size_t)ra.stage * a.a_stage_bytes);
          const uint32_t b_base = smem_u32(smB + (size_t)bstage * a.b_stage_bytes);
          for (int q = 0; q < taps; q++) {
            for (int kk = 0; kk < a.kb_chunks / 2; kk++) {
              const uint64_t ad =
                  make_desc(a_base + (uint32_t)((2 * kk * a.arows + q) * 16), a.arows * 16, 128);
              const uint64_t bd = make_desc(b_base + (uint32_t)(((q * a.kb_chunks + 2 * kk) * a.bn) * 16),
                                            a.bn * 16, 128);
              umma_f16(d_tmem, ad, bd, a.idesc, (kb | q | kk) != 0 ? 1u : 0u);
            }
          }
          umma_commit(smem_u32(&empty_a[ra.stage]));
          if (!a.resident) umma_commit(smem_u32(&empty_b[rb.stage]));
          if (kb == a.n_kblocks - 1) {
            umma_commit(smem_u32(&tmem_full[acc]));
            trace_ev(a, 1, ti, 3);
          }
        }
        __syncwarp();
        ra.advance(a.a_stages);
        if (!a.resident) rb.advance(a.b_stages);
      }
      b_waited = true;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp < EPI_WARP0) {
    // ================================ transform ================================
    const int xt = threadIdx.x - XF_WARP0 * 32;   // 0..127
    Ring ra;
    const bool do_prelu = p.has_prelu_in != 0;
    const float slope = p.prelu_in;
    int ti = 0;
    for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
      const int b = mt / a.m_tiles;
      const int m0 = (mt - b * a.m_tiles) * BM;
      const int j0 = m0 + p.tap_off;
      const int r_lo = j0 < 0 ? -j0 : 0;
      int r_hi = p.t_in - j0;
      r_hi = r_hi > a.arows ? a.arows : r_hi;
      const bool interior = (r_lo == 0 && r_hi == a.arows);
      for (int kb = 0; kb < a.n_kblocks; kb++) {
        mbar_wait(smem_u32(&full_a[ra.stage]), ra.phase);
        if (xt == 0 && kb == 0) trace_ev(a, 2, ti, 0);
        if (xt == 0 && kb == a.n_kblocks - 1) trace_ev(a, 2, ti, 1);
        uint8_t* As = smA + (size_t)ra.stage * a.a_stage_bytes;
        int nvalid = a.cin_chunks - kb * a.kb_chunks;
        nvalid = nvalid > a.kb_chunks ? a.kb_chunks : (nvalid < 0 ? 0 : nvalid);
        if (do_prelu || !interior || nvalid < a.kb_chunks) {
          const int total = a.kb_chunks * a.arows;
          for (int i = xt; i < total; i += 128) {
            const int c = i / a.arows, rr = i - c * a.arows;
            uint4* ptr = reinterpret_cast<uint4*>(As + (size_t)i * 16);
            if (c >= nvalid || rr < r_lo || rr >= r_hi) {
              *ptr = make_uint4(0u, 0u, 0u, 0u);
            } else if (do_prelu) {
              uint4 v = *ptr;
              uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const float2 f = bf2_to_f2(w[k]);
                w[k] = f2_to_bf2(prelu_f(f.x, slope), prelu_f(f.y, slope));
              }
              *ptr = v;
            }
          }
          fence_proxy_async();
        }
        __syncwarp();
        if (lane == 0) mbar_arrive(smem_u32(&ready_a[ra.stage]));
        if (xt == 0 && kb == a.n_kblocks - 1) trace_ev(a, 2, ti, 2);
        ra.advance(a.a_stages);
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ti = 0;
    for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
      const int b = mt / a.m_tiles;
      const int m0 = (mt - b * a.m_tiles) * BM;
      const int j = m0 + row;
      if (row == 0) trace_ev(a, 3, ti, 0);
      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      if (row == 0) trace_ev(a, 3, ti, 1);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * a.bn);
      for (int c0 = 0; c0 < a.bn; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        if (c0 + 32 >= a.bn) {
          // accumulator fully read: hand the TMEM buffer back before the (long) store phase
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
          if (row == 0) trace_ev(a, 3, ti, 2);
        }
        if (j < p.rows) {
#pragma unroll
          for (int g = 0; g < 4; g++) {
            const int n = n0 + c0 + g * 8;
            if (n < p.n) epilogue_vec8(p, b, j, n, &r[g * 8]);
          }
        }
      }
      if (row == 0) trace_ev(a, 3, ti, 3);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, a.tmem_cols);
}

// ------------------------------------------------------------------------------------ host side
static int g_num_sms = 0;
long long* g_trace = nullptr;

int plan(const ou_conv_params* p, TcArgs* a) {
  if (p->s != 1) return OU_ERR_UNSUPPORTED;
  int bn = 0;
  for (int cand : {256, 128, 64, 32})
    if (p->npad % cand == 0) {
      bn = cand;
      break;
    }
  if (!bn || p->cin % 8 || p->cout % 8) return OU_ERR_UNSUPPORTED;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  a->p = *p;
  a->cin_chunks = p->cin / 8;
  a->arows = BM + p->taps - 1;
  a->bn = bn;
  a->n_ntiles = p->npad / bn;
  a->m_tiles = ceil_div(p->rows, BM);
  a->total_m_tiles = a->m_tiles * p->batch;
  const int budget = 220 * 1024 - 1024;   // dynamic smem minus barriers / bookkeeping
  bool ok = false;
  for (int kb : {64, 32, 16}) {
    if (p->kpad % kb) continue;
    const int kbc = kb / 8;
    const int a_bytes = kbc * a->arows * 16;
    const int b_bytes = p->taps * kbc * bn * 16;
    const int nkb = p->kpad / kb;
    if (nkb <= MAX_STAGES && nkb * b_bytes + 3 * a_bytes <= budget) {
      a->resident = 1;
      a->b_stages = nkb;
      int as = (budget - nkb * b_bytes) / a_bytes;
      a->a_stages = as > 8 ? 8 : as;
    } else if (3 * b_bytes + 3 * a_bytes <= budget) {
      a->resident = 0;
      a->b_stages = 3;
      int as = (budget - 3 * b_bytes) / a_bytes;
      a->a_stages = as > 6 ? 6 : as;
      int bs = (budget - a->a_stages * a_bytes) / b_bytes;
      a->b_stages = bs > 6 ? 6 : bs;
    } else {
      continue;
    }
    a->kb_chunks = kbc;
    a->n_kblocks = nkb;
    a->a_stage_bytes = a_bytes;
    a->b_stage_bytes = b_bytes;
    ok = true;
    break;
  }
  if (!ok) return OU_ERR_UNSUPPORTED;
  // instruction descriptor: D=f32, A=B=bf16, K-major both, N = bn, M = 128
  a->idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * bn)) cols <<= 1;
  a->tmem_cols = cols;
  return OU_OK;
}

int launch(const ou_conv_params* p, cudaStream_t st) {
  TcArgs a;
  int rc = plan(p, &a);
  if (rc) return rc;
  const size_t smem = (size_t)a.a_stages * a.a_stage_bytes + (size_t)a.b_stages * a.b_stage_bytes +
                      (5 * MAX_STAGES + 4) * sizeof(uint64_t) + 16;
  static size_t configured = 0;
  if (smem > configured) {
    cudaError_t e = cudaFuncSetAttribute(conv1d_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) {
      set_error("ou_conv1d(tc): cudaFuncSetAttribute(%zu): %s", smem, cudaGetErrorString(e));
      return OU_ERR_CUDA;
    }
    configured = smem;
  }
  int per = g_num_sms / a.n_ntiles;
  if (per < 1) per = 1;
  if (per > a.total_m_tiles) per = a.total_m_tiles;
  a.ctas_per_ntile = per;
  a.trace = g_trace;
  dim3 grid(per * a.n_ntiles);
  conv1d_tc_kernel<<<grid, NTHREADS, smem, st>>>(a);
  return check_launch("ou_conv1d(tc)");
}

}  // namespace tc
}  // namespace ou

extern "C" int ou_debug_set_trace(void* device_buffer) {
  ou::tc::g_trace = (long long*)device_buffer;
  return OU_OK;
}
