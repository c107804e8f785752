// tcgen05 (5th-gen tensor core) implicit-GEMM Conv1d for sm_100a -- the production path of
// ou_conv1d for stride-1 input geometry (conv1/conv2/conv3, transposed up convs, 1x1s, GRU input
// projections: > 80 % of the FLOPs of a score step).  Contract: include/ou_b200.h.
//
// Persistent, warp-specialised CTA (one per SM):
//   warp 0      producer   : TMA tensor loads (cp.async.bulk.tensor, UTMALDG) global -> smem rings
//                            in the hardware 128B/64B/32B swizzle; out-of-range rows ("same"
//                            padding, tile overrun) are zero-filled by the TMA unit
//   warp 1      MMA issuer : one elected lane issues tcgen05.mma.cta_group::1.kind::f16
//                            (bf16 x bf16 -> fp32 in TMEM); tcgen05.commit frees smem stages and
//                            publishes accumulators; owns the TMEM allocation
//   warps 2-5   transform  : only for layers with a fused input PReLU: element-wise pass over each
//                            landed A stage, fence.proxy.async, hand-over to the MMA warp
//   warps 6-9   epilogue   : tcgen05.ld accumulator rows -> registers; bias / FiLM staged in smem,
//                            residual vectors prefetched before the accumulator is ready; 16-byte
//                            bf16 stores in the blocked layout (or fp32 time-major)
// TMEM holds two accumulator buffers (2 x BN columns): the epilogue of tile i overlaps the MMAs of
// tile i+1.
//
// Operand layout: the blocked activation layout [B][C/CB][T][CB] makes one time step of one
// channel block a contiguous CB*2-byte row, i.e. a K-major operand row; a TMA box of (CB channels x
// 128+taps-1 time steps) lands as the canonical swizzled K-major tile.  A conv tap is a shift by
// whole rows: the A descriptor's start address advances by q rows (base_offset carries the swizzle
// phase), so ONE staged tile feeds all taps.  Weights are pre-packed per (tap, K block) as
// [npad][CB] K-major tiles; a CTA's N-slice stays resident in shared memory across all its M tiles
// when it fits (C <= 128), otherwise it streams from L2 through a ring.
#include <cuda.h>

#include <cstdlib>

#include "common.cuh"

namespace ou {
namespace tc {

constexpr int BM = 128;
constexpr int NTHREADS = 320;
constexpr int XF_WARP0 = 2, EPI_WARP0 = 6;
constexpr int MAX_A_STAGES = 8, MAX_B_STAGES = 16;

struct TcArgs {
  ou_conv_params p;
  int cb;              // channels per K block (= channel block of the input layout: 64 / 32 / 16)
  int row_bytes;       // cb * 2
  int n_kblocks;       // cin / cb
  int arows;           // rows per A stage: BM + taps - 1 (shared by all taps) or BM (per-tap mode)
  int a_per_tap;       // 1: one A stage per (K block, tap) loaded at shifted coordinates
  int bn;              // N per CTA tile
  int n_ntiles;        // npad / bn
  int m_tiles;         // ceil(rows / BM) per clip
  int total_m_tiles;   // m_tiles * batch
  int ctas_per_ntile;  // gridDim.x / n_ntiles
  int a_stages, b_stages;
  int resident;        // whole weight slice stays in smem (b_stages == taps * n_kblocks)
  uint32_t a_stage_bytes, b_stage_bytes;   // 1024-aligned strides
  uint32_t a_tx_bytes, b_tx_bytes;         // bytes one TMA box delivers
  uint32_t idesc;
  uint32_t tmem_cols;
  uint32_t desc_hi;    // descriptor bits [32,64) without base_offset: SBO, version, layout type
  int base_off_mode;   // 1: base_offset = (start_address >> 7) & 7   0: always 0
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
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            int c3, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
      "%4, %5}], [%6];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(uint32_t dst, const CUtensorMap* map, int c0, int c1, int c2,
                                            uint32_t bar) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, "
      "%4}], [%5];" ::"r"(dst),
      "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(bar)
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
__device__ __forceinline__ uint4 ldg_nc_v4(const void* p) {
  uint4 v;
  asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
               : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
               : "l"(p));
  return v;
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// K-major swizzled shared-memory matrix descriptor (cute::UMMA::SmemDescriptor, version 1):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major: 1) | [32,46) SBO>>4 (8 rows)
//   [46,48) version=1 | [49,52) base_offset | [61,64) layout (2 = 128B, 4 = 64B, 6 = 32B swizzle)
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t desc_hi, int base_off_mode) {
  uint32_t hi = desc_hi;
  if (base_off_mode) hi |= ((saddr >> 7) & 7u) << (49 - 32);
  return (uint64_t)((saddr >> 4) & 0x3FFFu) | (1ull << 16) | ((uint64_t)hi << 32);
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

__global__ void __launch_bounds__(NTHREADS, 1)
conv1d_tc_kernel(const TcArgs a, const __grid_constant__ CUtensorMap tm_a,
                 const __grid_constant__ CUtensorMap tm_w) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const ou_conv_params& p = a.p;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // shared memory carve-up (stages 1024-byte aligned for the swizzle pattern)
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smA = smem;
  uint8_t* smB = smA + (size_t)a.a_stages * a.a_stage_bytes;
  uint8_t* tail = smB + (size_t)a.b_stages * a.b_stage_bytes;
  uint64_t* full_a = reinterpret_cast<uint64_t*>(tail);   // [MAX_A_STAGES]
  uint64_t* ready_a = full_a + MAX_A_STAGES;
  uint64_t* empty_a = ready_a + MAX_A_STAGES;
  uint64_t* full_b = empty_a + MAX_A_STAGES;               // [MAX_B_STAGES]
  uint64_t* empty_b = full_b + MAX_B_STAGES;
  uint64_t* tmem_full = empty_b + MAX_B_STAGES;            // [2]
  uint64_t* tmem_empty = tmem_full + 2;                    // [2]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* bias_s = reinterpret_cast<float*>(tmem_slot + 4); // [bn]
  float* film_s = bias_s + a.bn;                            // [2 buffers][gamma bn | beta bn]

  const bool use_xf = p.has_prelu_in != 0;

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
  const int a_loads_per_kb = a.a_per_tap ? taps : 1;

  if (warp == 0) {
    // ================================ producer ================================
    if (lane == 0) {
      Ring ra, rb;
      bool b_loaded = false;
      int ti = 0;
      for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
        const int b = mt / a.m_tiles;
        const int m0 = (mt - b * a.m_tiles) * BM;
        for (int kb = 0; kb < a.n_kblocks; kb++) {
          for (int q = 0; q < taps; q++) {
            if (q < a_loads_per_kb) {
              mbar_wait(smem_u32(&empty_a[ra.stage]), ra.phase ^ 1);
              if (kb == 0 && q == 0) trace_ev(a, 0, ti, 0);
              const uint32_t bar = smem_u32(&full_a[ra.stage]);
              mbar_arrive_expect_tx(bar, a.a_tx_bytes);
              // rows [j0, j0 + arows) of clip b, channel block kb; TMA zero-fills rows outside [0, T)
              const int j0 = m0 + p.tap_off + (a.a_per_tap ? q : 0);
              tma_load_4d(smem_u32(smA + (size_t)ra.stage * a.a_stage_bytes), &tm_a, 0, j0, kb, b, bar);
              ra.advance(a.a_stages);
            }
            if (!(a.resident && b_loaded)) {
              mbar_wait(smem_u32(&empty_b[rb.stage]), rb.phase ^ 1);
              const uint32_t bar = smem_u32(&full_b[rb.stage]);
              mbar_arrive_expect_tx(bar, a.b_tx_bytes);
              tma_load_3d(smem_u32(smB + (size_t)rb.stage * a.b_stage_bytes), &tm_w, 0, n0,
                          q * a.n_kblocks + kb, bar);
              rb.advance(a.b_stages);
            }
          }
        }
        trace_ev(a, 0, ti, 1);
        b_loaded = true;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ================================
    Ring ra, rb;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool b_waited = false;
    const int k16_steps = a.cb / 16;
    int ti = 0;
    for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
      mbar_wait(smem_u32(&tmem_empty[acc]), acc_phase ^ 1);
      tc_fence_after();
      if (lane == 0) trace_ev(a, 1, ti, 0);
      const uint32_t d_tmem = tmem_base + (uint32_t)(acc * a.bn);
      for (int kb = 0; kb < a.n_kblocks; kb++) {
        for (int q = 0; q < taps; q++) {
          if (q < a_loads_per_kb) {
            mbar_wait(smem_u32(use_xf ? &ready_a[ra.stage] : &full_a[ra.stage]), ra.phase);
            if (lane == 0 && kb == 0 && q == 0) trace_ev(a, 1, ti, 1);
          }
          int bstage;
          if (a.resident) {
            bstage = kb * taps + q;
            if (!b_waited) mbar_wait(smem_u32(&full_b[bstage]), 0);
          } else {
            bstage = rb.stage;
            mbar_wait(smem_u32(&full_b[bstage]), rb.phase);
          }
          tc_fence_after();
          if (lane == 0) {
            const uint32_t a_base = smem_u32(smA + (size_t)ra.stage * a.a_stage_bytes) +
                                    (a.a_per_tap ? 0u : (uint32_t)(q * a.row_bytes));
            const uint32_t b_base = smem_u32(smB + (size_t)bstage * a.b_stage_bytes);
            for (int kk = 0; kk < k16_steps; kk++) {
              const uint64_t ad = make_desc(a_base + kk * 32, a.desc_hi, a.base_off_mode);
              const uint64_t bd = make_desc(b_base + kk * 32, a.desc_hi, 0);
              umma_f16(d_tmem, ad, bd, a.idesc, (kb | q | kk) != 0 ? 1u : 0u);
            }
            if (!a.resident) umma_commit(smem_u32(&empty_b[bstage]));
            const bool a_done = a.a_per_tap || q == taps - 1;
            if (a_done) umma_commit(smem_u32(&empty_a[ra.stage]));
            if (kb == a.n_kblocks - 1 && q == taps - 1) {
              umma_commit(smem_u32(&tmem_full[acc]));
              trace_ev(a, 1, ti, 3);
            }
          }
          __syncwarp();
          if (!a.resident) rb.advance(a.b_stages);
          if (a.a_per_tap || q == taps - 1) ra.advance(a.a_stages);
        }
      }
      b_waited = true;
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp < EPI_WARP0) {
    // ================================ transform (fused input PReLU) ================================
    if (use_xf) {
      const int xt = threadIdx.x - XF_WARP0 * 32;   // 0..127
      Ring ra;
      const float slope = p.prelu_in;
      const int nvec = (int)(a.a_tx_bytes >> 4);
      int ti = 0;
      for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
        for (int kb = 0; kb < a.n_kblocks; kb++) {
          for (int q = 0; q < a_loads_per_kb; q++) {
            mbar_wait(smem_u32(&full_a[ra.stage]), ra.phase);
            if (xt == 0 && kb == 0 && q == 0) trace_ev(a, 2, ti, 0);
            uint4* As = reinterpret_cast<uint4*>(smA + (size_t)ra.stage * a.a_stage_bytes);
#pragma unroll 4
            for (int i = xt; i < nvec; i += 128) {
              uint4 v = As[i];
              uint32_t* w = reinterpret_cast<uint32_t*>(&v);
#pragma unroll
              for (int k = 0; k < 4; k++) {
                const float2 f = bf2_to_f2(w[k]);
                w[k] = f2_to_bf2(prelu_f(f.x, slope), prelu_f(f.y, slope));
              }
              As[i] = v;
            }
            fence_proxy_async();
            __syncwarp();
            if (lane == 0) mbar_arrive(smem_u32(&ready_a[ra.stage]));
            if (xt == 0 && kb == a.n_kblocks - 1) trace_ev(a, 2, ti, 2);
            ra.advance(a.a_stages);
          }
        }
      }
    }
  } else {
    // ================================ epilogue ================================
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int row = quarter * 32 + lane;
    const int et = threadIdx.x - EPI_WARP0 * 32;  // 0..127
    const int cbo = cl_cb(p.cout);
    const bool blocked_out = p.out_f32_tm == nullptr;
    const __nv_bfloat16* add1 = (const __nv_bfloat16*)p.add1;
    const __nv_bfloat16* add2 = (const __nv_bfloat16*)p.add2;
    for (int i = et; i < a.bn; i += 128) bias_s[i] = (p.bias && n0 + i < p.n) ? p.bias[n0 + i] : 0.f;
    int acc = 0;
    uint32_t acc_phase = 0;
    int ti = 0;
    for (int mt = mt0; mt < a.total_m_tiles; mt += mt_stride, ti++) {
      const int b = mt / a.m_tiles;
      const int m0 = (mt - b * a.m_tiles) * BM;
      const int j = m0 + row;
      const bool row_ok = j < p.rows;
      float* film = film_s + acc * 2 * a.bn;
      if (p.gamma) {
        for (int i = et; i < a.bn; i += 128) {
          const int n = n0 + i;
          const int co = n % p.cout;
          const bool ok = n < p.n;
          film[i] = ok ? p.gamma[(size_t)b * p.film_bstride + co] : 0.f;
          film[a.bn + i] = ok ? p.beta[(size_t)b * p.film_bstride + co] : 0.f;
        }
      }
      epi_bar_sync();   // bias_s / film_s visible to the four epilogue warps
      if (row == 0) trace_ev(a, 3, ti, 0);

      // residual vectors of the first 32-column chunk are fetched while the MMAs still run
      uint4 pre1[4], pre2[4];
      auto out_offset = [&](int n) -> long {
        const int ph = n / p.cout;
        const int co = n - ph * p.cout;
        const int t = j * p.up + ph;
        if (!row_ok || n >= p.n || t >= p.t_out) return -1;
        return (long)cl_off(b, co, t, p.cout, p.t_out, cbo);
      };
      auto prefetch = [&](int c0) {
#pragma unroll
        for (int g = 0; g < 4; g++) {
          pre1[g] = make_uint4(0u, 0u, 0u, 0u);
          pre2[g] = make_uint4(0u, 0u, 0u, 0u);
          if (blocked_out && (add1 || add2)) {
            const long off = out_offset(n0 + c0 + g * 8);
            if (off >= 0) {
              if (add1) pre1[g] = ldg_nc_v4(add1 + off);
              if (add2) pre2[g] = ldg_nc_v4(add2 + off);
            }
          }
        }
      };
      prefetch(0);

      mbar_wait(smem_u32(&tmem_full[acc]), acc_phase);
      tc_fence_after();
      if (row == 0) trace_ev(a, 3, ti, 1);
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (uint32_t)(acc * a.bn);
      for (int c0 = 0; c0 < a.bn; c0 += 32) {
        uint32_t r[32];
        tmem_ld32(taddr + (uint32_t)c0, r);
        tmem_ld_wait();
        const bool last = c0 + 32 >= a.bn;
        if (last) {
          // accumulator fully read: hand the TMEM buffer back before the store phase
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(smem_u32(&tmem_empty[acc]));
          if (row == 0) trace_ev(a, 3, ti, 2);
        }
        uint4 cur1[4], cur2[4];
#pragma unroll
        for (int g = 0; g < 4; g++) cur1[g] = pre1[g], cur2[g] = pre2[g];
        if (!last) prefetch(c0 + 32);
#pragma unroll
        for (int g = 0; g < 4; g++) {
          const int nl = c0 + g * 8;          // column within the CTA tile
          const int n = n0 + nl;
          if (!row_ok || n >= p.n) continue;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] = __uint_as_float(r[g * 8 + i]) + bias_s[nl + i];
          if (!blocked_out) {
            float4* dst = reinterpret_cast<float4*>(p.out_f32_tm + ((size_t)b * p.rows + j) * p.n + n);
            dst[0] = make_float4(v[0], v[1], v[2], v[3]);
            dst[1] = make_float4(v[4], v[5], v[6], v[7]);
            continue;
          }
          const long off = out_offset(n);
          if (off < 0) continue;
          if (add1) {
            const uint32_t* pa = reinterpret_cast<const uint32_t*>(&cur1[g]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float2 f = bf2_to_f2(pa[i]);
              v[2 * i] += f.x, v[2 * i + 1] += f.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] *= p.scale1;
          if (add2) {
            const uint32_t* pa = reinterpret_cast<const uint32_t*>(&cur2[g]);
#pragma unroll
            for (int i = 0; i < 4; i++) {
              const float2 f = bf2_to_f2(pa[i]);
              v[2 * i] += f.x, v[2 * i + 1] += f.y;
            }
          }
#pragma unroll
          for (int i = 0; i < 8; i++) v[i] *= p.scale2;
          if (p.gamma) {
#pragma unroll
            for (int i = 0; i < 8; i++) v[i] = fmaf(film[nl + i], v[i], film[a.bn + nl + i]);
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

static int env_int(const char* name, int dflt) {
  const char* e = getenv(name);
  return e ? atoi(e) : dflt;
}

int plan(const ou_conv_params* p, TcArgs* a) {
  if (p->s != 1 || p->w_tc == nullptr) return OU_ERR_UNSUPPORTED;
  int bn = 0;
  for (int cand : {256, 128, 64, 32})
    if (p->npad % cand == 0) {
      bn = cand;
      break;
    }
  if (!bn || p->cin % 16 || p->cout % 16) return OU_ERR_UNSUPPORTED;
  if (g_num_sms == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&g_num_sms, cudaDevAttrMultiProcessorCount, dev);
    if (g_num_sms <= 0) g_num_sms = 148;
  }
  static const int force_per_tap = env_int("OU_TC_PER_TAP", -1);
  static const int base_off_mode = env_int("OU_TC_BASEOFF", 1);
  a->p = *p;
  a->cb = cl_cb(p->cin);
  a->row_bytes = a->cb * 2;
  a->n_kblocks = p->cin / a->cb;
  a->bn = bn;
  a->n_ntiles = p->npad / bn;
  a->m_tiles = ceil_div(p->rows, BM);
  a->total_m_tiles = a->m_tiles * p->batch;
  // One staged A tile serves all taps through row-shifted descriptors when the row is a full
  // 128-byte swizzle span; narrower rows (C = 32, 48, 80, 96) load one tile per tap instead.
  a->a_per_tap = (p->taps > 1 && a->row_bytes < 128) ? 1 : 0;
  if (force_per_tap >= 0 && p->taps > 1) a->a_per_tap = force_per_tap;
  a->base_off_mode = base_off_mode;
  a->arows = a->a_per_tap ? BM : BM + p->taps - 1;
  a->a_tx_bytes = (uint32_t)(a->arows * a->row_bytes);
  a->b_tx_bytes = (uint32_t)(bn * a->row_bytes);
  a->a_stage_bytes = (a->a_tx_bytes + 1023u) & ~1023u;
  a->b_stage_bytes = (a->b_tx_bytes + 1023u) & ~1023u;
  const int budget = 232448 - 2048 - 5 * bn * 4;   // 227 KB minus alignment slack, barriers, bias / FiLM
  const int nb_all = p->taps * a->n_kblocks;
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
  a->idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(bn >> 3) << 17) | ((uint32_t)(BM >> 4) << 24);
  uint32_t cols = 32;
  while (cols < (uint32_t)(2 * bn)) cols <<= 1;
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
    // activations: [B][C/CB][T][CB] bf16 -> dims (CB, T, C/CB, B)
    cuuint64_t dims[4] = {(cuuint64_t)a.cb, (cuuint64_t)p->t_in, (cuuint64_t)a.n_kblocks, (cuuint64_t)p->batch};
    cuuint64_t strides[3] = {(cuuint64_t)a.row_bytes, (cuuint64_t)p->t_in * a.row_bytes,
                             (cuuint64_t)p->t_in * a.row_bytes * a.n_kblocks};
    cuuint32_t box[4] = {(cuuint32_t)a.cb, (cuuint32_t)a.arows, 1, 1};
    cuuint32_t estr[4] = {1, 1, 1, 1};
    CUresult r = g_encode(&tm_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(p->x), dims, strides,
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
    CUresult r = g_encode(&tm_w, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, const_cast<void*>(p->w_tc), dims, strides,
                          box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, swizzle_for(a.row_bytes),
                          CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      set_error("ou_conv1d(tc): cuTensorMapEncodeTiled(W) failed with %d", (int)r);
      return OU_ERR_CUDA;
    }
  }
  const size_t smem = 1024 + (size_t)a.a_stages * a.a_stage_bytes + (size_t)a.b_stages * a.b_stage_bytes +
                      (3 * MAX_A_STAGES + 2 * MAX_B_STAGES + 4) * sizeof(uint64_t) + 16 +
                      (size_t)5 * a.bn * sizeof(float);
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
  conv1d_tc_kernel<<<grid, NTHREADS, smem, st>>>(a, tm_a, tm_w);
  return check_launch("ou_conv1d(tc)");
}

}  // namespace tc
}  // namespace ou

extern "C" int ou_debug_set_trace(void* device_buffer) {
  ou::tc::g_trace = (long long*)device_buffer;
  return OU_OK;
}
