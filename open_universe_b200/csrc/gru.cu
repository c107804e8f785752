// Bidirectional GRU recurrence as a persistent thread-block-cluster kernel (see ou_b200.h).
//
// One cluster of CS CTAs owns one (direction, group of BG clips).  The 3H x H recurrent matrix is
// split by hidden unit across the cluster and kept REGISTER-resident for the whole sequence
// (each thread holds a 64-wide slice of one gate row), so a time step costs no weight traffic at
// all: h_{t-1} (fp32, BG x H) is read from shared memory as warp-broadcast float4s, partial dot
// products are combined through shared memory, the owning CTA applies the gate math in fp32 and
// pushes its slice of h_t into every CTA of the cluster through distributed shared memory; one
// cluster barrier per step.  Everything is fp32 (the recurrence is the precision-critical part).
#include <cooperative_groups.h>
#include <cuda_fp16.h>

#include <cstdlib>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace ou {
namespace tc {
extern long long* g_trace;
}

// recurrent-matrix columns held per thread: 64 (fewer partial sums) or 32 (twice the threads, half the
// serial FMA chain per step); OU_GRU_KPT selects at run time, default chosen from measurements

struct GruArgs {
  long long* trace;   // debug (ou_debug_set_trace): clock64 stamps of CTA 0 / thread 0, [64 steps][8 events]
  const float* gx;
  const float* w_hh;
  const float* b_hh;
  const act_t* add;
  act_t* out;
  float scale;
  int batch, t;
};

// __expf has ~2 ulp error: |error| of the gates ~1e-7, far below the bf16 output rounding
__device__ __forceinline__ float sigmoid_f(float x) { return __fdividef(1.f, 1.f + __expf(-x)); }
__device__ __forceinline__ float tanh_f(float x) { return __fdividef(2.f, 1.f + __expf(-2.f * x)) - 1.f; }

template <int H, int CS, int BG, int KPT>
__global__ void __launch_bounds__(3 * (H / CS) * (H / KPT)) gru_cluster_kernel(const GruArgs a) {
  constexpr int HS = H / CS;       // hidden units owned by this CTA
  constexpr int ROWS = 3 * HS;     // gate rows owned by this CTA
  constexpr int KS = H / KPT;      // split of the dot product across threads
  constexpr int NT = ROWS * KS;
  static_assert(H % CS == 0 && H % KPT == 0, "bad GRU shape");
  static_assert(HS * BG <= NT, "not enough threads for the gate stage");
  constexpr int FIN_THREADS = ((HS * BG + 31) / 32) * 32;   // gate-stage warps (whole warps)
  static_assert(FIN_THREADS <= NT, "gate-stage warps exceed the CTA");

  __shared__ __align__(16) float h_buf[2][BG][H];
  __shared__ float part[KS][ROWS][BG];
  __shared__ __align__(16) float h_stage[BG][HS];   // this CTA's new h slice, pushed as float4 vectors

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int dir = cid & 1;
  const int b0 = (cid >> 1) * BG;
  const int tid = threadIdx.x;
  const int ks = tid / ROWS;
  const int row = tid - ks * ROWS;
  const int gate = row / HS;
  const int u = row - gate * HS;
  const int T = a.t;

  // register-resident slice of W_hh: row (gate*H + rank*HS + u), columns [ks*KPT, ks*KPT + KPT)
  float w[KPT];
  {
    const float4* src = reinterpret_cast<const float4*>(
        a.w_hh + ((size_t)dir * 3 * H + (size_t)gate * H + rank * HS + u) * H + ks * KPT);
#pragma unroll
    for (int k = 0; k < KPT / 4; k++) {
      float4 v = src[k];
      w[4 * k] = v.x, w[4 * k + 1] = v.y, w[4 * k + 2] = v.z, w[4 * k + 3] = v.w;
    }
  }
  for (int i = tid; i < 2 * BG * H; i += NT) (&h_buf[0][0][0])[i] = 0.f;

  // gate-stage role: one thread per (owned unit, clip)
  const bool fin = tid < HS * BG;
  const int fb = tid % BG, fu = tid / BG;
  const int hu = rank * HS + fu;             // unit index within the direction
  const bool fvalid = fin && (b0 + fb) < a.batch;
  float bhr = 0.f, bhz = 0.f, bhn = 0.f;
  const float* gxp = nullptr;
  size_t out_base = 0;
  if (fvalid) {
    const float* bh = a.b_hh + (size_t)dir * 3 * H;
    bhr = bh[hu], bhz = bh[H + hu], bhn = bh[2 * H + hu];
    gxp = a.gx + f32blk_off(b0 + fb, dir * 3 * H + hu, 0, 6 * H, T);   // + 16 per time step, gates H columns apart
    const int ch = dir * H + hu;
    out_base = cl_off(b0 + fb, ch, 0, 2 * H, T, cl_cb(2 * H));
  }
  cluster.sync();

  // input pre-activations are prefetched one step ahead (they come from HBM / L2)
  const size_t GSTRIDE = (size_t)(H / 16) * T * 16;   // H columns further in the blocked fp32 layout
  float gxr = 0.f, gxz = 0.f, gxn = 0.f;
  if (fvalid) {
    const float* g = gxp + (size_t)(dir ? T - 1 : 0) * 16;
    gxr = __ldg(g), gxz = __ldg(g + GSTRIDE), gxn = __ldg(g + 2 * GSTRIDE);
  }
  for (int step = 0; step < T; step++) {
    const int t = dir ? (T - 1 - step) : step;
    const int cur = step & 1;
    float nxr = 0.f, nxz = 0.f, nxn = 0.f;
    if (fvalid && step + 1 < T) {
      const float* g = gxp + (size_t)(dir ? t - 1 : t + 1) * 16;
      nxr = __ldg(g), nxz = __ldg(g + GSTRIDE), nxn = __ldg(g + 2 * GSTRIDE);
    }
    float acc[BG];
#pragma unroll
    for (int bb = 0; bb < BG; bb++) acc[bb] = 0.f;
#pragma unroll
    for (int k = 0; k < KPT / 4; k++) {
#pragma unroll
      for (int bb = 0; bb < BG; bb++) {
        const float4 hv = *reinterpret_cast<const float4*>(&h_buf[cur][bb][ks * KPT + 4 * k]);
        acc[bb] = fmaf(w[4 * k], hv.x, acc[bb]);
        acc[bb] = fmaf(w[4 * k + 1], hv.y, acc[bb]);
        acc[bb] = fmaf(w[4 * k + 2], hv.z, acc[bb]);
        acc[bb] = fmaf(w[4 * k + 3], hv.w, acc[bb]);
      }
    }
#pragma unroll
    for (int bb = 0; bb < BG; bb++) part[ks][row][bb] = acc[bb];
    __syncthreads();
    float h_buf_own_new = 0.f;
    if (fin) {
      float hr = bhr, hz = bhz, hn = bhn;
#pragma unroll
      for (int k = 0; k < KS; k++) {
        hr += part[k][fu][fb];
        hz += part[k][HS + fu][fb];
        hn += part[k][2 * HS + fu][fb];
      }
      const float r = sigmoid_f(gxr + hr);
      const float z = sigmoid_f(gxz + hz);
      const float n = tanh_f(gxn + r * hn);
      const float hprev = h_buf[cur][fb][hu];
      const float hnew = fvalid ? (1.f - z) * n + z * hprev : 0.f;
      h_buf_own_new = hnew;
      h_stage[fb][fu] = hnew;
    }
    // push the slice to every CTA of the cluster as 16-byte DSMEM stores (4x fewer SM-to-SM packets
    // than per-float stores): BG * HS / 4 vectors per destination
    if (tid < FIN_THREADS) {   // whole warps: named barrier among the gate-stage warps only
      asm volatile("bar.sync 2, %0;" ::"n"(FIN_THREADS) : "memory");
      constexpr int VEC = BG * HS / 4;
      for (int i = tid; i < VEC * CS; i += FIN_THREADS) {
        const int c = i / VEC, v = i - c * VEC;
        const int bb = v / (HS / 4), u4 = v - bb * (HS / 4);
        const float4 val = *reinterpret_cast<const float4*>(&h_stage[bb][u4 * 4]);
        float4* dst = reinterpret_cast<float4*>(&h_buf[cur ^ 1][bb][rank * HS + u4 * 4]);
        *cluster.map_shared_rank(dst, c) = val;
      }
    }
    // split cluster barrier: release the DSMEM pushes, do the global store, then acquire
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    if (fvalid) {
      const float hnew = h_buf_own_new;
      const size_t off = out_base + (size_t)t * cl_cb(2 * H);
      float v = hnew;
      if (a.add) v += act_to_f(a.add[off]);
      a.out[off] = f_to_act(v * a.scale);
    }
    gxr = nxr, gxz = nxz, gxn = nxn;
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
}

template <int H, int CS, int BG, int KPT>
static int launch_gru_k(const GruArgs& a, cudaStream_t st) {
  constexpr int NT = 3 * (H / CS) * (H / KPT);
  auto kern = gru_cluster_kernel<H, CS, BG, KPT>;
  if (CS > 8) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
    if (e != cudaSuccess) {
      set_error("ou_gru_bidir: non-portable cluster size: %s", cudaGetErrorString(e));
      return OU_ERR_CUDA;
    }
  }
  const int clusters = 2 * ceil_div(a.batch, BG);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * CS);
  cfg.blockDim = dim3(NT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    set_error("ou_gru_bidir: launch H=%d CS=%d: %s", H, CS, cudaGetErrorString(e));
    return OU_ERR_CUDA;
  }
  return check_launch("ou_gru_bidir");
}

// How many clusters of CS CTAs can be co-resident (GPC granularity: 14 x 8 on a 148-SM B200).
template <int H, int CS>
static int max_clusters() {
  static int cached = 0;
  if (cached) return cached;
  constexpr int NT = 3 * (H / CS) * (H / 64);
  auto kern = gru_cluster_kernel<H, CS, 4, 64>;
  if (CS > 8) cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(CS * 64);
  cfg.blockDim = dim3(NT);
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess || n < 2) {
    cudaGetLastError();
    n = 2 * (148 / CS / 2);
  }
  cached = n;
  return cached;
}

template <int H, int CS, int BG>
static int launch_gru_bg(const GruArgs& a, cudaStream_t st) {
  static int kpt = 0;
  if (kpt == 0) {
    const char* e = getenv("OU_GRU_KPT");
    kpt = e ? atoi(e) : 64;
  }
  if constexpr (H <= 256 && 3 * (H / CS) * (H / 32) <= 1024) {
    if (kpt == 32) return launch_gru_k<H, CS, BG, 32>(a, st);
  }
  return launch_gru_k<H, CS, BG, 64>(a, st);
}

// Pick the clips-per-cluster so that all clusters run in ONE wave (the recurrence is latency
// bound: a second wave doubles the time, a wider cluster only adds FMA work).
template <int H, int CS>
static int launch_gru(const GruArgs& a, cudaStream_t st) {
  const int groups_max = max_clusters<H, CS>() / 2;
  int bg = ceil_div(a.batch, groups_max > 0 ? groups_max : 1);
  constexpr int NT = 3 * (H / CS) * (H / 64);
  constexpr int BG_MAX = NT / (H / CS);   // the gate stage needs one thread per (unit, clip)
  if (bg <= 2) return launch_gru_bg<H, CS, 2>(a, st);
  if (bg == 3) return launch_gru_bg<H, CS, 3>(a, st);
  if (bg == 4) return launch_gru_bg<H, CS, 4>(a, st);
  if (bg == 5) return launch_gru_bg<H, CS, 5>(a, st);
  if constexpr (BG_MAX >= 8) {
    if (bg == 6) return launch_gru_bg<H, CS, 6>(a, st);
    return launch_gru_bg<H, CS, 8>(a, st);
  } else {
    return launch_gru_bg<H, CS, 6>(a, st);
  }
}


// ------------------------------------------------------------------------------------------------
// Helpers of the tensor-core kernel below.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void gru_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "GRU_WAIT:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra GRU_DONE;\n"
      "bra GRU_WAIT;\n"
      "GRU_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
// debug (ou_debug_set_trace): thread 0 of CTA 0 stamps clock64() for steps 100..163, 8 events per step
#define GRU_STAMP(ev) \
  if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && step >= 100 && step < 164) a.trace[(step - 100) * 8 + (ev)] = clock64();

// ------------------------------------------------------------------------------------------------
// fp16 tensor-core variant (all hidden sizes; the default).  fp16 has the 11-bit significand of the
// TF32 the reference's own CUDA path (cuDNN, allow_tf32) uses for the recurrent product; it is used
// for W_hh and for the exchanged hidden state only: fp32 accumulation, fp32 gates and an fp32 copy of
// h in the owning thread.  (A TF32 mma.sync predecessor with a per-step cluster-wide st.async exchange
// of fp32 state and an intra-CTA barrier ran at 1.36 us per step; this one at 0.8 us.)  What bounds a step of the recurrence is the all-to-all
// exchange of h_t over the SM-to-SM network, which moves roughly one st.async packet per clock per
// SM whatever its size (measured: 512 packets of 16 B and 512 of 8 B both cost ~600 clk), so the
// design minimises PACKETS: fp16 state, 16-byte packets (one clip x 8 units), and as few clip slots
// per cluster as co-residency allows.  There is NO intra-CTA barrier: a warp owns 8 hidden units and
// computes their three gate rows over the FULL K itself as two m16n8k16 tiles, [r(8) | z(8)] and
// [n(8) | 8 zero rows], so every lane ends up with r, z, n of unit g for clips 2*t4 and 2*t4+1
// without any shuffle, applies the gates, and after an 8-shuffle all-gather lane g pushes the two
// (clip, 8 units) vectors into CTA g.  K is visited in a permuted order (lane t4 owns 8 consecutive
// columns of every 32) so that one 16-byte shared load feeds two k16 steps; W_hh fragments use the
// same permutation.  h_buf holds 16-byte (octet, clip) cells at index octet * 10 + clip: the
// B-fragment loads of a quarter warp then hit 32 distinct banks.
template <int H, int CS_ = 8>
struct G16 {
  static constexpr int CS = CS_;         // CTAs per cluster: 8, or 12 (non-portable) so that H = 384 keeps one warp per scheduler
  static constexpr int HS = H / CS;      // hidden units per CTA
  static constexpr int NW = HS / 8;      // warps (8 units each)
  static constexpr int NT = NW * 32;
  static constexpr int KS = H / 16;      // k16 steps
  static constexpr int CELLS = (H / 8) * 10;   // 16-byte cells per h buffer
  // input ring (per warp): XD stages of [8 clip slots][XCLIP floats: gates r, z, n x 8 units, padded so that the
  // four clip pairs of a quarter warp hit distinct banks] + [8 clip slots][8 units] of the 16-bit residual
  static constexpr int XD = 4;
  static constexpr int XCLIP = 28;
  static constexpr int XSTAGE = 8 * XCLIP * 4 + 8 * 16;   // bytes: 1 024
  static_assert(H % 64 == 0 && H % CS == 0 && HS % 8 == 0 && KS % 2 == 0 && CS >= 4 && CS <= 16, "bad GRU shape");
};

// 1 / (1 + 2^x): ex2.approx.ftz + rcp.approx.ftz, ~2 ulp; the argument never leaves [-126, 126] in any
// regime that matters (2^x underflows to 0 -> 1, overflows to inf -> 0: both the correct limits)
__device__ __forceinline__ float rcp1p_ex2(float x) {
  float e, r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(x));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(1.f + e));
  return r;
}

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
  return r;
}
__device__ __forceinline__ void mma_f16(float (&c)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3,
                                        uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void st_async_u4(uint32_t remote_addr, uint4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"(remote_bar)
               : "memory");
}
__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ float lds_f32(uint32_t addr) {
  float v;
  asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ unsigned short lds_u16(uint32_t addr) {
  unsigned short v;
  asm volatile("ld.shared.u16 %0, [%1];" : "=h"(v) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t r;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(sel));
  return r;
}

// BG: clip slots per cluster (4 or 8; the MMA n dimension is always 8); NCH: accumulator chains per
// tile (measured: 4 chains are no faster than 2 -- the k loop is bound by HMMA issue, ~10 clk each);
// LEAN: branch-free gate math on bare ex2 / rcp (halves the gate phase)
template <int H, int BG, int NCH, bool LEAN, int CSZ>
__global__ void __launch_bounds__(G16<H, CSZ>::NT) gru_cluster_f16_kernel(const GruArgs a) {
  using G = G16<H, CSZ>;
  constexpr int CS = G::CS, HS = G::HS, KS = G::KS;
  __shared__ __align__(16) uint4 h_buf[2][G::CELLS];
  __shared__ __align__(16) uint8_t x_ring[G::NW][G::XD][G::XSTAGE];
  __shared__ __align__(8) unsigned long long h_full[2];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int dir = cid & 1;
  const int b0 = (cid >> 1) * BG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int T = a.t;

  // gate stage of this lane: unit hu, clip slots 2*t4 and 2*t4 + 1
  const int hu = rank * HS + 8 * warp + g;
  const int octet = rank * (HS / 8) + warp;

  // register-resident fp16 A fragments.  tile 1: row g = r(unit g), row g + 8 = z(unit g);
  // tile 2: row g = n(unit g), rows 8..15 zero
  uint32_t wr[KS][2], wz[KS][2], wn[KS][2];
  {
    const float* w_r = a.w_hh + ((size_t)dir * 3 * H + hu) * H;
    const float* w_z = w_r + (size_t)H * H;
    const float* w_n = w_z + (size_t)H * H;
#pragma unroll
    for (int i = 0; i < KS; i++) {
      const int p = 32 * (i >> 1) + 8 * t4 + 4 * (i & 1);
      const float4 vr = *reinterpret_cast<const float4*>(w_r + p);
      const float4 vz = *reinterpret_cast<const float4*>(w_z + p);
      const float4 vn = *reinterpret_cast<const float4*>(w_n + p);
      wr[i][0] = pack_h2(vr.x, vr.y), wr[i][1] = pack_h2(vr.z, vr.w);
      wz[i][0] = pack_h2(vz.x, vz.y), wz[i][1] = pack_h2(vz.z, vz.w);
      wn[i][0] = pack_h2(vn.x, vn.y), wn[i][1] = pack_h2(vn.z, vn.w);
    }
  }
  for (int i = tid; i < 2 * G::CELLS; i += G::NT) (&h_buf[0][0])[i] = make_uint4(0, 0, 0, 0);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&h_full[0]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // h_{-1} = 0 is already in h_buf[0]: complete phase 0 of its barrier by hand
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0) : "memory");
  }

  const bool slot_ok = 2 * t4 < BG;               // lanes whose clip slots exist in this cluster
  bool valid[2];
  float bhr = 0.f, bhz = 0.f, bhn = 0.f;
  float hprev[2] = {0.f, 0.f};
  {
    const float* bh = a.b_hh + (size_t)dir * 3 * H;
    bhr = bh[hu], bhz = bh[H + hu], bhn = bh[2 * H + hu];
  }
  const bool has_add = a.add != nullptr;
  const int ocb = cl_cb(2 * H);
  const int t_first = dir ? T - 1 : 0;
  // output element of this lane's (unit, clip) pairs at the first time step, advanced by one row per step
  act_t* outp[2] = {a.out, a.out};
  const ptrdiff_t out_step = dir ? -(ptrdiff_t)ocb : (ptrdiff_t)ocb;
#pragma unroll
  for (int e = 0; e < 2; e++) {
    const int clip = b0 + 2 * t4 + e;
    valid[e] = slot_ok && clip < a.batch;
    if (valid[e]) outp[e] = a.out + cl_off(clip, dir * H + hu, t_first, 2 * H, T, ocb);
  }
  // push target of this lane: CTA g; cells (octet, clip 2*t4) and (octet, clip 2*t4 + 1) of its h_buf[0]
  const bool push1 = CS >= 8 || g < CS;            // clusters of fewer than 8 CTAs: lanes g >= CS have no target
  const uint32_t dst1 = push1 ? (uint32_t)g : 0u;
  const uint32_t dst_h = mapa_u32((uint32_t)__cvta_generic_to_shared(&h_buf[0][octet * 10 + 2 * t4]), dst1);
  const uint32_t dst_bar = mapa_u32(bar0, dst1);
  // clusters of more than 8 CTAs: lanes g < CS - 8 also feed CTA g + 8
  const bool push2 = CS > 8 && g + 8 < CS;
  const uint32_t dst2 = push2 ? (uint32_t)(g + 8) : (uint32_t)g;
  const uint32_t dst_h2 = mapa_u32((uint32_t)__cvta_generic_to_shared(&h_buf[0][octet * 10 + 2 * t4]), dst2);
  const uint32_t dst_bar2 = mapa_u32(bar0, dst2);

  // Input pre-activations (and the residual) stream from HBM / L2 through a per-warp ring of XD stages in
  // shared memory, filled with 16-byte cp.async XD - 1 steps ahead: per step and warp 8 clips x 3 gates x 8
  // units of fp32 gx (48 chunks) + 8 clips x 8 units of the 16-bit residual (8 chunks) = 56 chunks, at most
  // two per lane, each lane advancing its own source pointer by one time step.  (The previous version loaded
  // them into registers two steps ahead: with the loop not unrolled that takes register moves which wait on the
  // loads, per-step address arithmetic and four branches -- 499 of the 1 434 cycles of a step.)
  const uint32_t ring = (uint32_t)__cvta_generic_to_shared(&x_ring[warp][0][0]);
  const char* src[2] = {nullptr, nullptr};
  ptrdiff_t src_step[2] = {0, 0};
  uint32_t dst_off[2] = {0, 0};
#pragma unroll
  for (int j = 0; j < 2; j++) {
    const int c = lane + 32 * j;                    // chunk of the warp's stage
    if (c < 48) {
      const int cs = c / 6, gate = (c % 6) >> 1, half = c & 1, clip = b0 + cs;
      if (cs < BG && clip < a.batch) {
        src[j] = reinterpret_cast<const char*>(
            a.gx + f32blk_off(clip, dir * 3 * H + gate * H + rank * HS + 8 * warp + 4 * half, t_first, 6 * H, T));
        src_step[j] = (dir ? -1 : 1) * (ptrdiff_t)(16 * sizeof(float));
        dst_off[j] = (uint32_t)((cs * G::XCLIP + gate * 8 + half * 4) * 4);
      }
    } else if (c < 56 && has_add) {
      const int cs = c - 48, clip = b0 + cs;
      if (cs < BG && clip < a.batch) {
        src[j] = reinterpret_cast<const char*>(a.add + cl_off(clip, dir * H + rank * HS + 8 * warp, t_first, 2 * H, T, ocb));
        src_step[j] = out_step * (ptrdiff_t)sizeof(act_t);
        dst_off[j] = (uint32_t)(8 * G::XCLIP * 4 + cs * 16);
      }
    }
  }
  const bool has_src0 = src[0] != nullptr, has_src1 = src[1] != nullptr;
  auto prefetch = [&](int stage, bool in_range) {
    if (has_src0 && in_range) cp_async16(ring + (uint32_t)stage * G::XSTAGE + dst_off[0], src[0]);
    if (has_src1 && in_range) cp_async16(ring + (uint32_t)stage * G::XSTAGE + dst_off[1], src[1]);
    asm volatile("cp.async.commit_group;" ::: "memory");
    src[0] += src_step[0], src[1] += src_step[1];
  };
  // the ring of a warp is private to it: zero it (unused clip slots are read, never loaded) before the fill
  for (int i = lane; i < G::XD * G::XSTAGE / 16; i += 32)
    reinterpret_cast<uint4*>(&x_ring[warp][0][0])[i] = make_uint4(0, 0, 0, 0);
  __syncwarp();
  cluster.sync();
#pragma unroll
  for (int j = 0; j < G::XD - 1; j++) prefetch(j, j < T);
  // this lane's values inside a stage: gates r, z, n of unit g for clips 2*t4 and 2*t4 + 1; residual likewise
  const uint32_t rd_x = ring + (uint32_t)((2 * t4 * G::XCLIP + g) * 4);
  const uint32_t rd_a = ring + (uint32_t)(8 * G::XCLIP * 4 + 2 * t4 * 16 + g * 2);

  for (int step = 0; step < T; step++) {
    const int cur = step & 1;
    __syncwarp();   // every lane has read its values of step - 1 out of the stage that is refilled now
    prefetch((step + G::XD - 1) % G::XD, step + G::XD - 1 < T);
    // arm the barrier that collects h_t (H units x BG clips x 2 B), then wait for h_{t-1}
    if (tid == 0 && step + 1 < T)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * (cur ^ 1)),
                   "r"((uint32_t)(H * BG * 2))
                   : "memory");
    asm volatile("cp.async.wait_group %0;" ::"n"(G::XD - 1) : "memory");
    __syncwarp();
    float x0[2][3];
    unsigned short ad0[2];
    {
      const uint32_t so = (uint32_t)(step % G::XD) * G::XSTAGE;
#pragma unroll
      for (int e = 0; e < 2; e++) {
#pragma unroll
        for (int q = 0; q < 3; q++) x0[e][q] = lds_f32(rd_x + so + (uint32_t)((e * G::XCLIP + q * 8) * 4));
        ad0[e] = lds_u16(rd_a + so + (uint32_t)(e * 16));
      }
    }
    GRU_STAMP(0)
    gru_mbar_wait(bar0 + 8u * cur, (uint32_t)((step >> 1) & 1));
    GRU_STAMP(1)
    // NCH independent accumulator chains per tile: the k loop is bound by the HMMA accumulate latency
    float a1[NCH][4], a2[NCH][4];
#pragma unroll
    for (int c = 0; c < NCH; c++)
      a1[c][0] = a1[c][1] = a1[c][2] = a1[c][3] = a2[c][0] = a2[c][1] = a2[c][2] = a2[c][3] = 0.f;
    const uint4* hb = &h_buf[cur][t4 * 10 + g];
#pragma unroll
    for (int i = 0; i < KS; i += 2) {
      const uint4 v = hb[(i >> 1) * 40];   // octet 4 * (i / 2) + t4
      const int c = NCH == 4 ? (i & 2) : 0;
      mma_f16(a1[c], wr[i][0], wz[i][0], wr[i][1], wz[i][1], v.x, v.y);
      mma_f16(a2[c], wn[i][0], 0u, wn[i][1], 0u, v.x, v.y);
      mma_f16(a1[c + 1], wr[i + 1][0], wz[i + 1][0], wr[i + 1][1], wz[i + 1][1], v.z, v.w);
      mma_f16(a2[c + 1], wn[i + 1][0], 0u, wn[i + 1][1], 0u, v.z, v.w);
    }
    GRU_STAMP(2)
    float hnew[2];
    if (LEAN) {
      // gates of both (unit, clip) elements, branch-free so that their MUFU chains interleave
      constexpr float L2E = 1.4426950408889634f;
      float rr[2], zz[2], hn[2];
#pragma unroll
      for (int e = 0; e < 2; e++) {
        float hr = a1[0][e] + a1[1][e], hz = a1[0][2 + e] + a1[1][2 + e];
        hn[e] = a2[0][e] + a2[1][e];
        if (NCH == 4) {
          hr += a1[2][e] + a1[3][e], hz += a1[2][2 + e] + a1[3][2 + e];
          hn[e] += a2[2][e] + a2[3][e];
        }
        hn[e] += bhn;
        rr[e] = rcp1p_ex2(-L2E * (x0[e][0] + (bhr + hr)));
        zz[e] = rcp1p_ex2(-L2E * (x0[e][1] + (bhz + hz)));
      }
#pragma unroll
      for (int e = 0; e < 2; e++) {
        const float n = fmaf(2.f, rcp1p_ex2(-2.f * L2E * fmaf(rr[e], hn[e], x0[e][2])), -1.f);
        const float hv = fmaf(zz[e], hprev[e] - n, n);   // (1 - z) n + z h
        hnew[e] = valid[e] ? hv : 0.f;
        hprev[e] = hnew[e];
      }
    } else {
#pragma unroll
      for (int e = 0; e < 2; e++) {
        float hr = a1[0][e] + a1[1][e], hz = a1[0][2 + e] + a1[1][2 + e], hn = a2[0][e] + a2[1][e];
        if (NCH == 4) {
          hr += a1[2][e] + a1[3][e], hz += a1[2][2 + e] + a1[3][2 + e];
          hn += a2[2][e] + a2[3][e];
        }
        const float r = sigmoid_f(x0[e][0] + bhr + hr);
        const float z = sigmoid_f(x0[e][1] + bhz + hz);
        const float n = tanh_f(x0[e][2] + r * (bhn + hn));
        hnew[e] = valid[e] ? (1.f - z) * n + z * hprev[e] : 0.f;
        hprev[e] = hnew[e];
      }
    }
    GRU_STAMP(3)
    if (step + 1 < T) {
      // all-gather the 8 units of the warp: w[j] = (h(unit j, clip 2*t4), h(unit j, clip 2*t4+1)) as f16x2
      const uint32_t hq = pack_h2(hnew[0], hnew[1]);
      uint32_t w[8];
#pragma unroll
      for (int j = 0; j < 8; j++) w[j] = __shfl_sync(0xffffffffu, hq, t4 + 4 * j);
      if (slot_ok && push1) {
        const uint32_t boff = (uint32_t)((cur ^ 1) * G::CELLS * 16);
        const uint4 lo4 = make_uint4(prmt(w[0], w[1], 0x5410), prmt(w[2], w[3], 0x5410),
                                     prmt(w[4], w[5], 0x5410), prmt(w[6], w[7], 0x5410));
        const uint4 hi4 = make_uint4(prmt(w[0], w[1], 0x7632), prmt(w[2], w[3], 0x7632),
                                     prmt(w[4], w[5], 0x7632), prmt(w[6], w[7], 0x7632));
        st_async_u4(dst_h + boff, lo4, dst_bar + 8u * (cur ^ 1));
        st_async_u4(dst_h + boff + 16u, hi4, dst_bar + 8u * (cur ^ 1));
        if (push2) {
          st_async_u4(dst_h2 + boff, lo4, dst_bar2 + 8u * (cur ^ 1));
          st_async_u4(dst_h2 + boff + 16u, hi4, dst_bar2 + 8u * (cur ^ 1));
        }
      }
    }
    GRU_STAMP(4)
#pragma unroll
    for (int e = 0; e < 2; e++) {
      if (valid[e]) *outp[e] = f_to_act((hnew[e] + act_bits_to_f(ad0[e])) * a.scale);
      outp[e] += out_step;
    }
    GRU_STAMP(5)
  }
  // no CTA may exit while a peer's stores to it could still be in flight
  cluster.sync();
}

template <int H, int BG, int CSZ = 8, int NCH = 2, bool LEAN = true>
static int launch_gru_f16_bg(const GruArgs& a, cudaStream_t st, int* max_clusters) {
  using G = G16<H, CSZ>;
  auto kern = gru_cluster_f16_kernel<H, BG, NCH, LEAN, CSZ>;
  const int clusters = 2 * ceil_div(a.batch, BG);
  if (CSZ > 8) {   // non-portable cluster size: opt in once per device
    static bool allowed[64] = {};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev >= 0 && dev < 64 && !allowed[dev]) {
      if (cudaFuncSetAttribute(kern, cudaFuncAttributeNonPortableClusterSizeAllowed, 1) != cudaSuccess) {
        cudaGetLastError();
        if (max_clusters) *max_clusters = 0;
        return OU_ERR_UNSUPPORTED;
      }
      allowed[dev] = true;
    }
  }
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * G::CS);
  cfg.blockDim = dim3(G::NT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  // Highest launch priority: when the host overlaps this latency-bound recurrence with the convolutions
  // of another half-batch (engine/runtime.py, two capture streams), its few CTAs must get SMs as soon as
  // a conv kernel of the other stream retires, not queue behind that stream's next persistent grid.
  static const int prio = [] {
    int least = 0, greatest = 0;
    if (cudaDeviceGetStreamPriorityRange(&least, &greatest) != cudaSuccess) cudaGetLastError();
    return greatest;
  }();
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = G::CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributePriority;
  attr[1].val.priority = prio;
  cfg.attrs = attr;
  cfg.numAttrs = 2;
  if (max_clusters) {   // query only
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) != cudaSuccess) {
      cudaGetLastError();
      n = 0;
    }
    *max_clusters = n;
    return OU_OK;
  }
  cudaError_t e = cudaLaunchKernelEx(&cfg, kern, a);
  if (e != cudaSuccess) {
    set_error("ou_gru_bidir(f16): launch H=%d BG=%d CS=%d: %s", H, BG, CSZ, cudaGetErrorString(e));
    return OU_ERR_CUDA;
  }
  return check_launch("ou_gru_bidir(f16)");
}

// 8 clip slots per cluster, 4 when the whole batch fits (half the exchange packets for the same cluster count:
// H = 384, 4 clips: 1 054 -> 945 us).  Measured on B200: 4 slots with twice the clusters is no faster while
// every cluster has its SMs to itself and slower once two CTAs share an SM; OU_GRU_BG=4 / 8 forces either.
// H = 384 splits into 48 units = 6 warps per CTA over a cluster of 8, so two of the four schedulers issue the
// HMMAs of two warps.  A cluster of 12 (non-portable size; 4 warps per CTA) was measured: 932 us with 8 slots
// but 1 055 us with 4 (longer exchange), and its 24 CTAs do not fit the SMs the pipelined sampler reserves
// for the recurrence -- kept behind OU_GRU_CS=12 for A/B runs only.
template <int H>
static int launch_gru_f16(const GruArgs& a, cudaStream_t st, int cs_req) {
  static const int forced = [] { const char* e = getenv("OU_GRU_BG"); return e ? atoi(e) : 0; }();
  static const int forced_cs = [] { const char* e = getenv("OU_GRU_CS"); return e ? atoi(e) : 0; }();
  const bool use4 = forced == 4 || (forced == 0 && a.batch <= 4);
  if constexpr (H <= 256) {
    // cluster of 4 (two warps per scheduler: a step takes 1 589 instead of 1 214 cycles, but the recurrence holds
    // half the SMs: 16 x 0.71 ms instead of 32 x 0.53 ms for 16 clips): for hosts that overlap the recurrence with
    // other work and care about SM time, not latency (ou_gru_bidir_ex cluster_ctas = 4; OU_GRU_CS=4 / 8 forces)
    if (forced_cs == 4 || (forced_cs == 0 && cs_req == 4)) return use4 ? launch_gru_f16_bg<H, 4, 4>(a, st, nullptr) : launch_gru_f16_bg<H, 8, 4>(a, st, nullptr);
  }
  if constexpr (H == 384) {
    const int clusters = 2 * ceil_div(a.batch, use4 ? 4 : 8);
    if (forced_cs == 12 && clusters <= 6) {
      static int fit[2] = {-1, -1};   // clusters of 12 that can be co-resident (per BG variant; -1: not asked yet)
      int& f = fit[use4 ? 0 : 1];
      if (f < 0) {
        int n = 0;
        const int rc = use4 ? launch_gru_f16_bg<H, 4, 12>(a, st, &n) : launch_gru_f16_bg<H, 8, 12>(a, st, &n);
        f = rc == OU_OK ? n : 0;
      }
      if (f >= clusters) return use4 ? launch_gru_f16_bg<H, 4, 12>(a, st, nullptr) : launch_gru_f16_bg<H, 8, 12>(a, st, nullptr);
    }
  }
  return use4 ? launch_gru_f16_bg<H, 4>(a, st, nullptr) : launch_gru_f16_bg<H, 8>(a, st, nullptr);
}

// OU_GRU_IMPL = f16 (default) | fma (CUDA cores, fp32-exact recurrence: A/B reference)
static int gru_impl() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("OU_GRU_IMPL");
    cached = (e && e[0] == 'f' && e[1] == 'm') ? 0 : 2;
  }
  return cached;
}

}  // namespace ou

extern "C" int ou_gru_ctas(int hidden, int batch, int cluster_ctas) {
  const int cs = (cluster_ctas == 4 && hidden <= 256) ? 4 : 8;
  const int bg = batch <= 4 ? 4 : 8;
  return 2 * ((batch + bg - 1) / bg) * cs;
}

extern "C" int ou_gru_bidir(const float* gx, const float* w_hh, const float* b_hh, const void* add,
                            float scale, void* out, int batch, int t, int hidden, void* stream) {
  return ou_gru_bidir_ex(gx, w_hh, b_hh, add, scale, out, batch, t, hidden, 0, stream);
}

extern "C" int ou_gru_bidir_ex(const float* gx, const float* w_hh, const float* b_hh, const void* add,
                               float scale, void* out, int batch, int t, int hidden, int cluster_ctas,
                               void* stream) {
  OU_REQUIRE(gx && w_hh && b_hh && out, "ou_gru_bidir: null pointer");
  OU_REQUIRE(batch > 0 && t > 0, "ou_gru_bidir: empty problem");
  OU_REQUIRE(cluster_ctas == 0 || cluster_ctas == 4 || cluster_ctas == 8, "ou_gru_bidir_ex: cluster_ctas must be 0, 4 or 8");
  ou::GruArgs a{ou::tc::g_trace, gx, w_hh, b_hh, (const act_t*)add, (act_t*)out, scale, batch, t};
  cudaStream_t st = (cudaStream_t)stream;
  const int impl = ou::gru_impl();
  switch (hidden) {
    case 128: return impl == 2 ? ou::launch_gru_f16<128>(a, st, cluster_ctas) : ou::launch_gru<128, 4>(a, st);
    case 256:
      return impl == 2 ? ou::launch_gru_f16<256>(a, st, cluster_ctas) : ou::launch_gru<256, 8>(a, st);
    case 384: return impl == 2 ? ou::launch_gru_f16<384>(a, st, cluster_ctas) : ou::launch_gru<384, 16>(a, st);
    default:
      ou::set_error("ou_gru_bidir: hidden size %d has no kernel (128, 256, 384)", hidden);
      return OU_ERR_UNSUPPORTED;
  }
}
