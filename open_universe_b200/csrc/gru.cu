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
  const __nv_bfloat16* add;
  __nv_bfloat16* out;
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
    gxp = a.gx + (size_t)(b0 + fb) * T * 6 * H + (size_t)dir * 3 * H + hu;
    const int ch = dir * H + hu;
    out_base = cl_off(b0 + fb, ch, 0, 2 * H, T, cl_cb(2 * H));
  }
  cluster.sync();

  // input pre-activations are prefetched one step ahead (they come from HBM / L2)
  float gxr = 0.f, gxz = 0.f, gxn = 0.f;
  if (fvalid) {
    const float* g = gxp + (size_t)(dir ? T - 1 : 0) * 6 * H;
    gxr = __ldg(g), gxz = __ldg(g + H), gxn = __ldg(g + 2 * H);
  }
  for (int step = 0; step < T; step++) {
    const int t = dir ? (T - 1 - step) : step;
    const int cur = step & 1;
    float nxr = 0.f, nxz = 0.f, nxn = 0.f;
    if (fvalid && step + 1 < T) {
      const float* g = gxp + (size_t)(dir ? t - 1 : t + 1) * 6 * H;
      nxr = __ldg(g), nxz = __ldg(g + H), nxn = __ldg(g + 2 * H);
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
      if (a.add) v += __bfloat162float(a.add[off]);
      a.out[off] = __float2bfloat16(v * a.scale);
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
// Tensor-core variant for H = 256 (UNIVERSE / UNIVERSE++ 16 kHz): the per-step mat-vec
// W_hh[96 x 256] . h[256 x clips] of each CTA runs on mma.sync m16n8k8 TF32 (fp32 accumulate) with
// the weight fragments register-resident; up to 8 clips per cluster cost the same MMAs, so a batch
// of 32 needs only 8 clusters.  TF32 is what the reference's own CUDA path (cuDNN, allow_tf32)
// uses for the recurrent product; the hidden state itself, the gates and the z*h blend stay fp32
// (each gate-stage thread keeps its h in a register).
constexpr int GTC_H = 256, GTC_CS = 8, GTC_HS = 32, GTC_ROWS = 96, GTC_NT = 384, GTC_HP = GTC_H + 4;

__device__ __forceinline__ uint32_t to_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return r;
}
__device__ __forceinline__ void mma_tf32(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
      "{%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

// Exchange of the new hidden state: every gate-stage warp holds, after four shuffles, float4s of
// (clip, 4 consecutive units) and sends them with st.async straight from registers into h_buf of all
// eight CTAs; each store signals complete_tx on the DESTINATION CTA's mbarrier, so a CTA starts the
// next step as soon as the 8 KB of h_t have landed in its own shared memory -- no cluster barrier,
// no staging buffer.  h_buf is double buffered; a CTA can only push h_t after it has received all of
// h_{t-1}, i.e. after every peer has finished reading the buffer h_t goes into.
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t cta) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(cta));
  return r;
}
__device__ __forceinline__ void st_async_v4(uint32_t remote_addr, float4 v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.v4.b32 [%0], {%1, %2, %3, %4}, [%5];"
               ::"r"(remote_addr), "r"(__float_as_uint(v.x)), "r"(__float_as_uint(v.y)),
               "r"(__float_as_uint(v.z)), "r"(__float_as_uint(v.w)), "r"(remote_bar)
               : "memory");
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

__global__ void __launch_bounds__(GTC_NT) gru_cluster_tc_kernel(const GruArgs a) {
  constexpr int H = GTC_H, CS = GTC_CS, HS = GTC_HS, ROWS = GTC_ROWS, HP = GTC_HP, BG = 8;
  // h (TF32-rounded) of 8 clip slots, row stride H+4 floats: the B-fragment loads (clip = lane/4,
  // k = lane%4) then hit 32 distinct banks
  __shared__ __align__(16) float h_buf[2][BG][HP];
  __shared__ float part[2][ROWS][BG];
  __shared__ __align__(8) unsigned long long h_full[2];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int dir = cid & 1;
  const int b0 = (cid >> 1) * BG;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int g = lane >> 2, t4 = lane & 3;
  const int T = a.t;

  // warp -> (16-row tile, K half); weight fragments for its 16 k8-steps stay in registers
  const int mt = warp % 6, khalf = warp / 6;
  const int gate = mt >> 1, u0 = (mt & 1) * 16;
  uint32_t wfrag[16][4];
  {
    const float* wbase = a.w_hh + ((size_t)dir * 3 * H + (size_t)gate * H + rank * HS + u0) * H;
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const int k = (khalf * 16 + i) * 8 + t4;
      wfrag[i][0] = to_tf32(wbase[(size_t)g * H + k]);
      wfrag[i][1] = to_tf32(wbase[(size_t)(g + 8) * H + k]);
      wfrag[i][2] = to_tf32(wbase[(size_t)g * H + k + 4]);
      wfrag[i][3] = to_tf32(wbase[(size_t)(g + 8) * H + k + 4]);
    }
  }
  for (int i = tid; i < 2 * BG * HP; i += GTC_NT) (&h_buf[0][0][0])[i] = 0.f;
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(&h_full[0]);
  if (tid == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0));
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar0 + 8));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    // h_{-1} = 0 is already in h_buf[0]: complete phase 0 of its barrier by hand
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar0) : "memory");
  }

  // gate-stage role (warps 0-7): thread <-> (owned unit fu, clip fb); lane = (fu % 4) * 8 + fb, so a
  // warp owns 4 consecutive units of all 8 clips
  const bool fin = tid < HS * BG;
  const int fb = tid % BG, fu = tid / BG;
  const int hu = rank * HS + fu;
  const bool fvalid = fin && (b0 + fb) < a.batch;
  float bhr = 0.f, bhz = 0.f, bhn = 0.f, hprev = 0.f;
  const float* gxp = nullptr;
  size_t out_base = 0;
  if (fvalid) {
    const float* bh = a.b_hh + (size_t)dir * 3 * H;
    bhr = bh[hu], bhz = bh[H + hu], bhn = bh[2 * H + hu];
    gxp = a.gx + (size_t)(b0 + fb) * T * 6 * H + (size_t)dir * 3 * H + hu;
    out_base = cl_off(b0 + fb, dir * H + hu, 0, 2 * H, T, cl_cb(2 * H));
  }
  const bool has_add = a.add != nullptr;
  // push targets of this lane: clip fb, units rank*HS + 4*warp .. +3, destination CTAs lane/8 and
  // lane/8 + 4 (shared::cluster addresses of their h_buf[0] element and h_full[0])
  uint32_t dst_h[2], dst_bar[2];
  {
    const uint32_t local = (uint32_t)__cvta_generic_to_shared(&h_buf[0][fb][rank * HS + 4 * (warp & 7)]);
#pragma unroll
    for (int j = 0; j < 2; j++) {
      dst_h[j] = mapa_u32(local, (uint32_t)((lane >> 3) + 4 * j));
      dst_bar[j] = mapa_u32(bar0, (uint32_t)((lane >> 3) + 4 * j));
    }
  }
  cluster.sync();

  float gxr = 0.f, gxz = 0.f, gxn = 0.f, addv = 0.f;
  if (fvalid) {
    const int t = dir ? T - 1 : 0;
    const float* gp = gxp + (size_t)t * 6 * H;
    gxr = __ldg(gp), gxz = __ldg(gp + H), gxn = __ldg(gp + 2 * H);
    if (has_add) addv = __bfloat162float(a.add[out_base + (size_t)t * cl_cb(2 * H)]);
  }
  for (int step = 0; step < T; step++) {
    const int t = dir ? (T - 1 - step) : step;
    const int cur = step & 1;
    // next step's inputs (HBM / L2) are fetched a whole step ahead
    float nxr = 0.f, nxz = 0.f, nxn = 0.f, nadd = 0.f;
    if (fvalid && step + 1 < T) {
      const int tn = dir ? t - 1 : t + 1;
      const float* gp = gxp + (size_t)tn * 6 * H;
      nxr = __ldg(gp), nxz = __ldg(gp + H), nxn = __ldg(gp + 2 * H);
      if (has_add) nadd = __bfloat162float(a.add[out_base + (size_t)tn * cl_cb(2 * H)]);
    }
    // arm the barrier that collects h_t (8 CTAs x 32 units x 8 clips x 4 B), then wait for h_{t-1}
    if (tid == 0 && step + 1 < T)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar0 + 8u * (cur ^ 1)),
                   "r"((uint32_t)(CS * HS * BG * 4))
                   : "memory");
#define GRU_STAMP(ev) \
  if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && step >= 100 && step < 164) a.trace[(step - 100) * 8 + (ev)] = clock64();
    GRU_STAMP(0)
    gru_mbar_wait(bar0 + 8u * cur, (uint32_t)((step >> 1) & 1));
    GRU_STAMP(1)
    // partial products of this warp: 16 rows x 8 clips over its 128 columns, four accumulators to
    // shorten the dependent MMA chain
    float acc[4][4];
#pragma unroll
    for (int c = 0; c < 4; c++) acc[c][0] = acc[c][1] = acc[c][2] = acc[c][3] = 0.f;
    const float* hb = &h_buf[cur][g][khalf * 128 + t4];
#pragma unroll
    for (int i = 0; i < 16; i += 4) {
#pragma unroll
      for (int c = 0; c < 4; c++)
        mma_tf32(acc[c], wfrag[i + c], __float_as_uint(hb[(i + c) * 8]), __float_as_uint(hb[(i + c) * 8 + 4]));
    }
    {
      const int r0 = mt * 16 + g;
      part[khalf][r0][2 * t4] = (acc[0][0] + acc[1][0]) + (acc[2][0] + acc[3][0]);
      part[khalf][r0][2 * t4 + 1] = (acc[0][1] + acc[1][1]) + (acc[2][1] + acc[3][1]);
      part[khalf][r0 + 8][2 * t4] = (acc[0][2] + acc[1][2]) + (acc[2][2] + acc[3][2]);
      part[khalf][r0 + 8][2 * t4 + 1] = (acc[0][3] + acc[1][3]) + (acc[2][3] + acc[3][3]);
    }
    GRU_STAMP(2)
    __syncthreads();
    GRU_STAMP(3)
    if (fin) {   // whole warps 0-7
      const float hr = bhr + part[0][fu][fb] + part[1][fu][fb];
      const float hz = bhz + part[0][HS + fu][fb] + part[1][HS + fu][fb];
      const float hn = bhn + part[0][2 * HS + fu][fb] + part[1][2 * HS + fu][fb];
      const float r = sigmoid_f(gxr + hr);
      const float z = sigmoid_f(gxz + hz);
      const float n = tanh_f(gxn + r * hn);
      const float hnew = fvalid ? (1.f - z) * n + z * hprev : 0.f;
      hprev = hnew;
      GRU_STAMP(4)
      if (step + 1 < T) {
        const float hq = __uint_as_float(to_tf32(hnew));
        float4 v;
        v.x = __shfl_sync(0xffffffffu, hq, fb);
        v.y = __shfl_sync(0xffffffffu, hq, fb + 8);
        v.z = __shfl_sync(0xffffffffu, hq, fb + 16);
        v.w = __shfl_sync(0xffffffffu, hq, fb + 24);
        const uint32_t boff = (uint32_t)((cur ^ 1) * BG * HP * 4);
#pragma unroll
        for (int j = 0; j < 2; j++) st_async_v4(dst_h[j] + boff, v, dst_bar[j] + 8u * (cur ^ 1));
      }
      GRU_STAMP(5)
      if (fvalid) {
        const size_t off = out_base + (size_t)t * cl_cb(2 * H);
        a.out[off] = __float2bfloat16((hnew + addv) * a.scale);
      }
    }
    gxr = nxr, gxz = nxz, gxn = nxn, addv = nadd;
    GRU_STAMP(6)
  }
  // no CTA may exit while a peer's stores to it could still be in flight
  cluster.sync();
}

static int launch_gru_tc(const GruArgs& a, cudaStream_t st) {
  // 8 clip slots per cluster cost the same MMAs as 1: use as few clusters as the batch allows
  const int clusters = 2 * ceil_div(a.batch, 8);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(clusters * GTC_CS);
  cfg.blockDim = dim3(GTC_NT);
  cfg.dynamicSmemBytes = 0;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = GTC_CS;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  cudaError_t e = cudaLaunchKernelEx(&cfg, gru_cluster_tc_kernel, a);
  if (e != cudaSuccess) {
    set_error("ou_gru_bidir(tc): launch: %s", cudaGetErrorString(e));
    return OU_ERR_CUDA;
  }
  return check_launch("ou_gru_bidir(tc)");
}

// OU_GRU_IMPL=fma forces the CUDA-core kernel for H = 256 (A/B timing, fp32-exact recurrence)
static bool gru_use_tc() {
  static int cached = -1;
  if (cached < 0) {
    const char* e = getenv("OU_GRU_IMPL");
    cached = (e && e[0] == 'f') ? 0 : 1;
  }
  return cached == 1;
}

}  // namespace ou

extern "C" int ou_gru_bidir(const float* gx, const float* w_hh, const float* b_hh, const void* add,
                            float scale, void* out, int batch, int t, int hidden, void* stream) {
  OU_REQUIRE(gx && w_hh && b_hh && out, "ou_gru_bidir: null pointer");
  OU_REQUIRE(batch > 0 && t > 0, "ou_gru_bidir: empty problem");
  ou::GruArgs a{ou::tc::g_trace, gx, w_hh, b_hh, (const __nv_bfloat16*)add, (__nv_bfloat16*)out, scale, batch, t};
  cudaStream_t st = (cudaStream_t)stream;
  switch (hidden) {
    case 128: return ou::launch_gru<128, 4>(a, st);
    case 256: return ou::gru_use_tc() ? ou::launch_gru_tc(a, st) : ou::launch_gru<256, 8>(a, st);
    case 384: return ou::launch_gru<384, 16>(a, st);
    default:
      ou::set_error("ou_gru_bidir: hidden size %d has no kernel (128, 256, 384)", hidden);
      return OU_ERR_UNSUPPORTED;
  }
}
