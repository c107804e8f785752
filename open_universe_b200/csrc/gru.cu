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

__global__ void __launch_bounds__(GTC_NT) gru_cluster_tc_kernel(const GruArgs a, const int bg) {
  constexpr int H = GTC_H, CS = GTC_CS, HS = GTC_HS, ROWS = GTC_ROWS, HP = GTC_HP;
  // h (TF32-rounded) of 8 clip slots, row stride H+4 floats: the B-fragment loads (clip = lane/4,
  // k = lane%4) then hit 32 distinct banks
  __shared__ __align__(16) float h_buf[2][8][HP];
  __shared__ float part[2][ROWS][8];
  __shared__ __align__(16) float h_stage[8][HS];

  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const int cid = blockIdx.x / CS;
  const int dir = cid & 1;
  const int b0 = (cid >> 1) * bg;
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
  for (int i = tid; i < 2 * 8 * HP; i += GTC_NT) (&h_buf[0][0][0])[i] = 0.f;

  // gate-stage role: thread <-> (owned unit fu, clip fb)
  const int fin_threads = ((HS * bg + 31) / 32) * 32;
  const bool fin = tid < HS * bg;
  const int fb = tid % bg, fu = tid / bg;
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
  cluster.sync();

  float gxr = 0.f, gxz = 0.f, gxn = 0.f;
  if (fvalid) {
    const float* gp = gxp + (size_t)(dir ? T - 1 : 0) * 6 * H;
    gxr = __ldg(gp), gxz = __ldg(gp + H), gxn = __ldg(gp + 2 * H);
  }
  for (int step = 0; step < T; step++) {
    const int t = dir ? (T - 1 - step) : step;
    const int cur = step & 1;
    float nxr = 0.f, nxz = 0.f, nxn = 0.f;
    if (fvalid && step + 1 < T) {
      const float* gp = gxp + (size_t)(dir ? t - 1 : t + 1) * 6 * H;
      nxr = __ldg(gp), nxz = __ldg(gp + H), nxn = __ldg(gp + 2 * H);
    }
#define GRU_STAMP(ev) \
  if (a.trace != nullptr && blockIdx.x == 0 && tid == 0 && step >= 100 && step < 164) a.trace[(step - 100) * 8 + (ev)] = clock64();
    GRU_STAMP(0)
    // partial products of this warp: 16 rows x 8 clips over its 128 columns, two accumulators to
    // halve the dependent MMA chain
    float acc0[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    const float* hb = &h_buf[cur][g][khalf * 128 + t4];
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      mma_tf32(acc0, wfrag[i], __float_as_uint(hb[i * 8]), __float_as_uint(hb[i * 8 + 4]));
      mma_tf32(acc1, wfrag[i + 1], __float_as_uint(hb[(i + 1) * 8]), __float_as_uint(hb[(i + 1) * 8 + 4]));
    }
    {
      const int r0 = mt * 16 + g;
      part[khalf][r0][2 * t4] = acc0[0] + acc1[0];
      part[khalf][r0][2 * t4 + 1] = acc0[1] + acc1[1];
      part[khalf][r0 + 8][2 * t4] = acc0[2] + acc1[2];
      part[khalf][r0 + 8][2 * t4 + 1] = acc0[3] + acc1[3];
    }
    GRU_STAMP(1)
    __syncthreads();
    GRU_STAMP(2)
    float hnew_keep = 0.f;
    if (fin) {
      const float hr = bhr + part[0][fu][fb] + part[1][fu][fb];
      const float hz = bhz + part[0][HS + fu][fb] + part[1][HS + fu][fb];
      const float hn = bhn + part[0][2 * HS + fu][fb] + part[1][2 * HS + fu][fb];
      const float r = sigmoid_f(gxr + hr);
      const float z = sigmoid_f(gxz + hz);
      const float n = tanh_f(gxn + r * hn);
      const float hnew = fvalid ? (1.f - z) * n + z * hprev : 0.f;
      hprev = hnew;
      hnew_keep = hnew;
      h_stage[fb][fu] = __uint_as_float(to_tf32(hnew));
    }
    GRU_STAMP(3)
    if (tid < fin_threads) {   // whole warps: named barrier among the gate-stage warps only
      asm volatile("bar.sync 2, %0;" ::"r"(fin_threads) : "memory");
      const int vec = bg * HS / 4;
      for (int i = tid; i < vec * CS; i += fin_threads) {
        const int c = i / vec, v = i - c * vec;
        const int bb = v / (HS / 4), u4 = v - bb * (HS / 4);
        const float4 val = *reinterpret_cast<const float4*>(&h_stage[bb][u4 * 4]);
        float4* dst = reinterpret_cast<float4*>(&h_buf[cur ^ 1][bb][rank * HS + u4 * 4]);
        *cluster.map_shared_rank(dst, c) = val;
      }
    }
    GRU_STAMP(4)
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    GRU_STAMP(5)
    if (fvalid) {
      const size_t off = out_base + (size_t)t * cl_cb(2 * H);
      float v = hnew_keep;
      if (a.add) v += __bfloat162float(a.add[off]);
      a.out[off] = __float2bfloat16(v * a.scale);
    }
    gxr = nxr, gxz = nxz, gxn = nxn;
    GRU_STAMP(6)
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
    GRU_STAMP(7)
  }
}

static int launch_gru_tc(const GruArgs& a, cudaStream_t st) {
  // clips per cluster: up to 8 cost the same MMAs; use as few clusters as the batch allows
  int bg = a.batch < 8 ? a.batch : 8;
  const int clusters = 2 * ceil_div(a.batch, bg);
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
  cudaError_t e = cudaLaunchKernelEx(&cfg, gru_cluster_tc_kernel, a, bg);
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
