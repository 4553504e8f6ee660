// HBM-bound passes: instance-norm statistics/apply, activations, reflect halos, the
// blur/bilinear resampling stencils and the modulated-conv side reductions.
// All kernels use channel-vector (8-wide, 128-bit for bf16) NHWC access when the
// tensors allow it and fall back to scalar access for C==1 images.
#include <cstdlib>
#include <type_traits>

#include "common.cuh"

namespace otm {

thread_local char g_err[512] = {0};
std::atomic<int64_t> g_launches{0};

static inline int ew_grid(long long total, int threads) {
  long long blocks = (total + threads - 1) / threads;
  long long cap = (long long)num_sms() * 16;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

// ---------------------------------------------------------------------------
// generic element-wise driver: F::operator()(n,h,w,c0)
// A CTA owns `rows` consecutive (n,h) image rows; a thread owns one (w, channel-vector) column
// of them.  All index arithmetic is 32-bit, (n,h) is CTA-uniform, and the (w,cv) split is a
// shift when C/V is a power of two -- the first version spent ~390 instructions per 16-byte
// vector on 64-bit div/mod and reached only 0.4-1.4 TB/s (profiles/r1_ew_kernels.md).
// ---------------------------------------------------------------------------
constexpr int PF_ROWS = 4;   // element-wise passes: prefetch this many image rows ahead
constexpr int PF_ITERS = 4;  // reductions: prefetch this many loop iterations ahead

template <int V, typename F>
__global__ void __launch_bounds__(256) ew_kernel(F f, int N, int H, int W, int C, int cv_shift,
                                                 int rows) {
  const int CV = C / V;
  const int WC = W * CV;
  const int nrows = N * H;
  const int r0 = blockIdx.x * rows;
  const int r1 = min(r0 + rows, nrows);
  for (int i = threadIdx.x; i < WC; i += 256) {
    int w, cv;
    if (cv_shift >= 0) { w = i >> cv_shift; cv = i & (CV - 1); }
    else { w = i / CV; cv = i - w * CV; }
    typename F::State st;
    int cur_n = -1;
    for (int r = r0; r < r1; ++r) {
      const int n = r / H, h = r - n * H;
      if (r + PF_ROWS < r1) {
        const int rp = r + PF_ROWS;
        const int np = rp / H;
        f.prefetch(np, rp - np * H, w, cv * V);
      }
      if (n != cur_n) { f.prepare(n, cv * V, st); cur_n = n; }  // per-(n, channel) constants
      f(n, h, w, cv * V, st);
    }
  }
}

// Persistent form: the work is cut into items of 256 channel-vectors (one image row = nblk
// items) and every CTA owns a CONTIGUOUS range of items whose length differs by at most one
// between CTAs.  grid = 4 CTAs per SM, all co-resident (64 registers/thread): no partial last
// wave (the row-chunk form above ran 1.04-1.15 waves of 8 CTAs/SM-sized grids on 2-4 resident
// CTAs/SM), and the per-(n, channel) constants are re-read only when the image changes.
// resident CTAs per SM of a functor's persistent kernel (register budget 65536 / (256 * OCC));
// functors that hold a block of outputs per thread declare `static constexpr int OCC = 2`
#ifndef OTM_EW_OCC
#define OTM_EW_OCC 4
#endif
#ifndef OTM_RED_OCC
#define OTM_RED_OCC 4
#endif
template <typename F, typename = void>
struct ew_occ { static constexpr int value = OTM_EW_OCC; };
template <typename F>
struct ew_occ<F, std::void_t<decltype(F::OCC)>> { static constexpr int value = F::OCC; };

template <int V, typename F>
__global__ void __launch_bounds__(256, ew_occ<F>::value)
ew_persist_kernel(F f, int N, int H, int W, int C, int cv_shift, int nblk, int items, int chunk) {
  const int CV = C / V;
  const int WC = W * CV;
  const int nrows = N * H;
  typename F::State st;
  int cur_n = -1, cur_cv = -1;
  if (chunk > 0) {
    // block-cyclic order: CTA b takes chunks b, b + G, ... of `chunk` consecutive items, so at any
    // time the grid streams through one contiguous window of the tensors
    for (int c0 = blockIdx.x * chunk; c0 < items; c0 += gridDim.x * chunk) {
      const int c1 = min(c0 + chunk, items);
      for (int it = c0; it < c1; ++it) {
        const int r = it / nblk, cb = it - r * nblk;
        const int n = r / H, h = r - n * H;
        const int i = cb * 256 + threadIdx.x;
        if (i < WC) {
          int w, cv;
          if (cv_shift >= 0) { w = i >> cv_shift; cv = i & (CV - 1); }
          else { w = i / CV; cv = i - w * CV; }
          if (n != cur_n || cv != cur_cv) { f.prepare(n, cv * V, st); cur_n = n; cur_cv = cv; }
          f(n, h, w, cv * V, st);
        }
      }
    }
    return;
  }
  const int t0 = (int)((long long)items * blockIdx.x / gridDim.x);
  const int t1 = (int)((long long)items * (blockIdx.x + 1) / gridDim.x);
  if (t0 >= t1) return;
  int r = t0 / nblk, cb = t0 - r * nblk;
  int n = r / H, h = r - n * H;
  for (int it = t0; it < t1; ++it) {
    const int i = cb * 256 + threadIdx.x;
    if (i < WC) {
      int w, cv;
      if (cv_shift >= 0) { w = i >> cv_shift; cv = i & (CV - 1); }
      else { w = i / CV; cv = i - w * CV; }
      if (r + PF_ROWS < nrows) {
        int hp = h + PF_ROWS, np = n;
        while (hp >= H) { hp -= H; ++np; }
        f.prefetch(np, hp, w, cv * V);
      }
      if (n != cur_n || cv != cur_cv) { f.prepare(n, cv * V, st); cur_n = n; cur_cv = cv; }
      f(n, h, w, cv * V, st);
    }
    if (++cb == nblk) {
      cb = 0; ++r;
      if (++h == H) { h = 0; ++n; }
    }
  }
}

template <int V, typename F>
static int launch_ew(F f, int N, int H, int W, int C, cudaStream_t st) {
  const long long total = (long long)N * H * W * (C / V);
  if (total == 0) return OTM_OK;
  const int CV = C / V;
  int cv_shift = -1;
  if ((CV & (CV - 1)) == 0) { cv_shift = 0; while ((1 << cv_shift) < CV) ++cv_shift; }
  const int nrows = N * H;
  constexpr int mode = 1;  // persistent CTAs (the one-shot grid below serves item counts >= 2^30)
  if (mode == 1) {
    const int nblk = (W * CV + 255) / 256;
    const long long items = (long long)nrows * nblk;
    if (items < (1ll << 30)) {
      long long grid = (long long)num_sms() * ew_occ<F>::value;
      if (grid > items) grid = items;
      constexpr int chunk = 0;
      ew_persist_kernel<V, F><<<(int)grid, 256, 0, st>>>(f, N, H, W, C, cv_shift, nblk, (int)items,
                                                         chunk);
      OTM_LAUNCH_CHECK();
      return OTM_OK;
    }
  }
  // enough CTAs for ~8 per SM; a bounded number of rows per CTA
  constexpr int rows_cap = 8;
  constexpr int ctas_per_sm = 8;
  int rows = nrows / (num_sms() * ctas_per_sm);
  if (rows < 1) rows = 1;
  if (rows > rows_cap) rows = rows_cap;
  const int grid = (nrows + rows - 1) / rows;
  ew_kernel<V, F><<<grid, 256, 0, st>>>(f, N, H, W, C, cv_shift, rows);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

// ---------------------------------------------------------------------------
// generic per-(n,c) reduction driver.
//   F::NQ quantities; F::operator()(n,h,w,c0, acc[NQ][V]) accumulates (and may write
//   element-wise outputs); F::out_index(n,c,q) gives the atomicAdd target.
// grid = (pixel chunks, 1, N), 256 threads = lanes (channel vectors) x rows (pixels).
// ---------------------------------------------------------------------------
template <int V, typename F>
__global__ void __launch_bounds__(256, OTM_RED_OCC) nc_reduce_kernel(F f, int H, int W, int C, int lanes,
                                                        int pix_per_cta, float* out) {
  constexpr int NQ = F::NQ;
  __shared__ float red[256 * NQ * V];
  const int CV = C / V;
  const int rows = 256 / lanes;
  const int lane = threadIdx.x % lanes;
  const int row = threadIdx.x / lanes;
  const int n = blockIdx.z;
  const int HW = H * W;
  const int p0 = blockIdx.x * pix_per_cta;
  const int p1 = min(HW, p0 + pix_per_cta);
  for (int cv0 = 0; cv0 < CV; cv0 += lanes) {
    const int cv = cv0 + lane;
    float acc[NQ][V];
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int i = 0; i < V; ++i) acc[q][i] = 0.f;
    if (cv < CV) {
      typename F::State st;
      f.prepare(n, cv * V, st);
      for (int p = p0 + row; p < p1; p += rows) {
        int h = p / W, w = p - h * W;
        const int pp = p + PF_ITERS * rows;
        if (pp < p1) {
          const int hp = pp / W;
          f.prefetch(n, hp, pp - hp * W, cv * V);
        }
        f(n, h, w, cv * V, acc, st);
      }
    }
#pragma unroll
    for (int q = 0; q < NQ; ++q)
#pragma unroll
      for (int i = 0; i < V; ++i) red[(q * V + i) * 256 + threadIdx.x] = acc[q][i];
    __syncthreads();
    // tree over rows (rows is a power of two)
    for (int s = rows / 2; s > 0; s >>= 1) {
      if (row < s) {
#pragma unroll
        for (int q = 0; q < NQ * V; ++q)
          red[q * 256 + threadIdx.x] += red[q * 256 + threadIdx.x + s * lanes];
      }
      __syncthreads();
    }
    if (row == 0 && cv < CV) {
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int i = 0; i < V; ++i)
          atomicAdd(out + f.out_index(n, cv * V + i, q), red[(q * V + i) * 256 + threadIdx.x]);
    }
    __syncthreads();
  }
}

template <int V, typename F>
static int launch_nc_reduce(F f, int N, int H, int W, int C, float* out, cudaStream_t st) {
  if (N == 0 || H * W == 0) return OTM_OK;
  int CV = C / V;
  int lanes = 1;
  while (lanes * 2 <= CV && lanes * 2 <= 256) lanes *= 2;
  int HW = H * W;
  int rows = 256 / lanes;
  // all CTAs co-resident (4 per SM at 64 registers): chunks * N <= 4 * SMs, so there is no
  // partial second wave; at least `rows*4` pixels per CTA
  constexpr int mode = 1;  // persistent CTAs (the one-shot grid below serves item counts >= 2^30)
  int want_chunks = mode == 1 ? (num_sms() * OTM_RED_OCC) / N : (num_sms() * 4 + N - 1) / N;
  if (want_chunks < 1) want_chunks = 1;
  int pix = (HW + want_chunks - 1) / want_chunks;
  int min_pix = rows * 4;
  if (pix < min_pix) pix = min_pix;
  int chunks = (HW + pix - 1) / pix;
  dim3 grid(chunks, 1, N);
  nc_reduce_kernel<V, F><<<grid, 256, 0, st>>>(f, H, W, C, lanes, pix, out);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

// ---------------------------------------------------------------------------
// instance-norm statistics
// ---------------------------------------------------------------------------
// Sums are accumulated about a per-(n,c) shift k = x[n,0,0,c] so that the one-pass variance
// E[(x-k)^2] - E[x-k]^2 does not cancel catastrophically when |mean| >> std (the parity mode
// needs ~1e-7 relative statistics: a 4e-6 error flips LeakyReLU masks in the 2x4-pixel layers).
template <typename T, int V>
struct StatsF {
  __device__ void prefetch(int, int, int, int) const {}  // pure read reduction: measured slower with prefetch
  static constexpr int NQ = 2;
  View x;
  int C;
  struct State { float k[V]; };
  __device__ void prepare(int n, int c, State& st) const { load_vec<T, V>(vptr<T>(x, n, 0, 0, c), st.k); }
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[2][V], const State& st) const {
    float v[V];
    load_vec<T, V>(vptr<T>(x, n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float d = v[i] - st.k[i];
      acc[0][i] += d;
      acc[1][i] += d * d;
    }
  }
  __device__ int out_index(int n, int c, int q) const { return (n * C + c) * 2 + q; }
};

template <typename T>
__global__ void stats_finalize_kernel(const float* ws, float* stats, View x, int count, float inv_n,
                                      float eps) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  int n = i / x.c, c = i - n * x.c;
  float k = to_f(*vptr<T>(x, n, 0, 0, c));
  float s = ws[2 * i], ss = ws[2 * i + 1];
  float dm = s * inv_n;
  float var = fmaxf(ss * inv_n - dm * dm, 0.f);
  stats[2 * i] = k + dm;
  stats[2 * i + 1] = rsqrtf(var + eps);
}

// (mean, rstd) from plain sums accumulated by a conv epilogue (otm_conv_fwd_args.stat_sums)
__global__ void stats_from_sums_kernel(const float* __restrict__ sums, float* __restrict__ stats,
                                       int count, float inv_n, float eps) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  const float m = sums[2 * i] * inv_n;
  const float var = fmaxf(sums[2 * i + 1] * inv_n - m * m, 0.f);
  stats[2 * i] = m;
  stats[2 * i + 1] = rsqrtf(var + eps);
}

// ---------------------------------------------------------------------------
// norm + act (+ residual) (+ reflect halo)
// ---------------------------------------------------------------------------
template <typename T, int V>
struct NormActF {
  __device__ void prefetch(int n, int h, int w, int c) const {
    prefetch_l2<T>(x, n, h, w, c);
    prefetch_l2<T>(res, n, h, w, c);
  }
  View x, res, y;
  const float* stats;
  int act, halo, C;
  struct State { float mean[V], rstd[V]; };
  __device__ void prepare(int n, int c, State& st) const {
    if (stats) {
      const float* p = stats + ((long long)n * C + c) * 2;
#pragma unroll
      for (int i = 0; i < V; ++i) { st.mean[i] = p[2 * i]; st.rstd[i] = p[2 * i + 1]; }
    }
  }
  __device__ void operator()(int n, int h, int w, int c, const State& st) const {
    float v[V];
    load_vec<T, V>(vptr<T>(x, n, h, w, c), v);
    if (stats) {
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] = (v[i] - st.mean[i]) * st.rstd[i];
    }
    act_fwd_vec<V>(v, act);
    if (res.ptr) {
      float r[V];
      load_vec<T, V>(vptr<T>(res, n, h, w, c), r);
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] += r[i];
    }
    store_halo<T, V>(y, halo, n, h, w, c, v);
  }
};

template <int MAXC>
__device__ __forceinline__ int down_bwd_taps(int i, int n_in, int n_out, float scale,
                                             int (&js)[MAXC], float (&wj)[MAXC]);

// backward: shared recompute of (ga, gn, pre).  DOWN is a compile-time switch so the common
// instances do not pay the register cost of the stencil-transpose path.
template <typename T, int V, bool DOWN>
struct NormActBwdBase {
  __device__ void prefetch(int n, int h, int w, int c) const {
    if constexpr (!DOWN) prefetch_l2<T>(g, n, h, w, c);
    prefetch_l2<T>(g2, n, h, w, c);
    prefetch_l2<T>(x, n, h, w, c);
  }
  View g, g2, x;
  const float* stats;
  int act, g_halo, C;
  int g_down;       // 1: g is the gradient of DownSample(act(norm(x))) at half resolution and the
  float sch, scw;   //    transposed blur+bilinear stencil is applied on load (no ga tensor)
  struct State { float mean[V], rstd[V], m1[V], m2[V]; };
  __device__ void prepare_stats(int n, int c, State& st) const {
    if (stats) {
      const float* p = stats + ((long long)n * C + c) * 2;
#pragma unroll
      for (int i = 0; i < V; ++i) { st.mean[i] = p[2 * i]; st.rstd[i] = p[2 * i + 1]; }
    }
  }
  __device__ void compute(int n, int h, int w, int c, float (&ga)[V], float (&gn)[V],
                          float (&pre)[V], const State& st) const {
    if constexpr (DOWN) {
      int jh[8], jw[8];
      float wh[8], ww[8];
      const int nh = down_bwd_taps<8>(h, x.h, g.h, sch, jh, wh);
      const int nw = down_bwd_taps<8>(w, x.w, g.w, scw, jw, ww);
#pragma unroll
      for (int i = 0; i < V; ++i) ga[i] = 0.f;
      for (int a = 0; a < nh; ++a)
        for (int b = 0; b < nw; ++b) {
          float t[V];
          load_fold<T, V>(g, g_halo, n, jh[a], jw[b], c, t);
          const float wgt = wh[a] * ww[b];
#pragma unroll
          for (int i = 0; i < V; ++i) ga[i] += wgt * t[i];
        }
    } else {
      load_fold<T, V>(g, g_halo, n, h, w, c, ga);
    }
    if (g2.ptr) {
      float t[V];
      load_vec<T, V>(vptr<T>(g2, n, h, w, c), t);
#pragma unroll
      for (int i = 0; i < V; ++i) ga[i] += t[i];
    }
    if (x.ptr) {
      load_vec<T, V>(vptr<T>(x, n, h, w, c), pre);
    } else {
#pragma unroll
      for (int i = 0; i < V; ++i) pre[i] = 0.f;
    }
    if (stats) {
#pragma unroll
      for (int i = 0; i < V; ++i) pre[i] = (pre[i] - st.mean[i]) * st.rstd[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) gn[i] = ga[i];
    act_bwd_vec<V>(gn, pre, act);
  }
};

template <typename T, int V, bool DOWN>
struct NormActBwdReduceF : NormActBwdBase<T, V, DOWN> {
  static constexpr int NQ = 2;
  using State = typename NormActBwdBase<T, V, DOWN>::State;
  __device__ void prepare(int n, int c, State& st) const { this->prepare_stats(n, c, st); }
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[2][V], const State& st) const {
    float ga[V], gn[V], pre[V];
    this->compute(n, h, w, c, ga, gn, pre, st);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      acc[0][i] += gn[i];
      acc[1][i] += gn[i] * pre[i];
    }
  }
  __device__ int out_index(int n, int c, int q) const { return (n * this->C + c) * 2 + q; }
};

template <typename T, int V, bool DOWN>
struct NormActBwdApplyF : NormActBwdBase<T, V, DOWN> {
  View gx, gres;
  const float* sums;
  float inv_hw;
  using State = typename NormActBwdBase<T, V, DOWN>::State;
  __device__ void prepare(int n, int c, State& st) const {
    this->prepare_stats(n, c, st);
    if (this->stats) {
      const float* sm = sums + ((long long)n * this->C + c) * 2;
#pragma unroll
      for (int i = 0; i < V; ++i) { st.m1[i] = sm[2 * i] * inv_hw; st.m2[i] = sm[2 * i + 1] * inv_hw; }
    }
  }
  __device__ void operator()(int n, int h, int w, int c, const State& st) const {
    float ga[V], gn[V], pre[V];
    this->compute(n, h, w, c, ga, gn, pre, st);
    if (gres.ptr) store_vec<T, V>(vptr_mut<T>(gres, n, h, w, c), ga);
    if (this->stats) {
#pragma unroll
      for (int i = 0; i < V; ++i) gn[i] = st.rstd[i] * (gn[i] - st.m1[i] - pre[i] * st.m2[i]);
    }
    store_vec<T, V>(vptr_mut<T>(gx, n, h, w, c), gn);
  }
};

// ---------------------------------------------------------------------------
// Second derivative of InstanceNorm (+ LeakyReLU / ReLU): the R1 gradient penalty differentiates
// a BACKWARD pass (BASELINE config 5).  First backward (otm_norm_act_bwd), per (n,c) plane of N
// pixels, yh = (x - mean) * r:
//     gy = g * act'(yh) ;  gx = B(gy) = r * (gy - mean(gy) - yh * mean(gy * yh))
// Given gg = dL/d(gx):   B is linear and self-adjoint in gy, so
//     dL/d(g)  = act'(yh) * B(gg)
//     dL/d(x)_j = -r^2 * [ b (gg_j - mean(gg)) + a (gy_j - mean(gy))
//                          + yh_j (mean(gg * gy) - mean(gg) mean(gy) - 3 a b) ]
//   with a = mean(gg * yh), b = mean(gy * yh)  (act' is piecewise constant: no term from it).
// Pass 1 reduces the five means, pass 2 writes both outputs.
// ---------------------------------------------------------------------------
template <typename T, int V>
struct Norm2Base {
  View g, gg, x;
  const float* stats;
  int act, C;
  struct State { float mean[V], rstd[V], mg[V], mgg[V], a[V], b[V], t[V]; };
  __device__ void prefetch(int, int, int, int) const {}
  __device__ void prepare_stats(int n, int c, State& st) const {
    const float* p = stats + ((long long)n * C + c) * 2;
#pragma unroll
    for (int i = 0; i < V; ++i) { st.mean[i] = p[2 * i]; st.rstd[i] = p[2 * i + 1]; }
  }
  __device__ void load(int n, int h, int w, int c, const State& st, float (&gy)[V], float (&ggv)[V],
                       float (&yh)[V], float (&mask)[V]) const {
    load_vec<T, V>(vptr<T>(g, n, h, w, c), gy);
    load_vec<T, V>(vptr<T>(gg, n, h, w, c), ggv);
    load_vec<T, V>(vptr<T>(x, n, h, w, c), yh);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      yh[i] = (yh[i] - st.mean[i]) * st.rstd[i];
      mask[i] = act == OTM_ACT_NONE ? 1.f : (yh[i] > 0.f ? 1.f : (act == OTM_ACT_LRELU ? 0.2f : 0.f));
      gy[i] *= mask[i];
    }
  }
};

template <typename T, int V>
struct Norm2ReduceF : Norm2Base<T, V> {
  static constexpr int NQ = 5;
  using State = typename Norm2Base<T, V>::State;
  __device__ void prepare(int n, int c, State& st) const { this->prepare_stats(n, c, st); }
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[5][V], const State& st) const {
    float gy[V], ggv[V], yh[V], mask[V];
    this->load(n, h, w, c, st, gy, ggv, yh, mask);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      acc[0][i] += gy[i];
      acc[1][i] += ggv[i];
      acc[2][i] += ggv[i] * yh[i];
      acc[3][i] += gy[i] * yh[i];
      acc[4][i] += ggv[i] * gy[i];
    }
  }
  __device__ int out_index(int n, int c, int q) const { return (n * this->C + c) * 5 + q; }
};

template <typename T, int V>
struct Norm2ApplyF : Norm2Base<T, V> {
  View dg, dx;
  const float* sums;
  float inv_hw;
  using State = typename Norm2Base<T, V>::State;
  __device__ void prepare(int n, int c, State& st) const {
    this->prepare_stats(n, c, st);
    const float* sm = sums + ((long long)n * this->C + c) * 5;
#pragma unroll
    for (int i = 0; i < V; ++i) {
      st.mg[i] = sm[5 * i] * inv_hw;
      st.mgg[i] = sm[5 * i + 1] * inv_hw;
      st.a[i] = sm[5 * i + 2] * inv_hw;
      st.b[i] = sm[5 * i + 3] * inv_hw;
      st.t[i] = sm[5 * i + 4] * inv_hw - st.mgg[i] * st.mg[i] - 3.f * st.a[i] * st.b[i];
    }
  }
  __device__ void operator()(int n, int h, int w, int c, const State& st) const {
    float gy[V], ggv[V], yh[V], mask[V], o1[V], o2[V];
    this->load(n, h, w, c, st, gy, ggv, yh, mask);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      o1[i] = mask[i] * st.rstd[i] * (ggv[i] - st.mgg[i] - yh[i] * st.a[i]);
      o2[i] = -st.rstd[i] * st.rstd[i] *
              (st.b[i] * (ggv[i] - st.mgg[i]) + st.a[i] * (gy[i] - st.mg[i]) + yh[i] * st.t[i]);
    }
    if (dg.ptr) store_vec<T, V>(vptr_mut<T>(dg, n, h, w, c), o1);
    if (dx.ptr) store_vec<T, V>(vptr_mut<T>(dx, n, h, w, c), o2);
  }
};

// ---------------------------------------------------------------------------
// resampling: shared 1-D tap generators
// ---------------------------------------------------------------------------
// DownSample = blur then bilinear to n_out = n_in/2 (align_corners=False, scale n_in/n_out).
// For output index j: up to 4 source positions base-1+k (clamped) with merged weights.
__device__ __forceinline__ void down_taps(int j, int n_in, float scale, int (&pos)[4],
                                          float (&wt)[4]) {
  float src = fmaxf((j + 0.5f) * scale - 0.5f, 0.f);
  int i0 = min((int)floorf(src), n_in - 1);
  int i1 = min(i0 + 1, n_in - 1);
  float lam = src - (float)i0;
  const float wa = 1.f - lam;
  const float l0 = (i1 == i0) ? lam : 0.f;  // degenerate edge: both taps on the same row
  const float l1 = (i1 == i0) ? 0.f : lam;
  wt[0] = 0.25f * (wa + l0);
  wt[1] = 0.5f * (wa + l0) + 0.25f * l1;
  wt[2] = 0.25f * (wa + l0) + 0.5f * l1;
  wt[3] = 0.25f * l1;
#pragma unroll
  for (int k = 0; k < 4; ++k) pos[k] = min(max(i0 - 1 + k, 0), n_in - 1);
}

// UpSample = bilinear x2 then blur.  For output j in [0, 2n): positions base+k, k<4.
// Interior outputs have the closed form  even j=2m: (.3125,.625,.0625) on x[m-1..m+1],
// odd j=2m+1: (.0625,.625,.3125); the clamped borders take the general path.
__device__ __forceinline__ void up_taps(int j, int n_in, int (&pos)[4], float (&wt)[4]) {
  const int n_out = 2 * n_in;
  if (j >= 2 && j < n_out - 2) {
    const int m = j >> 1;
    pos[0] = m - 1; pos[1] = m; pos[2] = m + 1; pos[3] = m + 1;
    const bool odd = j & 1;
    wt[0] = odd ? 0.0625f : 0.3125f; wt[1] = 0.625f; wt[2] = odd ? 0.3125f : 0.0625f; wt[3] = 0.f;
    return;
  }
  wt[0] = wt[1] = wt[2] = wt[3] = 0.f;
  int jm = max(j - 1, 0);
  float srcm = fmaxf((jm + 0.5f) * 0.5f - 0.5f, 0.f);
  int base = (int)floorf(srcm);
#pragma unroll
  for (int d = -1; d <= 1; ++d) {
    int jj = min(max(j + d, 0), n_out - 1);
    float kd = (d == 0) ? 0.5f : 0.25f;
    float src = fmaxf((jj + 0.5f) * 0.5f - 0.5f, 0.f);
    int i0 = min((int)floorf(src), n_in - 1);
    int i1 = min(i0 + 1, n_in - 1);
    float lam = src - (float)i0;
    const int a = i0 - base, b = i1 - base;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      wt[k] += (k == a ? kd * (1.f - lam) : 0.f) + (k == b ? kd * lam : 0.f);
  }
#pragma unroll
  for (int k = 0; k < 4; ++k) pos[k] = min(base + k, n_in - 1);
}

template <typename T, int V>
struct DownF {
  __device__ void prefetch(int, int, int, int) const {}
  View x, y;
  const float* stats;
  int act, halo, C;
  float sch, scw;
  struct State { float mean[V], rstd[V]; };
  __device__ void prepare(int n, int c, State& st) const {
    if (stats) {
      const float* p = stats + ((long long)n * C + c) * 2;
#pragma unroll
      for (int i = 0; i < V; ++i) { st.mean[i] = p[2 * i]; st.rstd[i] = p[2 * i + 1]; }
    }
  }
  __device__ void operator()(int n, int ho, int wo, int c, const State& st) const {
    int ph[4], pw[4];
    float wh[4], ww[4];
    down_taps(ho, x.h, sch, ph, wh);
    down_taps(wo, x.w, scw, pw, ww);
    const float (&mean)[V] = st.mean;
    const float (&rstd)[V] = st.rstd;
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wh[a] == 0.f) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float wgt = wh[a] * ww[b];
        if (wgt == 0.f) continue;
        float v[V];
        load_vec<T, V>(vptr<T>(x, n, ph[a], pw[b], c), v);
        if (stats) {
#pragma unroll
          for (int i = 0; i < V; ++i) v[i] = (v[i] - mean[i]) * rstd[i];
        }
        act_fwd_vec<V>(v, act);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += wgt * v[i];
      }
    }
    store_halo<T, V>(y, halo, n, ho, wo, c, acc);
  }
};

// gather form of the transpose: for input index i, list (j, weight) of outputs touching it.
// Output j touches i iff floor(src_j) in [i-2, i+1] (positions i0-1..i0+2, clamped at the ends).
template <int MAXC>
__device__ __forceinline__ int down_bwd_taps(int i, int n_in, int n_out, float scale,
                                             int (&js)[MAXC], float (&wj)[MAXC]) {
  int cnt = 0;
  if (n_in == 2 * n_out && i >= 2 && i <= n_in - 3) {  // exact /2: taps (.125,.375,.375,.125)
    const int m = i >> 1;
    if (i & 1) { js[0] = m; wj[0] = 0.375f; js[1] = m + 1; wj[1] = 0.125f; }
    else { js[0] = m - 1; wj[0] = 0.125f; js[1] = m; wj[1] = 0.375f; }
    return 2;
  }
  const float inv = 1.f / scale;
  int lo = max(0, (int)floorf((i - 1.5f) * inv - 0.5f) - 1);
  int hi = min(n_out - 1, (int)ceilf((i + 2.5f) * inv - 0.5f) + 1);
  for (int j = lo; j <= hi; ++j) {
    int pos[4];
    float wt[4];
    down_taps(j, n_in, scale, pos, wt);
    float w = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (pos[k] == i) w += wt[k];
    if (w != 0.f && cnt < MAXC) { js[cnt] = j; wj[cnt] = w; ++cnt; }
  }
  return cnt;
}

template <int MAXC>
__device__ __forceinline__ int up_bwd_taps(int i, int n_in, int (&js)[MAXC], float (&wj)[MAXC]) {
  const int n_out = 2 * n_in;
  if (i >= 2 && i <= n_in - 3) {  // interior: outputs 2i-2 .. 2i+3
    const float w6[6] = {0.0625f, 0.3125f, 0.625f, 0.625f, 0.3125f, 0.0625f};
#pragma unroll
    for (int k = 0; k < 6; ++k) { js[k] = 2 * i - 2 + k; wj[k] = w6[k]; }
    return 6;
  }
  int cnt = 0;
  int lo = max(0, 2 * i - 4), hi = min(n_out - 1, 2 * i + 5);
  for (int j = lo; j <= hi; ++j) {
    int pos[4];
    float wt[4];
    up_taps(j, n_in, pos, wt);
    float w = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k)
      if (pos[k] == i) w += wt[k];
    if (w != 0.f && cnt < MAXC) { js[cnt] = j; wj[cnt] = w; ++cnt; }
  }
  return cnt;
}

template <typename T, int V>
struct DownBwdF {
  __device__ void prefetch(int, int, int, int) const {}
  View g, ga;
  int g_halo;
  float sch, scw;
  struct State {};
  __device__ void prepare(int, int, State&) const {}
  __device__ void operator()(int n, int h, int w, int c, const State&) const {
    int jh[8], jw[8];
    float wh[8], ww[8];
    int nh = down_bwd_taps<8>(h, ga.h, g.h, sch, jh, wh);
    int nw = down_bwd_taps<8>(w, ga.w, g.w, scw, jw, ww);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        float v[V];
        load_fold<T, V>(g, g_halo, n, jh[a], jw[b], c, v);
        float wgt = wh[a] * ww[b];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += wgt * v[i];
      }
    store_vec<T, V>(vptr_mut<T>(ga, n, h, w, c), acc);
  }
};

template <typename T, int V>
struct UpF {
  __device__ void prefetch(int, int, int, int) const {}
  View x, y;
  int halo;
  const float* scale;  // [n, C] or NULL
  int C;
  struct State { float sc[V]; };
  __device__ void prepare(int n, int c, State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) st.sc[i] = scale ? scale[(long long)n * C + c + i] : 1.f;
  }
  __device__ void operator()(int n, int ho, int wo, int c, const State& st) const {
    int ph[4], pw[4];
    float wh[4], ww[4];
    up_taps(ho, x.h, ph, wh);
    up_taps(wo, x.w, pw, ww);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      if (wh[a] == 0.f) continue;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        float wgt = wh[a] * ww[b];
        if (wgt == 0.f) continue;
        float v[V];
        load_vec<T, V>(vptr<T>(x, n, ph[a], pw[b], c), v);
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += wgt * v[i];
      }
    }
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] *= st.sc[i];
    store_halo<T, V>(y, halo, n, ho, wo, c, acc);
  }
};

template <typename T, int V>
struct UpBwdF {
  __device__ void prefetch(int, int, int, int) const {}
  View g, gx;
  int g_halo;
  const float* scale;  // [n, C] or NULL
  int C;
  struct State { float sc[V]; };
  __device__ void prepare(int n, int c, State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) st.sc[i] = scale ? scale[(long long)n * C + c + i] : 1.f;
  }
  __device__ void operator()(int n, int h, int w, int c, const State& st) const {
    int jh[10], jw[10];
    float wh[10], ww[10];
    int nh = up_bwd_taps<10>(h, gx.h, jh, wh);
    int nw = up_bwd_taps<10>(w, gx.w, jw, ww);
    float acc[V];
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] = 0.f;
    for (int a = 0; a < nh; ++a)
      for (int b = 0; b < nw; ++b) {
        float v[V];
        load_fold<T, V>(g, g_halo, n, jh[a], jw[b], c, v);
        float wgt = wh[a] * ww[b];
#pragma unroll
        for (int i = 0; i < V; ++i) acc[i] += wgt * v[i];
      }
#pragma unroll
    for (int i = 0; i < V; ++i) acc[i] *= st.sc[i];
    store_vec<T, V>(vptr_mut<T>(gx, n, h, w, c), acc);
  }
};

// Block forms of the exact-halving DownSample (n_in == 2 n_out: blur + bilinear collapses to the
// separable taps (.125,.375,.375,.125) on x[2j-1 .. 2j+2]).  One thread produces a 2x2 output
// block from a 6x6 input window (9 loads per output vector instead of 16; the norm + activation
// in front of the stencil is applied 36 instead of 64 times), resp. the gradients of a 2x2 INPUT
// block from the 3x3 output gradients that touch it (2.25 loads per vector instead of 4).
// Blocks whose window leaves the image, and all odd sizes, take the per-pixel functors above.
template <typename T, int V>
struct Down2x2F {
  static constexpr int OCC = 2;
  __device__ void prefetch(int, int, int, int) const {}
  View x, y;
  const float* stats;
  int act, halo, C;
  struct State { float mean[V], rstd[V]; };
  __device__ void prepare(int n, int c, State& st) const {
    if (stats) {
      const float* p = stats + ((long long)n * C + c) * 2;
#pragma unroll
      for (int i = 0; i < V; ++i) { st.mean[i] = p[2 * i]; st.rstd[i] = p[2 * i + 1]; }
    }
  }
  // (h2, w2) index 2x2 blocks of y
  __device__ void operator()(int n, int h2, int w2, int c, const State& st) const {
    const int ho = 2 * h2, wo = 2 * w2;
    const int r0 = 2 * ho - 1, c0 = 2 * wo - 1;  // window origin in x
    if (r0 >= 0 && r0 + 5 < x.h && c0 >= 0 && c0 + 5 < x.w && ho + 1 < y.h && wo + 1 < y.w) {
      const float k4[4] = {0.125f, 0.375f, 0.375f, 0.125f};
      float acc[2][2][V];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[a][b][i] = 0.f;
#pragma unroll
      for (int r = 0; r < 6; ++r) {
        float h0[V], h1[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { h0[i] = 0.f; h1[i] = 0.f; }
#pragma unroll
        for (int q = 0; q < 6; ++q) {
          float v[V];
          load_vec<T, V>(vptr<T>(x, n, r0 + r, c0 + q, c), v);
          if (stats) {
#pragma unroll
            for (int i = 0; i < V; ++i) v[i] = (v[i] - st.mean[i]) * st.rstd[i];
          }
          act_fwd_vec<V>(v, act);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            if (q < 4) h0[i] = fmaf(k4[q], v[i], h0[i]);
            if (q >= 2) h1[i] = fmaf(k4[q - 2], v[i], h1[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          if (r < 4) { acc[0][0][i] = fmaf(k4[r], h0[i], acc[0][0][i]); acc[0][1][i] = fmaf(k4[r], h1[i], acc[0][1][i]); }
          if (r >= 2) { acc[1][0][i] = fmaf(k4[r - 2], h0[i], acc[1][0][i]); acc[1][1][i] = fmaf(k4[r - 2], h1[i], acc[1][1][i]); }
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) store_halo<T, V>(y, halo, n, ho + a, wo + b, c, acc[a][b]);
      return;
    }
    DownF<T, V> f{x, y, stats, act, halo, C, (float)x.h / (float)y.h, (float)x.w / (float)y.w};
    typename DownF<T, V>::State fs;
#pragma unroll
    for (int i = 0; i < V; ++i) { fs.mean[i] = st.mean[i]; fs.rstd[i] = st.rstd[i]; }
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        if (ho + a < y.h && wo + b < y.w) f(n, ho + a, wo + b, c, fs);
  }
};

template <typename T, int V>
struct DownBwd2x2F {
  static constexpr int OCC = 3;
  __device__ void prefetch(int, int, int, int) const {}
  View g, ga;  // g: [n, H/2, W/2, c] (g_halo == 0 only), ga: [n, H, W, c]
  struct State {};
  __device__ void prepare(int, int, State&) const {}
  // (h2, w2) index 2x2 blocks of ga: rows 2 h2, 2 h2 + 1 are touched by output rows h2-1 .. h2+1
  __device__ void operator()(int n, int h2, int w2, int c, const State&) const {
    const int h = 2 * h2, w = 2 * w2;
    if (h >= 2 && h + 1 <= ga.h - 3 && w >= 2 && w + 1 <= ga.w - 3) {
      // input row 2a   <- outputs a-1 (.125), a (.375);  row 2a+1 <- outputs a (.375), a+1 (.125)
      float t[3][3][V];
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int q = 0; q < 3; ++q) load_vec<T, V>(vptr<T>(g, n, h2 - 1 + r, w2 - 1 + q, c), t[r][q]);
      float he[3][V], ho[3][V];  // horizontal pass: even / odd input column
#pragma unroll
      for (int r = 0; r < 3; ++r)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          he[r][i] = 0.125f * t[r][0][i] + 0.375f * t[r][1][i];
          ho[r][i] = 0.375f * t[r][1][i] + 0.125f * t[r][2][i];
        }
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.125f * he[0][i] + 0.375f * he[1][i];
      store_vec<T, V>(vptr_mut<T>(ga, n, h, w, c), o);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.125f * ho[0][i] + 0.375f * ho[1][i];
      store_vec<T, V>(vptr_mut<T>(ga, n, h, w + 1, c), o);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.375f * he[1][i] + 0.125f * he[2][i];
      store_vec<T, V>(vptr_mut<T>(ga, n, h + 1, w, c), o);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.375f * ho[1][i] + 0.125f * ho[2][i];
      store_vec<T, V>(vptr_mut<T>(ga, n, h + 1, w + 1, c), o);
      return;
    }
    DownBwdF<T, V> f{g, ga, 0, (float)ga.h / (float)g.h, (float)ga.w / (float)g.w};
    typename DownBwdF<T, V>::State fs;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        if (h + a < ga.h && w + b < ga.w) f(n, h + a, w + b, c, fs);
  }
};

// Block forms of the x2 up-sampling stencils.  The per-output functors above spend ~200
// instructions per 16-byte vector on tap generation and issue 9 (forward) / 36 (backward) loads
// per output vector; they ran at 0.8-1.2 TB/s.  Here one thread produces a 2x2 block:
//   forward : the four outputs of input pixel (h, w) from its 3x3 neighbourhood, separable,
//             constant weights (even j=2m: .3125,.625,.0625 on m-1..m+1; odd: mirrored);
//   backward: the input gradients of the 2x2 pixel block (2h..2h+1, 2w..2w+1) from the 8x8
//             output-gradient window (rows 4h-2 .. 4h+5), horizontal pass first.
// Blocks that touch the clamped image border take the general tap path of the functors above.
template <typename T, int V>
struct Up2x2F {
  static constexpr int OCC = 2;
  __device__ void prefetch(int, int, int, int) const {}
  View x, y;
  int halo;
  const float* scale;  // [n, C] or NULL
  int C;
  struct State { float sc[V]; };
  __device__ void prepare(int n, int c, State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) st.sc[i] = scale ? scale[(long long)n * C + c + i] : 1.f;
  }
  __device__ void put(int n, int h, int w, int c, float (&o)[V], const State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] *= st.sc[i];
    store_halo<T, V>(y, halo, n, h, w, c, o);
  }
  __device__ void operator()(int n, int h, int w, int c, const State& st) const {
    if (h >= 1 && h < x.h - 1 && w >= 1 && w < x.w - 1) {
      float he[3][V], ho[3][V];
#pragma unroll
      for (int r = 0; r < 3; ++r) {
        float a[V], b[V], d[V];
        load_vec<T, V>(vptr<T>(x, n, h - 1 + r, w - 1, c), a);
        load_vec<T, V>(vptr<T>(x, n, h - 1 + r, w, c), b);
        load_vec<T, V>(vptr<T>(x, n, h - 1 + r, w + 1, c), d);
#pragma unroll
        for (int i = 0; i < V; ++i) {
          he[r][i] = 0.3125f * a[i] + 0.625f * b[i] + 0.0625f * d[i];
          ho[r][i] = 0.0625f * a[i] + 0.625f * b[i] + 0.3125f * d[i];
        }
      }
      float o[V];
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.3125f * he[0][i] + 0.625f * he[1][i] + 0.0625f * he[2][i];
      put(n, 2 * h, 2 * w, c, o, st);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.3125f * ho[0][i] + 0.625f * ho[1][i] + 0.0625f * ho[2][i];
      put(n, 2 * h, 2 * w + 1, c, o, st);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.0625f * he[0][i] + 0.625f * he[1][i] + 0.3125f * he[2][i];
      put(n, 2 * h + 1, 2 * w, c, o, st);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] = 0.0625f * ho[0][i] + 0.625f * ho[1][i] + 0.3125f * ho[2][i];
      put(n, 2 * h + 1, 2 * w + 1, c, o, st);
      return;
    }
    // border block: same separable 3x3 -> 2x2 scheme with per-axis weights from the general
    // tap generator (window clamped into the image)
    const int hs = min(max(h - 1, 0), x.h - 3), ws = min(max(w - 1, 0), x.w - 3);
    float wr[2][3], wc[2][3];
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      int pos[4];
      float wt[4];
      up_taps(2 * h + a, x.h, pos, wt);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += (pos[k] == hs + t) ? wt[k] : 0.f;
        wr[a][t] = acc;
      }
      up_taps(2 * w + a, x.w, pos, wt);
#pragma unroll
      for (int t = 0; t < 3; ++t) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += (pos[k] == ws + t) ? wt[k] : 0.f;
        wc[a][t] = acc;
      }
    }
    float hv[2][3][V];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
      float a[V], b[V], d[V];
      load_vec<T, V>(vptr<T>(x, n, hs + r, ws, c), a);
      load_vec<T, V>(vptr<T>(x, n, hs + r, ws + 1, c), b);
      load_vec<T, V>(vptr<T>(x, n, hs + r, ws + 2, c), d);
#pragma unroll
      for (int q = 0; q < 2; ++q)
#pragma unroll
        for (int i = 0; i < V; ++i) hv[q][r][i] = wc[q][0] * a[i] + wc[q][1] * b[i] + wc[q][2] * d[i];
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        float o[V];
#pragma unroll
        for (int i = 0; i < V; ++i)
          o[i] = wr[a][0] * hv[q][0][i] + wr[a][1] * hv[q][1][i] + wr[a][2] * hv[q][2][i];
        put(n, 2 * h + a, 2 * w + q, c, o, st);
      }
  }
};

template <typename T, int V>
struct UpBwd2x2F {
  static constexpr int OCC = 2;
  __device__ void prefetch(int, int, int, int) const {}
  View g, gx;  // g: [n, 2H, 2W, c] output gradient (g_halo == 0 only), gx: [n, H, W, c]
  const float* scale;  // [n, C] or NULL
  int C;
  struct State { float sc[V]; };
  __device__ void prepare(int n, int c, State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) st.sc[i] = scale ? scale[(long long)n * C + c + i] : 1.f;
  }
  __device__ void put(int n, int h, int w, int c, float (&o)[V], const State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) o[i] *= st.sc[i];
    store_vec<T, V>(vptr_mut<T>(gx, n, h, w, c), o);
  }
  // (h2, w2) index 2x2 blocks of gx
  __device__ void operator()(int n, int h2, int w2, int c, const State& st) const {
    const int h = 2 * h2, w = 2 * w2;
    if (h >= 2 && h + 1 <= gx.h - 3 && w >= 2 && w + 1 <= gx.w - 3) {
      // horizontal pass: window columns 2w-2 .. 2w+5; input col w uses 0..5, col w+1 uses 2..7
      const float k6[6] = {0.0625f, 0.3125f, 0.625f, 0.625f, 0.3125f, 0.0625f};
      float acc[2][2][V];
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b)
#pragma unroll
          for (int i = 0; i < V; ++i) acc[a][b][i] = 0.f;
#pragma unroll
      for (int r = 0; r < 8; ++r) {
        float h0[V], h1[V];
#pragma unroll
        for (int i = 0; i < V; ++i) { h0[i] = 0.f; h1[i] = 0.f; }
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          float v[V];
          load_vec<T, V>(vptr<T>(g, n, 2 * h - 2 + r, 2 * w - 2 + q, c), v);
#pragma unroll
          for (int i = 0; i < V; ++i) {
            if (q < 6) h0[i] = fmaf(k6[q], v[i], h0[i]);
            if (q >= 2) h1[i] = fmaf(k6[q - 2], v[i], h1[i]);
          }
        }
#pragma unroll
        for (int i = 0; i < V; ++i) {
          if (r < 6) { acc[0][0][i] = fmaf(k6[r], h0[i], acc[0][0][i]); acc[0][1][i] = fmaf(k6[r], h1[i], acc[0][1][i]); }
          if (r >= 2) { acc[1][0][i] = fmaf(k6[r - 2], h0[i], acc[1][0][i]); acc[1][1][i] = fmaf(k6[r - 2], h1[i], acc[1][1][i]); }
        }
      }
#pragma unroll
      for (int a = 0; a < 2; ++a)
#pragma unroll
        for (int b = 0; b < 2; ++b) put(n, h + a, w + b, c, acc[a][b], st);
      return;
    }
    // border block: the same 8x8 -> 2x2 separable scheme with per-axis weights derived from
    // the forward tap generator (window clamped into the image; an odd trailing row/column of gx
    // gets zero weights and is not stored)
    const int hs = min(max(2 * h - 2, 0), g.h - 8), ws = min(max(2 * w - 2, 0), g.w - 8);
    float wr[2][8], wc[2][8];
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      int pos[4];
      float wt[4];
      up_taps(hs + r, gx.h, pos, wt);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += (pos[k] == h + a) ? wt[k] : 0.f;
        wr[a][r] = acc;
      }
      up_taps(ws + r, gx.w, pos, wt);
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        float acc = 0.f;
#pragma unroll
        for (int k = 0; k < 4; ++k) acc += (pos[k] == w + a) ? wt[k] : 0.f;
        wc[a][r] = acc;
      }
    }
    float acc[2][2][V];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
#pragma unroll
        for (int i = 0; i < V; ++i) acc[a][b][i] = 0.f;
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      float h0[V], h1[V];
#pragma unroll
      for (int i = 0; i < V; ++i) { h0[i] = 0.f; h1[i] = 0.f; }
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float v[V];
        load_vec<T, V>(vptr<T>(g, n, hs + r, ws + q, c), v);
#pragma unroll
        for (int i = 0; i < V; ++i) { h0[i] = fmaf(wc[0][q], v[i], h0[i]); h1[i] = fmaf(wc[1][q], v[i], h1[i]); }
      }
#pragma unroll
      for (int i = 0; i < V; ++i) {
        acc[0][0][i] = fmaf(wr[0][r], h0[i], acc[0][0][i]); acc[0][1][i] = fmaf(wr[0][r], h1[i], acc[0][1][i]);
        acc[1][0][i] = fmaf(wr[1][r], h0[i], acc[1][0][i]); acc[1][1][i] = fmaf(wr[1][r], h1[i], acc[1][1][i]);
      }
    }
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 2; ++b)
        if (h + a < gx.h && w + b < gx.w) put(n, h + a, w + b, c, acc[a][b], st);
  }
};

// ---------------------------------------------------------------------------
// modulated-conv side passes
// ---------------------------------------------------------------------------
template <typename T, int V>
struct ModOutF {
  __device__ void prefetch(int n, int h, int w, int c) const {
    prefetch_l2<T>(g, n, h, w, c);
    prefetch_l2<T>(g2, n, h, w, c);
    prefetch_l2<T>(out, n, h, w, c);
    prefetch_l2<T>(res, n, h, w, c);
  }
  static constexpr int NQ = 1;
  View g, g2, out, res, gy;
  int g_halo, act, C;
  const float* gys;  // [n, C] scale of the STORED gy (NULL = 1)
  struct State { float gs[V]; };
  __device__ void prepare(int n, int c, State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) st.gs[i] = gys ? gys[(long long)n * C + c + i] : 1.f;
  }
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[1][V], const State& st) const {
    float ga[V], o[V];
    load_fold<T, V>(g, g_halo, n, h, w, c, ga);
    if (g2.ptr) {
      float t[V];
      load_vec<T, V>(vptr<T>(g2, n, h, w, c), t);
#pragma unroll
      for (int i = 0; i < V; ++i) ga[i] += t[i];
    }
    load_vec<T, V>(vptr<T>(out, n, h, w, c), o);
    if (act == OTM_ACT_RELU) {
#pragma unroll
      for (int i = 0; i < V; ++i) ga[i] = o[i] > 0.f ? ga[i] : 0.f;
    }
    if (res.ptr) {
      float r[V];
      load_vec<T, V>(vptr<T>(res, n, h, w, c), r);
#pragma unroll
      for (int i = 0; i < V; ++i) o[i] -= r[i];
    }
#pragma unroll
    for (int i = 0; i < V; ++i) { acc[0][i] += ga[i] * o[i]; ga[i] *= st.gs[i]; }
    if (gy.ptr) store_vec<T, V>(vptr_mut<T>(gy, n, h, w, c), ga);
  }
  __device__ int out_index(int n, int c, int) const { return n * C + c; }
};

template <typename T, int V>
struct ModInF {
  __device__ void prefetch(int n, int h, int w, int c) const {
    prefetch_l2<T>(g, n, h, w, c);
    prefetch_l2<T>(x, n, h, w, c);
    prefetch_l2<T>(gadd, n, h, w, c);
  }
  static constexpr int NQ = 1;
  View g, x, gadd, gx;
  const float* s;
  int g_halo, C, relu_mask;
  const float* gxs;  // [n, C] scale of the stored gx (NULL = 1)
  struct State { float s[V], gs[V]; };
  __device__ void prepare(int n, int c, State& st) const {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      st.s[i] = s[(long long)n * C + c + i];
      st.gs[i] = gxs ? gxs[(long long)n * C + c + i] : 1.f;
    }
  }
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[1][V], const State& st) const {
    float gt[V], xv[V];
    load_fold<T, V>(g, g_halo, n, h, w, c, gt);
    load_vec<T, V>(vptr<T>(x, n, h, w, c), xv);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      acc[0][i] += gt[i] * xv[i];
      gt[i] *= st.s[i];
    }
    if (gadd.ptr) {
      float t[V];
      load_vec<T, V>(vptr<T>(gadd, n, h, w, c), t);
#pragma unroll
      for (int i = 0; i < V; ++i) gt[i] += t[i];
    }
    if (!gx.ptr) return;
    if (relu_mask) {
#pragma unroll
      for (int i = 0; i < V; ++i) gt[i] = xv[i] != 0.f ? gt[i] : 0.f;
    }
#pragma unroll
    for (int i = 0; i < V; ++i) gt[i] *= st.gs[i];
    store_vec<T, V>(vptr_mut<T>(gx, n, h, w, c), gt);
  }
  __device__ int out_index(int n, int c, int) const { return n * C + c; }
};

template <typename T, int V>
struct ChannelSumF {
  __device__ void prefetch(int, int, int, int) const {}
  static constexpr int NQ = 1;
  View g;
  float scale;
  int per_sample, C;
  struct State {};
  __device__ void prepare(int, int, State&) const {}
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[1][V], const State&) const {
    float v[V];
    load_vec<T, V>(vptr<T>(g, n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) acc[0][i] += v[i] * scale;
  }
  __device__ int out_index(int n, int c, int) const { return per_sample ? n * C + c : c; }
};

template <typename T, int V>
struct AvgPoolBwdF {
  __device__ void prefetch(int, int, int, int) const {}
  View gx;
  const float* g;
  float inv_hw;
  int C;
  struct State {};
  __device__ void prepare(int, int, State&) const {}
  __device__ void operator()(int n, int h, int w, int c, const State&) const {
    float v[V];
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = g[(long long)n * C + c + i] * inv_hw;
    store_vec<T, V>(vptr_mut<T>(gx, n, h, w, c), v);
  }
};

template <typename TX, typename TY>
struct CastF {
  __device__ void prefetch(int, int, int, int) const {}
  View x, y;
  struct State {};
  __device__ void prepare(int, int, State&) const {}
  __device__ void operator()(int n, int h, int w, int c, const State&) const {
    *vptr_mut<TY>(y, n, h, w, c) = from_f<TY>(to_f(*vptr<TX>(x, n, h, w, c)));
  }
};

template <typename T, int V>
struct AddF {
  __device__ void prefetch(int, int, int, int) const {}
  View dst, src;
  struct State {};
  __device__ void prepare(int, int, State&) const {}
  __device__ void operator()(int n, int h, int w, int c, const State&) const {
    float a[V], b[V];
    load_vec_rw<T, V>(vptr<T>(dst, n, h, w, c), a);
    load_vec<T, V>(vptr<T>(src, n, h, w, c), b);
#pragma unroll
    for (int i = 0; i < V; ++i) a[i] += b[i];
    store_vec<T, V>(vptr_mut<T>(dst, n, h, w, c), a);
  }
};

// ---------------------------------------------------------------------------
// TMA-staged streaming form of the InstanceNorm / activation backward passes.
// The register-file drivers above keep (threads x 1-3 x 16 B) in flight and plateau at 2.5-3
// TB/s (torch's trivial add kernel: 5.2-6.5 TB/s on the same tensors, by running 2048 threads/SM
// with 8 loads each -- impossible with 30+ registers of per-channel state per thread).  Here the
// bytes in flight live in shared memory instead: one producer lane issues 1-D bulk copies
// (cp.async.bulk -> mbarrier) of whole image rows of every input into a ring of stages, 8
// consumer warps compute from shared memory and store straight to global.
//   stage = [g row (W + 2p pixels) | g alias row (reflect fold of rows 1 / H-2) | x row | g2 row]
// MODE 0: apply (writes gx / gres), MODE 1: the two per-(n,c) reductions.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t sb_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void sb_mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void sb_mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void sb_mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void sb_mbar_wait(uint32_t bar, uint32_t parity) {
  asm volatile(
      "{\n"
      ".reg .pred P1;\n"
      "SB_WAIT_LOOP:\n"
      "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
      "@P1 bra SB_DONE;\n"
      "bra SB_WAIT_LOOP;\n"
      "SB_DONE:\n"
      "}\n" ::"r"(bar),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void sb_bulk_load(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}

// position (in view coordinates) of the reflect-halo element that aliases interior index i of an
// extent-n axis with a halo of width p (nn.ReflectionPad2d), or NO_ALIAS; at most one exists
// when n >= 2p + 2
constexpr int NO_ALIAS = -1000000;
__device__ __forceinline__ int reflect_alias(int i, int n, int p) {
  if (p == 0) return NO_ALIAS;
  if (i >= 1 && i <= p) return -i;
  if (i >= n - 1 - p && i <= n - 2) return 2 * (n - 1) - i;
  return NO_ALIAS;
}

// Generic row-streaming kernel.  OP supplies up to three row inputs -- A (optionally the interior
// view of a reflect-padded tensor with halo <= 3: its aliases are folded on the fly), B, C -- and
//   NQ, State, prepare(n, c, State&), run(n, h, w, c, a[8], b[8], c[8], acc[NQ|1][8], State),
//   out_index(n, c, q) for the NQ per-(n,c) reductions (accumulated in registers per sample).
// consumer threads per CTA of the row-streaming kernel (+ one producer warp).  8 warps left an
// eligible warp in only 46 % of the cycles (profiles/r1_ncu_row_stream.md)
#ifndef OTM_RS_THREADS
#define OTM_RS_THREADS 512
#endif
constexpr int RS_THREADS = OTM_RS_THREADS;
#ifndef OTM_RS_GROUPS
#define OTM_RS_GROUPS 1
#endif
// consumer warp groups (when the ring has a multiple of it in stages).  Measured with 4 groups at
// 128x128 b32: InstanceNorm backward 1.79 -> 1.85 ms, mod_in 1.11 -> 1.19 ms, norm+act 0.49 -> 0.44
// ms per iteration -- the per-stage barrier round trip is NOT what holds these passes at ~3.7 TB/s;
// default 1 (every warp on every stage), the grouped path stays compiled and tested.
constexpr int RS_GROUPS = OTM_RS_GROUPS;

template <typename T, typename OP>
__global__ void __launch_bounds__(RS_THREADS + 32, 1)
row_stream_kernel(OP op, int N, int H, int W, int C, int stages_nseg, float* red_out) {
  // rows too long for a multi-stage ring are streamed as `nseg` column segments (one item = one
  // segment of one image row; the reflect-halo columns travel with the first / last segment)
  const int stages = stages_nseg & 0xff, nseg = (stages_nseg >> 8) & 0xff, ngroups = stages_nseg >> 16;
  const int Wseg = W / nseg;
  constexpr int V = 8;
  constexpr int NQ = OP::NQ;
  constexpr int NA = NQ > 0 ? NQ : 1;
  extern __shared__ __align__(128) unsigned char rs_smem[];
  const View va = op.in_a(), vb = op.in_b(), vc = op.in_c();
  const int p = op.a_halo();
  const int CV = C / V;
  const uint32_t a_row_bytes = (uint32_t)(Wseg + 2 * p) * C * sizeof(T);  // slot size (largest segment)
  const uint32_t x_row_bytes = (uint32_t)Wseg * C * sizeof(T);
  const uint32_t px_bytes = (uint32_t)C * sizeof(T);
  const bool has_b = vb.ptr != nullptr, has_c = vc.ptr != nullptr;
  const uint32_t off_alias = a_row_bytes, off_b = off_alias + (p ? a_row_bytes : 0);
  const uint32_t off_c = off_b + (has_b ? x_row_bytes : 0);
  const uint32_t stage_bytes = (off_c + (has_c ? x_row_bytes : 0) + 127u) & ~127u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(rs_smem + (size_t)stages * stage_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + stages;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s2 = 0; s2 < stages; ++s2) { sb_mbar_init(sb_smem(&full[s2]), 1); sb_mbar_init(sb_smem(&empty[s2]), RS_THREADS / 32 / ngroups); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int rows = N * H * nseg;  // items
  const int r0 = (int)((long long)rows * blockIdx.x / gridDim.x);
  const int r1 = (int)((long long)rows * (blockIdx.x + 1) / gridDim.x);
  if (warp == RS_THREADS / 32) {
    // ---------------- producer ----------------
    if (lane == 0) {
      for (int r = r0, k = 0; r < r1; ++r, ++k) {
        const int st = k % stages;
        sb_mbar_wait(sb_smem(&empty[st]), ((k / stages) & 1) ^ 1);
        const int row = r / nseg, seg = r - row * nseg;
        const int n = row / H, h = row - n * H;
        const int w_lo = seg * Wseg;
        // A columns [a0, a1): the segment plus the halo columns at the row ends
        const int a0 = seg == 0 ? -p : w_lo, a1 = w_lo + Wseg + (seg == nseg - 1 ? p : 0);
        const uint32_t a_bytes = (uint32_t)(a1 - a0) * px_bytes;
        const int alias = reflect_alias(h, H, p);
        const uint32_t bar = sb_smem(&full[st]);
        const uint32_t base = sb_smem(rs_smem + (size_t)st * stage_bytes);
        const uint32_t bytes = a_bytes + (alias != NO_ALIAS ? a_bytes : 0) + (has_b ? x_row_bytes : 0) +
                               (has_c ? x_row_bytes : 0);
        sb_mbar_expect_tx(bar, bytes);
        sb_bulk_load(base, vptr<T>(va, n, h, a0, 0), a_bytes, bar);
        if (alias != NO_ALIAS) sb_bulk_load(base + off_alias, vptr<T>(va, n, alias, a0, 0), a_bytes, bar);
        if (has_b) sb_bulk_load(base + off_b, vptr<T>(vb, n, h, w_lo, 0), x_row_bytes, bar);
        if (has_c) sb_bulk_load(base + off_c, vptr<T>(vc, n, h, w_lo, 0), x_row_bytes, bar);
      }
    }
    return;
  }
  // ---------------- consumers (RS_THREADS / 32 warps in `ngroups` groups) ----------------
  // The warps work in groups; group j consumes the items j, j + ngroups, ... of the CTA's range,
  // so a thread handles ngroups times more vectors per stage it waits for (one image row of 64
  // pixels x 128 channels is only 2 vectors per thread when all 512 threads share it: the barrier
  // round trip and the row bookkeeping were a third of the work per row).  `stages` is a multiple
  // of `ngroups`, so a ring stage always belongs to the same group and that group sees every phase
  // of its mbarriers (a parity wait must never be more than one phase behind).
  const int tid = threadIdx.x;  // 0..RS_THREADS-1
  const int GSZ = RS_THREADS / ngroups;
  const int grp = tid / GSZ, gt = tid % GSZ;
  const int cv = gt % CV;       // constant per thread: CV divides GSZ (a power of two)
  const int cv_sh = 31 - __clz(CV);
  typename OP::State st;
  float acc[NA][V];
#pragma unroll
  for (int q = 0; q < NA; ++q)
#pragma unroll
    for (int i = 0; i < V; ++i) acc[q][i] = 0.f;
  // per-sample flush of the register accumulators: the RS_THREADS / CV threads that share a
  // channel vector meet in shared memory first, so one atomic per (n, c, q) and CTA reaches L2
  float* red_sm = reinterpret_cast<float*>(bars + 2 * stages);  // [NQ * V][RS_THREADS]
  auto flush = [&](int n) {
    if (NQ > 0) {  // every consumer thread calls this once per sample of the CTA's range
#pragma unroll
      for (int q = 0; q < NQ; ++q)
#pragma unroll
        for (int i = 0; i < V; ++i) {
          red_sm[(q * V + i) * RS_THREADS + tid] = acc[q][i];
          acc[q][i] = 0.f;
        }
      asm volatile("bar.sync 1, %0;" ::"n"(RS_THREADS) : "memory");
      const int groups = RS_THREADS / CV;  // threads tid, tid + CV, ... hold the same channel vector
      for (int o = tid; o < NQ * V * CV; o += RS_THREADS) {
        const int c_v = o % CV, qi = o / CV;  // qi = q * V + i
        float sum = 0.f;
        for (int g2 = 0; g2 < groups; ++g2) sum += red_sm[qi * RS_THREADS + g2 * CV + c_v];
        atomicAdd(red_out + op.out_index(n, c_v * V + qi % V, qi / V), sum);
      }
      asm volatile("bar.sync 1, %0;" ::"n"(RS_THREADS) : "memory");
    }
  };
  const int ips = nseg * H;  // items per sample
  for (int n = r0 / ips; n <= (r1 - 1) / ips; ++n) {
    const int lo = max(r0, n * ips), hi = min(r1, (n + 1) * ips);
    op.prepare(n, cv * V, st);
    // this group's items of the sample: ring positions k = r - r0 with k % ngroups == grp
    for (int r = lo + ((grp - (lo - r0)) % ngroups + ngroups) % ngroups; r < hi; r += ngroups) {
      const int k = r - r0;
      const int stg = k % stages;
      const int row = r / nseg, seg = r - row * nseg;
      const int h = row - n * H;
      const int w_lo = seg * Wseg;
      const int a0 = seg == 0 ? -p : w_lo;  // first A column of the slot
      sb_mbar_wait(sb_smem(&full[stg]), (k / stages) & 1);
      const unsigned char* base = rs_smem + (size_t)stg * stage_bytes;
      const T* arow = reinterpret_cast<const T*>(base);
      const T* alrow = reinterpret_cast<const T*>(base + off_alias);
      const T* brow = reinterpret_cast<const T*>(base + off_b);
      const T* crow = reinterpret_cast<const T*>(base + off_c);
      const bool row_alias = reflect_alias(h, H, p) != NO_ALIAS;
      for (int i = gt; i < Wseg * CV; i += GSZ) {
        const int wl = i >> cv_sh;  // column within the segment
        const int w = w_lo + wl;
        float a[V], b[V], c[V];
        load_vec<T, V>(arow + (size_t)(w - a0) * C + cv * V, a);
        if (p) {
          const int wa = reflect_alias(w, W, p);  // lies in the same slot (first / last segment)
          if (wa != NO_ALIAS) {
            float t[V];
            load_vec<T, V>(arow + (size_t)(wa - a0) * C + cv * V, t);
#pragma unroll
            for (int e = 0; e < V; ++e) a[e] += t[e];
          }
          if (row_alias) {
            float t[V];
            load_vec<T, V>(alrow + (size_t)(w - a0) * C + cv * V, t);
#pragma unroll
            for (int e = 0; e < V; ++e) a[e] += t[e];
            if (wa != NO_ALIAS) {
              load_vec<T, V>(alrow + (size_t)(wa - a0) * C + cv * V, t);
#pragma unroll
              for (int e = 0; e < V; ++e) a[e] += t[e];
            }
          }
        }
        if (has_b) {
          load_vec<T, V>(brow + (size_t)wl * C + cv * V, b);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) b[e] = 0.f;
        }
        if (has_c) {
          load_vec<T, V>(crow + (size_t)wl * C + cv * V, c);
        } else {
#pragma unroll
          for (int e = 0; e < V; ++e) c[e] = 0.f;
        }
        op.run(n, h, w, cv * V, a, b, c, acc, st);
      }
      __syncwarp();
      if (lane == 0) sb_mbar_arrive(sb_smem(&empty[stg]));
    }
    flush(n);
  }
}

// InstanceNorm / activation backward: A = g (fold), B = x, C = g2.  MODE 0 apply, 1 reductions.
template <typename T, int MODE>
struct NabRowOp {
  static constexpr int NQ = MODE == 1 ? 2 : 0;
  NormActBwdApplyF<T, 8, false> f;
  using State = typename NormActBwdApplyF<T, 8, false>::State;
  __device__ View in_a() const { return f.g; }
  __device__ View in_b() const { return f.x; }
  __device__ View in_c() const { return f.g2; }
  __device__ int a_halo() const { return f.g_halo; }
  __device__ void prepare(int n, int c, State& st) const {
    if (MODE == 0) f.prepare(n, c, st); else f.prepare_stats(n, c, st);
  }
  __device__ int out_index(int n, int c, int q) const { return (n * f.C + c) * 2 + q; }
  template <int NA>
  __device__ void run(int n, int h, int w, int c, float (&ga)[8], float (&pre)[8], const float (&g2)[8],
                      float (&acc)[NA][8], const State& st) const {
    float gn[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) ga[e] += g2[e];
    if (f.stats) {
#pragma unroll
      for (int e = 0; e < 8; ++e) pre[e] = (pre[e] - st.mean[e]) * st.rstd[e];
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) gn[e] = ga[e];
    act_bwd_vec<8>(gn, pre, f.act);
    if (MODE == 1) {
#pragma unroll
      for (int e = 0; e < 8; ++e) { acc[0][e] += gn[e]; acc[NA - 1][e] += gn[e] * pre[e]; }
    } else {
      if (f.gres.ptr) store_vec<T, 8>(vptr_mut<T>(f.gres, n, h, w, c), ga);
      if (f.stats) {
#pragma unroll
        for (int e = 0; e < 8; ++e) gn[e] = st.rstd[e] * (gn[e] - st.m1[e] - pre[e] * st.m2[e]);
      }
      store_vec<T, 8>(vptr_mut<T>(f.gx, n, h, w, c), gn);
    }
  }
};

// modulated-conv input side: A = g (fold), B = x, C = gadd
template <typename T>
struct ModInRowOp {
  static constexpr int NQ = 1;
  ModInF<T, 8> f;
  using State = typename ModInF<T, 8>::State;
  __device__ View in_a() const { return f.g; }
  __device__ View in_b() const { return f.x; }
  __device__ View in_c() const { return f.gadd; }
  __device__ int a_halo() const { return f.g_halo; }
  __device__ void prepare(int n, int c, State& st) const { f.prepare(n, c, st); }
  __device__ int out_index(int n, int c, int) const { return n * f.C + c; }
  template <int NA>
  __device__ void run(int n, int h, int w, int c, float (&gt)[8], float (&xv)[8], const float (&ga)[8],
                      float (&acc)[NA][8], const State& st) const {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      acc[0][e] += gt[e] * xv[e];
      gt[e] = fmaf(gt[e], st.s[e], ga[e]);
    }
    if (!f.gx.ptr) return;
    if (f.relu_mask) {
#pragma unroll
      for (int e = 0; e < 8; ++e) gt[e] = xv[e] != 0.f ? gt[e] : 0.f;
    }
#pragma unroll
    for (int e = 0; e < 8; ++e) gt[e] *= st.gs[e];
    store_vec<T, 8>(vptr_mut<T>(f.gx, n, h, w, c), gt);
  }
};

// norm + act (+ residual) (+ reflect halo): A = x, B = residual
template <typename T>
struct NormActRowOp {
  static constexpr int NQ = 0;
  NormActF<T, 8> f;
  using State = typename NormActF<T, 8>::State;
  __device__ View in_a() const { return f.x; }
  __device__ View in_b() const { return f.res; }
  __device__ View in_c() const { return null_view_dev(); }
  __device__ int a_halo() const { return 0; }
  __device__ void prepare(int n, int c, State& st) const { f.prepare(n, c, st); }
  __device__ int out_index(int, int, int) const { return 0; }
  template <int NA>
  __device__ void run(int n, int h, int w, int c, float (&v)[8], float (&r)[8], const float (&)[8],
                      float (&)[NA][8], const State& st) const {
    if (f.stats) {
#pragma unroll
      for (int e = 0; e < 8; ++e) v[e] = (v[e] - st.mean[e]) * st.rstd[e];
    }
    act_fwd_vec<8>(v, f.act);
#pragma unroll
    for (int e = 0; e < 8; ++e) v[e] += r[e];
    store_halo<T, 8>(f.y, f.halo, n, h, w, c, v);
  }
};

// pure per-(n,c) reductions of one tensor: InstanceNorm statistics, channel sums
template <typename T>
struct StatsRowOp {
  static constexpr int NQ = 2;
  StatsF<T, 8> f;
  using State = typename StatsF<T, 8>::State;
  __device__ View in_a() const { return f.x; }
  __device__ View in_b() const { return null_view_dev(); }
  __device__ View in_c() const { return null_view_dev(); }
  __device__ int a_halo() const { return 0; }
  __device__ void prepare(int n, int c, State& st) const { f.prepare(n, c, st); }
  __device__ int out_index(int n, int c, int q) const { return f.out_index(n, c, q); }
  template <int NA>
  __device__ void run(int, int, int, int, float (&v)[8], float (&)[8], const float (&)[8],
                      float (&acc)[NA][8], const State& st) const {
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const float d = v[e] - st.k[e];
      acc[0][e] += d;
      acc[NA - 1][e] += d * d;
    }
  }
};

template <typename T>
struct ChannelSumRowOp {
  static constexpr int NQ = 1;
  ChannelSumF<T, 8> f;
  struct State {};
  __device__ View in_a() const { return f.g; }
  __device__ View in_b() const { return null_view_dev(); }
  __device__ View in_c() const { return null_view_dev(); }
  __device__ int a_halo() const { return 0; }
  __device__ void prepare(int, int, State&) const {}
  __device__ int out_index(int n, int c, int q) const { return f.out_index(n, c, q); }
  template <int NA>
  __device__ void run(int, int, int, int, float (&v)[8], float (&)[8], const float (&)[8],
                      float (&acc)[NA][8], const State&) const {
#pragma unroll
    for (int e = 0; e < 8; ++e) acc[0][e] += v[e] * f.scale;
  }
};

// Can (A with halo a_halo, B, C -> outputs o1, o2) go through row_stream_kernel?  Rows must be
// dense (pixel stride == C) and 16-byte aligned, CV must divide 256, halo 0 or 1, two stages
// must fit in shared memory.  Returns the number of stages (0 = not eligible) and the ring size.
static int row_stream_plan(const otm_tensor& A, int a_halo, const otm_tensor* B, const otm_tensor* Cc,
                           const otm_tensor* o1, const otm_tensor* o2, size_t* smem_bytes) {
  constexpr int use_stream = 1;
  if (!use_stream || a_halo > 3 || A.c % 8 != 0 || A.h < 2 * a_halo + 2 || A.w < 2 * a_halo + 2) return 0;
  // small tensors: the 148 x 288-thread persistent launch with its ring set-up costs more than
  // the register-file kernels
  constexpr long long min_bytes = 8ll << 20;
  if ((long long)A.n * A.h * A.w * A.c * (long long)dtype_size(A.dtype) < min_bytes) return 0;
  const int C = A.c, CV = C / 8;
  if (CV < 1 || 256 % CV != 0 || A.h < 4 || A.w < 4) return 0;
  auto dense = [&](const otm_tensor* t) {
    return !t || !t->ptr ||
           (t->sw == C && t->sh % 8 == 0 && t->sn % 8 == 0 && ((uintptr_t)t->ptr % 16 == 0) &&
            t->dtype == A.dtype);
  };
  if (!dense(&A) || !dense(B) || !dense(Cc) || !dense(o1) || !dense(o2)) return 0;
  const size_t es = dtype_size(A.dtype);
  const size_t scratch = 2 * 8 * RS_THREADS * sizeof(float);  // reduction flush scratch [NQ * V][threads]
  // whole rows if at least 3 stages fit, else 2 / 4 / 8 column segments per row (the 256- and
  // 512-channel tensors of the 256x256 / 512x512 configurations)
  for (int nseg = 1; nseg <= 8; nseg *= 2) {
    if (A.w % nseg != 0) break;
    const int wseg = A.w / nseg;
    if (wseg < 2 * a_halo + 2) break;
    const size_t arow = (size_t)(wseg + 2 * a_halo) * C * es, xrow = (size_t)wseg * C * es;
    if (arow % 16 || xrow % 16) return 0;
    const size_t stage =
        (arow * (a_halo ? 2 : 1) + ((B && B->ptr) ? xrow : 0) + ((Cc && Cc->ptr) ? xrow : 0) + 127) & ~(size_t)127;
    int stages = (int)((200 * 1024 - 256 - scratch) / stage);
    // prefer a ring of RS_GROUPS or 2 * RS_GROUPS stages (grouped consumers): split further
    const bool can_split = nseg < 8 && A.w % (2 * nseg) == 0 && A.w / (2 * nseg) >= 2 * a_halo + 2 &&
                           (RS_THREADS / RS_GROUPS) % CV == 0 && (A.w / (2 * nseg)) * CV >= RS_THREADS / RS_GROUPS;
    if (stages < (RS_GROUPS > 3 ? RS_GROUPS : 3) && can_split) continue;
    if (stages < 2) return 0;
    if (stages > 8) stages = 8;
    int groups = 1;
    if (RS_GROUPS > 1 && stages >= RS_GROUPS && (RS_THREADS / RS_GROUPS) % CV == 0) {
      groups = RS_GROUPS;
      stages -= stages % RS_GROUPS;
    }
    *smem_bytes = stage * stages + 2 * 8 * stages + scratch + 64;
    return stages | (nseg << 8) | (groups << 16);
  }
  return 0;
}

template <typename T, typename OP>
static int launch_row_stream(const OP& op, int N, int H, int W, int C, int stages, size_t smem,
                             float* red_out, cudaStream_t st) {
  auto kern = row_stream_kernel<T, OP>;
  OTM_ENSURE_SMEM(kern, 200 * 1024);
  int grid = num_sms();  // `stages` = stages | nseg << 8 | groups << 16 as row_stream_plan returns it
  if (grid > N * H) grid = N * H;
  kern<<<grid, RS_THREADS + 32, smem, st>>>(op, N, H, W, C, stages, red_out);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

// ---------------------------------------------------------------------------
// Windowed row streaming for the DownSample stencil (blur + bilinear, 4 x 4 taps, stride ~2):
// output row ho reads input rows i0-1 .. i0+2.  The producer lane bulk-copies every input row of
// the CTA's range ONCE into a ring of S >= 6 row slots; the consumers wait for the rows a new
// output row adds, release the rows it no longer needs, and gather their 16 taps from shared
// memory (the register-file kernel gathered them through L1/L2 at ~1 TB/s).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void down_row_window(int ho, int n_in, float scale, int& lo, int& hi) {
  const float src = fmaxf((ho + 0.5f) * scale - 0.5f, 0.f);
  const int i0 = min((int)floorf(src), n_in - 1);
  lo = max(i0 - 1, 0);
  hi = min(i0 + 2, n_in - 1);
}

template <typename T>
__global__ void __launch_bounds__(288, 1)
down_stream_kernel(DownF<T, 8> f, int N, int Ho, int Wo, int C, int S) {
  constexpr int V = 8;
  extern __shared__ __align__(128) unsigned char ds_smem[];
  const int Hi = f.x.h, Wi = f.x.w;
  const int CV = C / V;
  const uint32_t row_bytes = (uint32_t)Wi * C * sizeof(T);
  const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(ds_smem + (size_t)S * slot_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s2 = 0; s2 < S; ++s2) { sb_mbar_init(sb_smem(&full[s2]), 1); sb_mbar_init(sb_smem(&empty[s2]), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  const int rows = N * Ho;
  const int q0 = (int)((long long)rows * blockIdx.x / gridDim.x);
  const int q1 = (int)((long long)rows * (blockIdx.x + 1) / gridDim.x);
  if (warp == 8) {
    // ---------------- producer: every needed input row once, in order ----------------
    if (lane == 0) {
      int k = 0, cur_n = -1, next_in = 0;
      for (int q = q0; q < q1; ++q) {
        const int n = q / Ho, ho = q - n * Ho;
        int lo, hi;
        down_row_window(ho, Hi, f.sch, lo, hi);
        if (n != cur_n) { cur_n = n; next_in = lo; }
        for (; next_in <= hi; ++next_in, ++k) {
          const int st = k % S;
          sb_mbar_wait(sb_smem(&empty[st]), ((k / S) & 1) ^ 1);
          const uint32_t bar = sb_smem(&full[st]);
          sb_mbar_expect_tx(bar, row_bytes);
          sb_bulk_load(sb_smem(ds_smem + (size_t)st * slot_bytes), vptr<T>(f.x, n, next_in, 0, 0), row_bytes, bar);
        }
      }
    }
    return;
  }
  // ---------------- consumers ----------------
  const int tid = threadIdx.x;
  const int cv = tid % CV;
  const int cv_sh = 31 - __clz(CV);
  typename DownF<T, 8>::State st;
  int cur_n = -1;
  int first = 0, kbase = 0;   // input row `first` of the current image has load index kbase
  int waited_hi = -1;         // highest input row of the current image already waited for
  int released = 0;           // input rows [first, released) of the current image are released
  int kcount = 0;             // rows loaded so far (mirrors the producer's k)
  auto slot_of = [&](int r) { return (kbase + (r - first)) % S; };
  auto release_upto = [&](int upto) {  // warp-level: one arrival per warp and row
    // (the rows were rewritten in place through the generic proxy; order that before the bulk
    // copy that will overwrite the slot)
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncwarp();
    if (lane == 0)
      for (int r = released; r < upto; ++r) sb_mbar_arrive(sb_smem(&empty[slot_of(r)]));
    released = max(released, upto);
  };
  for (int q = q0; q < q1; ++q) {
    const int n = q / Ho, ho = q - n * Ho;
    int lo, hi;
    down_row_window(ho, Hi, f.sch, lo, hi);
    if (n != cur_n) {
      if (cur_n >= 0) release_upto(waited_hi + 1);
      cur_n = n;
      first = lo; kbase = kcount; waited_hi = lo - 1; released = lo;
      f.prepare(n, cv * V, st);
    }
    release_upto(lo);
    // rows this output row adds: wait, then normalise + activate them ONCE, in place (each input
    // element feeds up to 4 x 4 output taps; storing act(norm(x)) back at storage precision
    // leaves the stencil a plain weighted gather)
    const bool transform = f.stats != nullptr || f.act != OTM_ACT_NONE;
    for (int r = waited_hi + 1; r <= hi; ++r) {
      const int k = kbase + (r - first);
      sb_mbar_wait(sb_smem(&full[k % S]), (k / S) & 1);
      if (transform) {
        T* row = reinterpret_cast<T*>(ds_smem + (size_t)(k % S) * slot_bytes);
        for (int i = tid; i < Wi * CV; i += 256) {
          float v[V];
          load_vec<T, V>(row + (size_t)(i >> cv_sh) * C + cv * V, v);
          if (f.stats) {
#pragma unroll
            for (int e = 0; e < V; ++e) v[e] = (v[e] - st.mean[e]) * st.rstd[e];
          }
          act_fwd_vec<V>(v, f.act);
          store_vec<T, V>(row + (size_t)(i >> cv_sh) * C + cv * V, v);
        }
      }
    }
    if (hi > waited_hi) {
      kcount += hi - waited_hi;
      waited_hi = hi;
      if (transform) asm volatile("bar.sync 1, 256;" ::: "memory");  // transformed rows visible
    }
    int ph[4];
    float wh[4];
    down_taps(ho, Hi, f.sch, ph, wh);
    const T* rp[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) rp[a] = reinterpret_cast<const T*>(ds_smem + (size_t)slot_of(ph[a]) * slot_bytes);
    for (int i = tid; i < Wo * CV; i += 256) {
      const int wo = i >> cv_sh;
      int pw[4];
      float ww[4];
      down_taps(wo, Wi, f.scw, pw, ww);
      float acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
      for (int a = 0; a < 4; ++a) {
#pragma unroll
        for (int b = 0; b < 4; ++b) {
          const float wgt = wh[a] * ww[b];
          float v[V];
          load_vec<T, V>(rp[a] + (size_t)pw[b] * C + cv * V, v);
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += wgt * v[e];
        }
      }
      store_halo<T, V>(f.y, f.halo, n, ho, wo, cv * V, acc);
    }
  }
  if (cur_n >= 0) release_upto(waited_hi + 1);
}

// Transpose of the DownSample stencil in gather form with a FIXED candidate set: n_out = n_in / 2
// (DownSample always halves), so floor(src_j) is 2j or 2j+1 and input i can only be touched by the
// outputs j0 .. j0+2, j0 = max(0, (i-2) >> 1); each candidate's weight is the sum of its taps that
// land on i (zero if none).
__device__ __forceinline__ void down_bwd_taps3(int i, int n_in, int n_out, float scale,
                                               int (&js)[3], float (&wj)[3]) {
  const int j0 = max(0, (i - 2) >> 1);
#pragma unroll
  for (int c = 0; c < 3; ++c) {
    const int j = j0 + c;
    int pos[4];
    float wt[4];
    down_taps(min(j, n_out - 1), n_in, scale, pos, wt);
    float w = 0.f;
#pragma unroll
    for (int k = 0; k < 4; ++k) w += (pos[k] == i) ? wt[k] : 0.f;
    js[c] = min(j, n_out - 1);
    wj[c] = j < n_out ? w : 0.f;
  }
}

// Windowed row streaming of the DownSample backward: output (full-resolution) row h gathers the
// half-resolution gradient rows j0 .. j0+2, which the producer lane bulk-copies once each into a
// ring; the horizontal candidates / weights of every output column are tabulated in shared memory
// once per CTA.  The pass is then bound by writing ga.
template <typename T>
__global__ void __launch_bounds__(288, 1)
down_bwd_stream_kernel(View g, View ga, float sch, float scw, int C, int S) {
  constexpr int V = 8;
  extern __shared__ __align__(128) unsigned char db_smem[];
  const int N = ga.n, H = ga.h, W = ga.w, Ho = g.h, Wo = g.w;
  const int CV = C / V;
  const uint32_t row_bytes = (uint32_t)Wo * C * sizeof(T);
  const uint32_t slot_bytes = (row_bytes + 127u) & ~127u;
  uint64_t* bars = reinterpret_cast<uint64_t*>(db_smem + (size_t)S * slot_bytes);
  uint64_t* full = bars;
  uint64_t* empty = bars + S;
  int* tab_j = reinterpret_cast<int*>(bars + 2 * S);       // [W][3]
  float* tab_w = reinterpret_cast<float*>(tab_j + 3 * W);  // [W][3]
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (threadIdx.x == 0) {
    for (int s2 = 0; s2 < S; ++s2) { sb_mbar_init(sb_smem(&full[s2]), 1); sb_mbar_init(sb_smem(&empty[s2]), 8); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  for (int w = threadIdx.x; w < W; w += blockDim.x) {
    int js[3];
    float wj[3];
    down_bwd_taps3(w, W, Wo, scw, js, wj);
#pragma unroll
    for (int c = 0; c < 3; ++c) { tab_j[3 * w + c] = js[c]; tab_w[3 * w + c] = wj[c]; }
  }
  __syncthreads();
  const int rows = N * H;
  const int q0 = (int)((long long)rows * blockIdx.x / gridDim.x);
  const int q1 = (int)((long long)rows * (blockIdx.x + 1) / gridDim.x);
  auto window = [&](int h, int& lo, int& hi) {
    lo = min(max(0, (h - 2) >> 1), Ho - 1);
    hi = min(max(0, (h - 2) >> 1) + 2, Ho - 1);
  };
  if (warp == 8) {
    if (lane == 0) {
      int k = 0, cur_n = -1, next_in = 0;
      for (int q = q0; q < q1; ++q) {
        const int n = q / H, h = q - n * H;
        int lo, hi;
        window(h, lo, hi);
        if (n != cur_n) { cur_n = n; next_in = lo; }
        for (; next_in <= hi; ++next_in, ++k) {
          const int st = k % S;
          sb_mbar_wait(sb_smem(&empty[st]), ((k / S) & 1) ^ 1);
          const uint32_t bar = sb_smem(&full[st]);
          sb_mbar_expect_tx(bar, row_bytes);
          sb_bulk_load(sb_smem(db_smem + (size_t)st * slot_bytes), vptr<T>(g, n, next_in, 0, 0), row_bytes, bar);
        }
      }
    }
    return;
  }
  const int tid = threadIdx.x;
  const int cv = tid % CV;
  const int cv_sh = 31 - __clz(CV);
  int cur_n = -1, first = 0, kbase = 0, waited_hi = -1, released = 0, kcount = 0;
  auto slot_of = [&](int r) { return (kbase + (r - first)) % S; };
  auto release_upto = [&](int upto) {
    __syncwarp();
    if (lane == 0)
      for (int r = released; r < upto; ++r) sb_mbar_arrive(sb_smem(&empty[slot_of(r)]));
    released = max(released, upto);
  };
  for (int q = q0; q < q1; ++q) {
    const int n = q / H, h = q - n * H;
    int lo, hi;
    window(h, lo, hi);
    if (n != cur_n) {
      if (cur_n >= 0) release_upto(waited_hi + 1);
      cur_n = n;
      first = lo; kbase = kcount; waited_hi = lo - 1; released = lo;
    }
    release_upto(lo);
    for (int r = waited_hi + 1; r <= hi; ++r) {
      const int k = kbase + (r - first);
      sb_mbar_wait(sb_smem(&full[k % S]), (k / S) & 1);
    }
    if (hi > waited_hi) { kcount += hi - waited_hi; waited_hi = hi; }
    int jh[3];
    float wh[3];
    down_bwd_taps3(h, H, Ho, sch, jh, wh);
    const T* rp[3];
#pragma unroll
    for (int a = 0; a < 3; ++a)
      rp[a] = reinterpret_cast<const T*>(db_smem + (size_t)slot_of(min(max(jh[a], lo), hi)) * slot_bytes);
    for (int i = tid; i < W * CV; i += 256) {
      const int w = i >> cv_sh;
      float acc[V];
#pragma unroll
      for (int e = 0; e < V; ++e) acc[e] = 0.f;
#pragma unroll
      for (int b = 0; b < 3; ++b) {
        const int jw = tab_j[3 * w + b];
        const float wwb = tab_w[3 * w + b];
#pragma unroll
        for (int a = 0; a < 3; ++a) {
          float v[V];
          load_vec<T, V>(rp[a] + (size_t)jw * C + cv * V, v);
          const float wgt = wh[a] * wwb;
#pragma unroll
          for (int e = 0; e < V; ++e) acc[e] += wgt * v[e];
        }
      }
      store_vec<T, V>(vptr_mut<T>(ga, n, h, w, cv * V), acc);
    }
  }
  if (cur_n >= 0) release_upto(waited_hi + 1);
}

static int down_stream_slots(const otm_down_args* a, size_t* smem_bytes) {
  constexpr int use_stream = 1;
  const otm_tensor& x = a->x;
  if (!use_stream || x.c % 8 != 0) return 0;
  const int C = x.c, CV = C / 8;
  if (CV < 1 || 256 % CV != 0 || a->y.h < 2 || a->y.w < 2) return 0;
  if (x.sw != C || x.sh % 8 || x.sn % 8 || ((uintptr_t)x.ptr % 16)) return 0;
  const size_t es = dtype_size(x.dtype);
  constexpr long long min_bytes = 8ll << 20;
  if ((long long)x.n * x.h * x.w * C * (long long)es < min_bytes) return 0;
  const size_t row = ((size_t)x.w * C * es + 127) & ~(size_t)127;
  if (((size_t)x.w * C * es) % 16) return 0;
  int S = (int)((200 * 1024 - 256) / row);
  if (S < 6) return 0;
  if (S > 12) S = 12;
  *smem_bytes = row * S + 2 * 8 * S + 64;
  return S;
}

static bool same_shape(const otm_tensor& a, const otm_tensor& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c;
}

}  // namespace otm

using namespace otm;

// dispatch on (dtype, vector width) and run BODY with T and V defined
#define OTM_DISPATCH_TV(dt, vecok, ...)                                      \
  do {                                                                       \
    if ((dt) == OTM_BF16) {                                                  \
      using T = __nv_bfloat16;                                               \
      if (vecok) { constexpr int V = 8; __VA_ARGS__; }                       \
      else { constexpr int V = 1; __VA_ARGS__; }                             \
    } else {                                                                 \
      using T = float;                                                       \
      if (vecok) { constexpr int V = 8; __VA_ARGS__; }                       \
      else { constexpr int V = 1; __VA_ARGS__; }                             \
    }                                                                        \
  } while (0)

extern "C" {

const char* otm_last_error(void) { return g_err; }
int otm_version(void) { return 1; }
int64_t otm_launch_count(void) { return g_launches.load(); }

int otm_instnorm_finalize(const float* sums, float* stats, int32_t count, int32_t hw, float eps,
                          otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(sums && stats && count > 0 && hw > 0, "instnorm_finalize: bad argument");
  stats_from_sums_kernel<<<(count + 255) / 256, 256, 0, st>>>(sums, stats, count, 1.f / (float)hw, eps);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_instnorm_stats(const otm_tensor* x, float eps, float* ws, float* stats,
                       otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(x && x->ptr && ws && stats, "instnorm_stats: null argument");
  int count = x->n * x->c;
  OTM_CHECK_CUDA(cudaMemsetAsync(ws, 0, sizeof(float) * 2 * count, st));
  bool vok = vec_ok(*x, 8);
  int rc = OTM_OK;
  size_t rs_smem_bytes = 0;
  // pure read reductions through the row-streaming kernel: 31.7 vs 33.8 us on [64,128,64,64], 70.7
  // vs 80.9 us on [64,128,128,128] once the flush goes through shared memory (with one atomic per
  // thread it was slower); OTM_STREAM_REDUCE=0 selects the register-file reduction
  constexpr int stream_red = 1;
  const int rs_stages = (vok && stream_red)
                            ? row_stream_plan(*x, 0, nullptr, nullptr, nullptr, nullptr, &rs_smem_bytes) : 0;
  if (rs_stages) {
    if (x->dtype == OTM_BF16) {
      StatsRowOp<__nv_bfloat16> op{StatsF<__nv_bfloat16, 8>{make_view(*x), x->c}};
      rc = launch_row_stream<__nv_bfloat16>(op, x->n, x->h, x->w, x->c, rs_stages, rs_smem_bytes, ws, st);
    } else {
      StatsRowOp<float> op{StatsF<float, 8>{make_view(*x), x->c}};
      rc = launch_row_stream<float>(op, x->n, x->h, x->w, x->c, rs_stages, rs_smem_bytes, ws, st);
    }
  } else {
    OTM_DISPATCH_TV(x->dtype, vok, {
      StatsF<T, V> f{make_view(*x), x->c};
      rc = launch_nc_reduce<V>(f, x->n, x->h, x->w, x->c, ws, st);
    });
  }
  if (rc) return rc;
  if (x->dtype == OTM_BF16)
    stats_finalize_kernel<__nv_bfloat16><<<(count + 255) / 256, 256, 0, st>>>(
        ws, stats, make_view(*x), count, 1.f / (float)(x->h * x->w), eps);
  else
    stats_finalize_kernel<float><<<(count + 255) / 256, 256, 0, st>>>(
        ws, stats, make_view(*x), count, 1.f / (float)(x->h * x->w), eps);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_norm_act(const otm_norm_act_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->x.ptr && a->y.ptr, "norm_act: null tensor");
  OTM_REQUIRE(same_shape(a->x, a->y), "norm_act: x/y shape mismatch");
  OTM_REQUIRE(a->x.dtype == a->y.dtype, "norm_act: dtype mismatch");
  OTM_REQUIRE(a->y_halo >= 0 && a->y_halo < a->y.h && a->y_halo < a->y.w,
              "norm_act: reflect halo %d too large for %dx%d", a->y_halo, a->y.h, a->y.w);
  if (a->residual.ptr) {
    OTM_REQUIRE(same_shape(a->x, a->residual) && a->residual.dtype == a->x.dtype,
                "norm_act: residual mismatch");
  }
  bool vok = vec_ok(a->x, 8) && vec_ok(a->y, 8) && vec_ok(a->residual, 8);
  int rc = OTM_OK;
  {
    size_t smem = 0;
    const int stages = vok ? row_stream_plan(a->x, 0, &a->residual, nullptr, &a->y, nullptr, &smem) : 0;
    if (stages) {
      if (a->x.dtype == OTM_BF16) {
        NormActRowOp<__nv_bfloat16> op{NormActF<__nv_bfloat16, 8>{
            make_view(a->x), a->residual.ptr ? make_view(a->residual) : null_view(), make_view(a->y),
            a->stats, a->act, a->y_halo, a->x.c}};
        return launch_row_stream<__nv_bfloat16>(op, a->x.n, a->x.h, a->x.w, a->x.c, stages, smem, nullptr, st);
      }
      NormActRowOp<float> op{NormActF<float, 8>{
          make_view(a->x), a->residual.ptr ? make_view(a->residual) : null_view(), make_view(a->y),
          a->stats, a->act, a->y_halo, a->x.c}};
      return launch_row_stream<float>(op, a->x.n, a->x.h, a->x.w, a->x.c, stages, smem, nullptr, st);
    }
  }
  OTM_DISPATCH_TV(a->x.dtype, vok, {
    NormActF<T, V> f{make_view(a->x), a->residual.ptr ? make_view(a->residual) : null_view(),
                     make_view(a->y), a->stats, a->act, a->y_halo, a->x.c};
    rc = launch_ew<V>(f, a->x.n, a->x.h, a->x.w, a->x.c, st);
  });
  return rc;
}

static int norm_act_bwd_impl(const otm_norm_act_bwd_args* a, otm_stream stream);

// n0..n0+cnt samples of a tensor (NULL stays NULL)
static otm_tensor slice_n(const otm_tensor& t, int n0, int cnt) {
  otm_tensor r = t;
  if (t.ptr) {
    r.ptr = (char*)t.ptr + (size_t)n0 * (size_t)t.sn * dtype_size(t.dtype);
    r.n = cnt;
  }
  return r;
}

// InstanceNorm backward is two passes over (g, x): the per-(n,c) reductions, then the apply.
// OTM_L2_BLOCK_MB=<m> runs them back to back on blocks of samples of <= m MB so that the second
// pass could find the block's g and x in the 126 MB L2 (2R+1W of HBM traffic instead of 4R+1W).
// Measured on B200 at 32 MB blocks: SLOWER (3.65 vs 3.0 ms per iteration for all norm backwards,
// 1034 vs 1059 img/s) -- the 2-4x smaller launches lose more to their prologues and tails than
// the L2 hits give back.  Off by default; kept as a knob for the larger configs.
int otm_norm_act_bwd(const otm_norm_act_bwd_args* a, otm_stream stream) {
  constexpr long long block_bytes = 0;  // L2-blocked variant measured slower; kept for reference
  if (a && a->stats && a->gx.ptr && a->x.ptr && a->g.ptr && block_bytes > 0) {
    const long long per_sample =
        (long long)a->gx.h * a->gx.w * a->gx.c * (long long)dtype_size(a->gx.dtype) *
        (2 + (a->g2.ptr ? 1 : 0));
    int cn = (int)(block_bytes / (per_sample > 0 ? per_sample : 1));
    if (cn < 1) cn = 1;
    if (cn < a->gx.n) {
      const int nblk = (a->gx.n + cn - 1) / cn;
      cn = (a->gx.n + nblk - 1) / nblk;  // even blocks
      for (int n0 = 0; n0 < a->gx.n; n0 += cn) {
        const int cnt = a->gx.n - n0 < cn ? a->gx.n - n0 : cn;
        otm_norm_act_bwd_args b = *a;
        b.g = slice_n(a->g, n0, cnt); b.g2 = slice_n(a->g2, n0, cnt); b.x = slice_n(a->x, n0, cnt);
        b.gx = slice_n(a->gx, n0, cnt); b.gres = slice_n(a->gres, n0, cnt);
        b.stats = a->stats + (size_t)n0 * a->gx.c * 2;
        b.sums = a->sums ? a->sums + (size_t)n0 * a->gx.c * 2 : nullptr;
        int rc = norm_act_bwd_impl(&b, stream);
        if (rc) return rc;
      }
      return OTM_OK;
    }
  }
  return norm_act_bwd_impl(a, stream);
}

static int norm_act_bwd_impl(const otm_norm_act_bwd_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->g.ptr && a->gx.ptr, "norm_act_bwd: null tensor");
  // x (the forward input) may be omitted for a pure fold/add pass (no norm, no activation)
  OTM_REQUIRE(a->x.ptr || (a->act == OTM_ACT_NONE && !a->stats), "norm_act_bwd: x required");
  if (a->g_down) {
    OTM_REQUIRE(a->x.ptr && same_shape(a->gx, a->x) && a->g.n == a->x.n && a->g.c == a->x.c &&
                    a->g.h == a->x.h / 2 && a->g.w == a->x.w / 2,
                "norm_act_bwd: g_down needs g of shape [n, H/2, W/2, c]");
  } else {
    OTM_REQUIRE(same_shape(a->g, a->gx) && (!a->x.ptr || same_shape(a->gx, a->x)),
                "norm_act_bwd: shape mismatch");
  }
  OTM_REQUIRE(a->g.dtype == a->gx.dtype && (!a->x.ptr || a->gx.dtype == a->x.dtype),
              "norm_act_bwd: dtype");
  OTM_REQUIRE(!a->stats || a->sums, "norm_act_bwd: sums workspace required with stats");
  bool vok = vec_ok(a->g, 8) && vec_ok(a->x, 8) && vec_ok(a->gx, 8) && vec_ok(a->g2, 8) &&
             vec_ok(a->gres, 8);
  int rc = OTM_OK;
  const otm_tensor& sh = a->gx;
  const int C = sh.c;
#define OTM_NAB_BODY(DOWN)                                                                       \
  do {                                                                                           \
    if (a->stats) {                                                                              \
      NormActBwdReduceF<T, V, DOWN> r;                                                           \
      r.g = make_view(a->g); r.g2 = a->g2.ptr ? make_view(a->g2) : null_view();                  \
      r.x = make_view(a->x); r.stats = a->stats; r.act = a->act; r.g_halo = a->g_halo; r.C = C;  \
      r.g_down = a->g_down; r.sch = sch; r.scw = scw;                                            \
      rc = launch_nc_reduce<V>(r, sh.n, sh.h, sh.w, C, a->sums, st);                             \
    }                                                                                            \
    if (rc == OTM_OK) {                                                                          \
      NormActBwdApplyF<T, V, DOWN> f;                                                            \
      f.g = make_view(a->g); f.g2 = a->g2.ptr ? make_view(a->g2) : null_view();                  \
      f.x = a->x.ptr ? make_view(a->x) : null_view();                                            \
      f.stats = a->stats; f.act = a->act; f.g_halo = a->g_halo; f.C = C;                         \
      f.g_down = a->g_down; f.sch = sch; f.scw = scw;                                            \
      f.gx = make_view(a->gx); f.gres = a->gres.ptr ? make_view(a->gres) : null_view();          \
      f.sums = a->sums; f.inv_hw = 1.f / (float)(sh.h * sh.w);                                   \
      rc = launch_ew<V>(f, sh.n, sh.h, sh.w, C, st);                                             \
    }                                                                                            \
  } while (0)
  const float sch = a->g_down ? (float)a->x.h / (float)a->g.h : 1.f;
  const float scw = a->g_down ? (float)a->x.w / (float)a->g.w : 1.f;
  if (a->stats) OTM_CHECK_CUDA(cudaMemsetAsync(a->sums, 0, sizeof(float) * 2 * sh.n * C, st));
  {
    size_t smem = 0;
    const int stages = (vok && !a->g_down)
                           ? row_stream_plan(a->g, a->g_halo, &a->x, &a->g2, &a->gx, &a->gres, &smem)
                           : 0;
    if (stages) {
#define OTM_NAB_STREAM(T)                                                                          \
  do {                                                                                              \
    NormActBwdApplyF<T, 8, false> f;                                                                \
    f.g = make_view(a->g); f.g2 = a->g2.ptr ? make_view(a->g2) : null_view();                       \
    f.x = a->x.ptr ? make_view(a->x) : null_view();                                                 \
    f.stats = a->stats; f.act = a->act; f.g_halo = a->g_halo; f.C = C;                              \
    f.g_down = 0; f.sch = 1.f; f.scw = 1.f;                                                         \
    f.gx = make_view(a->gx); f.gres = a->gres.ptr ? make_view(a->gres) : null_view();               \
    f.sums = a->sums; f.inv_hw = 1.f / (float)(sh.h * sh.w);                                        \
    if (a->stats) {                                                                                 \
      NabRowOp<T, 1> r1{f};                                                                         \
      rc = launch_row_stream<T>(r1, sh.n, sh.h, sh.w, C, stages, smem, a->sums, st);                \
    }                                                                                               \
    if (rc == OTM_OK) {                                                                             \
      NabRowOp<T, 0> r0{f};                                                                         \
      rc = launch_row_stream<T>(r0, sh.n, sh.h, sh.w, C, stages, smem, nullptr, st);                \
    }                                                                                               \
  } while (0)
      if (sh.dtype == OTM_BF16) OTM_NAB_STREAM(__nv_bfloat16); else OTM_NAB_STREAM(float);
#undef OTM_NAB_STREAM
      return rc;
    }
  }
  OTM_DISPATCH_TV(sh.dtype, vok, {
    if (a->g_down) OTM_NAB_BODY(true);
    else OTM_NAB_BODY(false);
  });
#undef OTM_NAB_BODY
  return rc;
}

int otm_norm_act_bwd_bwd(const otm_tensor* g, const otm_tensor* gg, const otm_tensor* x,
                         const float* stats, int32_t act, const otm_tensor* dg, const otm_tensor* dx,
                         float* sums, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(g && gg && x && g->ptr && gg->ptr && x->ptr && stats && sums && dg && dx,
              "norm_act_bwd_bwd: null argument");
  OTM_REQUIRE(same_shape(*g, *x) && same_shape(*gg, *x) && g->dtype == x->dtype && gg->dtype == x->dtype,
              "norm_act_bwd_bwd: shape / dtype mismatch");
  OTM_REQUIRE(act == OTM_ACT_NONE || act == OTM_ACT_RELU || act == OTM_ACT_LRELU,
              "norm_act_bwd_bwd: piecewise-linear activations only");
  const int C = x->c;
  OTM_CHECK_CUDA(cudaMemsetAsync(sums, 0, sizeof(float) * 5 * x->n * C, st));
  const bool vok = vec_ok(*g, 8) && vec_ok(*gg, 8) && vec_ok(*x, 8) && vec_ok(*dg, 8) && vec_ok(*dx, 8);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(x->dtype, vok, {
    Norm2ReduceF<T, V> r;
    r.g = make_view(*g); r.gg = make_view(*gg); r.x = make_view(*x); r.stats = stats; r.act = act; r.C = C;
    rc = launch_nc_reduce<V>(r, x->n, x->h, x->w, C, sums, st);
    if (rc == OTM_OK) {
      Norm2ApplyF<T, V> f;
      f.g = make_view(*g); f.gg = make_view(*gg); f.x = make_view(*x); f.stats = stats; f.act = act; f.C = C;
      f.dg = dg->ptr ? make_view(*dg) : null_view(); f.dx = dx->ptr ? make_view(*dx) : null_view();
      f.sums = sums; f.inv_hw = 1.f / (float)(x->h * x->w);
      rc = launch_ew<V>(f, x->n, x->h, x->w, C, st);
    }
  });
  return rc;
}

int otm_down(const otm_down_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->x.ptr && a->y.ptr, "down: null tensor");
  OTM_REQUIRE(a->y.h == a->x.h / 2 && a->y.w == a->x.w / 2 && a->y.n == a->x.n &&
                  a->y.c == a->x.c && a->y.h > 0 && a->y.w > 0,
              "down: y must be [n, H/2, W/2, c]");
  OTM_REQUIRE(a->x.dtype == a->y.dtype, "down: dtype mismatch");
  OTM_REQUIRE(a->y_halo >= 0 && (a->y_halo == 0 || (a->y_halo < a->y.h && a->y_halo < a->y.w)),
              "down: halo too large");
  bool vok = vec_ok(a->x, 8) && vec_ok(a->y, 8);
  int rc = OTM_OK;
  // OTM_DOWN_FWD_BLOCK=1: 2x2 output blocks from a 6x6 window.  Measured SLOWER than the per-output
  // functor (1.54 vs 1.02 ms per iteration at 128x128, 7.2 vs 5.3 ms at 256x256: 128 registers ->
  // 2 CTAs/SM); off by default.  The backward block form (DownBwd2x2F) is faster and is on.
  {
    size_t smem = 0;
    const int S = vok ? down_stream_slots(a, &smem) : 0;
    if (S) {
      int grid = num_sms();
      if (grid > a->y.n * a->y.h) grid = a->y.n * a->y.h;
#define OTM_DOWN_STREAM(T)                                                                        \
  do {                                                                                             \
    DownF<T, 8> f{make_view(a->x), make_view(a->y), a->stats, a->act, a->y_halo, a->x.c,           \
                  (float)a->x.h / (float)a->y.h, (float)a->x.w / (float)a->y.w};                   \
    auto kern = down_stream_kernel<T>;                                                             \
    OTM_ENSURE_SMEM(kern, 200 * 1024);                                                                                              \
    kern<<<grid, 288, smem, st>>>(f, a->y.n, a->y.h, a->y.w, a->y.c, S);                           \
  } while (0)
      if (a->x.dtype == OTM_BF16) OTM_DOWN_STREAM(__nv_bfloat16); else OTM_DOWN_STREAM(float);
#undef OTM_DOWN_STREAM
      OTM_LAUNCH_CHECK();
      return OTM_OK;
    }
  }
  constexpr int blk = 0;  // the 2x2-block form measured slower (occupancy), DESIGN.md
  const bool even = a->x.h == 2 * a->y.h && a->x.w == 2 * a->y.w && a->y.h >= 4 && a->y.w >= 4;
  OTM_DISPATCH_TV(a->x.dtype, vok, {
    if (blk && V == 8 && even) {
      Down2x2F<T, V> f{make_view(a->x), make_view(a->y), a->stats, a->act, a->y_halo, a->x.c};
      rc = launch_ew<V>(f, a->y.n, (a->y.h + 1) / 2, (a->y.w + 1) / 2, a->y.c, st);
    } else {
      DownF<T, V> f{make_view(a->x), make_view(a->y), a->stats, a->act, a->y_halo, a->x.c,
                    (float)a->x.h / (float)a->y.h, (float)a->x.w / (float)a->y.w};
      rc = launch_ew<V>(f, a->y.n, a->y.h, a->y.w, a->y.c, st);
    }
  });
  return rc;
}

int otm_down_bwd(const otm_tensor* g, int32_t g_halo, const otm_tensor* ga, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(g && ga && g->ptr && ga->ptr, "down_bwd: null tensor");
  OTM_REQUIRE(g->h == ga->h / 2 && g->w == ga->w / 2 && g->n == ga->n && g->c == ga->c,
              "down_bwd: shape mismatch");
  OTM_REQUIRE(g->dtype == ga->dtype, "down_bwd: dtype mismatch");
  bool vok = vec_ok(*g, 8) && vec_ok(*ga, 8);
  int rc = OTM_OK;
  {
    // windowed row streaming (odd and even sizes): g rows are small, the pass is write-bound
    constexpr int use_stream = 1;
    constexpr long long min_bytes = 8ll << 20;
    const int C = g->c, CV = C / 8;
    const size_t es = dtype_size(g->dtype);
    const size_t row = ((size_t)g->w * C * es + 127) & ~(size_t)127;
    const bool ok = use_stream && vok && g_halo == 0 && C % 8 == 0 && CV >= 1 && 256 % CV == 0 &&
                    g->sw == C && g->sh % 8 == 0 && g->sn % 8 == 0 && ((uintptr_t)g->ptr % 16 == 0) &&
                    ((size_t)g->w * C * es) % 16 == 0 && g->h >= 3 && g->w >= 3 && ga->w <= 1024 &&
                    (long long)ga->n * ga->h * ga->w * C * (long long)es >= min_bytes;
    int S = ok ? (int)((160 * 1024) / row) : 0;
    if (S > 12) S = 12;
    if (S >= 5) {
      const size_t smem = row * S + 2 * 8 * S + (size_t)ga->w * 3 * 8 + 64;
      int grid = num_sms();
      if (grid > ga->n * ga->h) grid = ga->n * ga->h;
#define OTM_DBS(T)                                                                                 \
  do {                                                                                              \
    auto kern = down_bwd_stream_kernel<T>;                                                          \
    OTM_ENSURE_SMEM(kern, 200 * 1024);                                                                                               \
    kern<<<grid, 288, smem, st>>>(make_view(*g), make_view(*ga), (float)ga->h / (float)g->h,        \
                                  (float)ga->w / (float)g->w, C, S);                                \
  } while (0)
      if (g->dtype == OTM_BF16) OTM_DBS(__nv_bfloat16); else OTM_DBS(float);
#undef OTM_DBS
      OTM_LAUNCH_CHECK();
      return OTM_OK;
    }
  }
  constexpr int blk = 1;
  const bool even = ga->h == 2 * g->h && ga->w == 2 * g->w && g->h >= 4 && g->w >= 4;
  OTM_DISPATCH_TV(g->dtype, vok, {
    if (blk && V == 8 && even && g_halo == 0) {
      DownBwd2x2F<T, V> f{make_view(*g), make_view(*ga)};
      rc = launch_ew<V>(f, ga->n, ga->h / 2, ga->w / 2, ga->c, st);
    } else {
      DownBwdF<T, V> f{make_view(*g), make_view(*ga), g_halo, (float)ga->h / (float)g->h,
                       (float)ga->w / (float)g->w};
      rc = launch_ew<V>(f, ga->n, ga->h, ga->w, ga->c, st);
    }
  });
  return rc;
}

int otm_up(const otm_tensor* x, const otm_tensor* y, int32_t y_halo, const float* scale,
           otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(x && y && x->ptr && y->ptr, "up: null tensor");
  OTM_REQUIRE(y->h == 2 * x->h && y->w == 2 * x->w && y->n == x->n && y->c == x->c,
              "up: y must be [n, 2H, 2W, c]");
  OTM_REQUIRE(x->dtype == y->dtype, "up: dtype mismatch");
  bool vok = vec_ok(*x, 8) && vec_ok(*y, 8);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(x->dtype, vok, {
    if (V == 8 && x->h >= 4 && x->w >= 4) {
      Up2x2F<T, V> f{make_view(*x), make_view(*y), y_halo, scale, x->c};
      rc = launch_ew<V>(f, x->n, x->h, x->w, x->c, st);
    } else {
      UpF<T, V> f{make_view(*x), make_view(*y), y_halo, scale, x->c};
      rc = launch_ew<V>(f, y->n, y->h, y->w, y->c, st);
    }
  });
  return rc;
}

int otm_up_bwd(const otm_tensor* g, int32_t g_halo, const otm_tensor* gx, const float* scale,
               otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(g && gx && g->ptr && gx->ptr, "up_bwd: null tensor");
  OTM_REQUIRE(g->h == 2 * gx->h && g->w == 2 * gx->w && g->n == gx->n && g->c == gx->c,
              "up_bwd: shape mismatch");
  OTM_REQUIRE(g->dtype == gx->dtype, "up_bwd: dtype mismatch");
  bool vok = vec_ok(*g, 8) && vec_ok(*gx, 8);
  int rc = OTM_OK;
  // (A windowed shared-memory streaming form of this pass -- pairs of gx rows x column segments
  // from a ring of g row segments, every g byte bulk-copied once -- was built and measured at 352
  // vs 321 us per launch: the pass is ALU-bound (30 fp32 FMAs + conversions per output element for
  // the 6 x 6 tap transpose), not load-bound, and 8 consumer warps issue less than 32.)
  OTM_DISPATCH_TV(g->dtype, vok, {
    if (V == 8 && g_halo == 0 && gx->h >= 8 && gx->w >= 8) {
      UpBwd2x2F<T, V> f{make_view(*g), make_view(*gx), scale, gx->c};
      rc = launch_ew<V>(f, gx->n, (gx->h + 1) / 2, (gx->w + 1) / 2, gx->c, st);
    } else {
      UpBwdF<T, V> f{make_view(*g), make_view(*gx), g_halo, scale, gx->c};
      rc = launch_ew<V>(f, gx->n, gx->h, gx->w, gx->c, st);
    }
  });
  return rc;
}

int otm_mod_out(const otm_mod_out_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->g.ptr && a->out.ptr && a->P, "mod_out: null argument");
  OTM_REQUIRE(same_shape(a->g, a->out), "mod_out: shape mismatch");
  OTM_REQUIRE(a->act == OTM_ACT_NONE || a->act == OTM_ACT_RELU, "mod_out: act must be none/relu");
  const int C = a->out.c;
  OTM_CHECK_CUDA(cudaMemsetAsync(a->P, 0, sizeof(float) * a->out.n * C, st));
  bool vok = vec_ok(a->g, 8) && vec_ok(a->out, 8) && vec_ok(a->g2, 8) && vec_ok(a->res, 8) &&
             vec_ok(a->gy, 8);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(a->out.dtype, vok, {
    ModOutF<T, V> f{make_view(a->g),
                    a->g2.ptr ? make_view(a->g2) : null_view(),
                    make_view(a->out),
                    a->res.ptr ? make_view(a->res) : null_view(),
                    a->gy.ptr ? make_view(a->gy) : null_view(),
                    a->g_halo, a->act, C, a->gy_scale};
    rc = launch_nc_reduce<V>(f, a->out.n, a->out.h, a->out.w, C, a->P, st);
  });
  return rc;
}

int otm_mod_in(const otm_mod_in_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->g.ptr && a->x.ptr && a->Q && a->s, "mod_in: null argument");
  OTM_REQUIRE(same_shape(a->g, a->x) && (!a->gx.ptr || same_shape(a->gx, a->x)), "mod_in: shape mismatch");
  const int C = a->x.c;
  OTM_CHECK_CUDA(cudaMemsetAsync(a->Q, 0, sizeof(float) * a->x.n * C, st));
  bool vok = vec_ok(a->g, 8) && vec_ok(a->x, 8) && vec_ok(a->gadd, 8) && vec_ok(a->gx, 8);
  int rc = OTM_OK;
  {
    size_t smem = 0;
    const int stages = vok ? row_stream_plan(a->g, a->g_halo, &a->x, &a->gadd, &a->gx, nullptr, &smem) : 0;
    if (stages) {
      if (a->x.dtype == OTM_BF16) {
        ModInRowOp<__nv_bfloat16> op{ModInF<__nv_bfloat16, 8>{
            make_view(a->g), make_view(a->x), a->gadd.ptr ? make_view(a->gadd) : null_view(),
            a->gx.ptr ? make_view(a->gx) : null_view(), a->s, a->g_halo, C, a->relu_mask, a->gx_scale}};
        return launch_row_stream<__nv_bfloat16>(op, a->x.n, a->x.h, a->x.w, C, stages, smem, a->Q, st);
      }
      ModInRowOp<float> op{ModInF<float, 8>{
          make_view(a->g), make_view(a->x), a->gadd.ptr ? make_view(a->gadd) : null_view(),
          a->gx.ptr ? make_view(a->gx) : null_view(), a->s, a->g_halo, C, a->relu_mask, a->gx_scale}};
      return launch_row_stream<float>(op, a->x.n, a->x.h, a->x.w, C, stages, smem, a->Q, st);
    }
  }
  OTM_DISPATCH_TV(a->x.dtype, vok, {
    ModInF<T, V> f{make_view(a->g), make_view(a->x),
                   a->gadd.ptr ? make_view(a->gadd) : null_view(),
                   a->gx.ptr ? make_view(a->gx) : null_view(), a->s, a->g_halo, C, a->relu_mask,
                   a->gx_scale};
    rc = launch_nc_reduce<V>(f, a->x.n, a->x.h, a->x.w, C, a->Q, st);
  });
  return rc;
}

int otm_channel_sum(const otm_tensor* g, float* out, int32_t accumulate, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(g && g->ptr && out, "channel_sum: null argument");
  if (!accumulate) OTM_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * g->c, st));
  bool vok = vec_ok(*g, 8);
  int rc = OTM_OK;
  {
    size_t smem = 0;
    constexpr int stream_red = 1;
    const int stages = (vok && stream_red) ? row_stream_plan(*g, 0, nullptr, nullptr, nullptr, nullptr, &smem) : 0;
    if (stages) {
      if (g->dtype == OTM_BF16) {
        ChannelSumRowOp<__nv_bfloat16> op{ChannelSumF<__nv_bfloat16, 8>{make_view(*g), 1.f, 0, g->c}};
        return launch_row_stream<__nv_bfloat16>(op, g->n, g->h, g->w, g->c, stages, smem, out, st);
      }
      ChannelSumRowOp<float> op{ChannelSumF<float, 8>{make_view(*g), 1.f, 0, g->c}};
      return launch_row_stream<float>(op, g->n, g->h, g->w, g->c, stages, smem, out, st);
    }
  }
  OTM_DISPATCH_TV(g->dtype, vok, {
    ChannelSumF<T, V> f{make_view(*g), 1.f, 0, g->c};
    rc = launch_nc_reduce<V>(f, g->n, g->h, g->w, g->c, out, st);
  });
  return rc;
}

int otm_avgpool(const otm_tensor* x, float* out, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(x && x->ptr && out, "avgpool: null argument");
  OTM_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * x->n * x->c, st));
  bool vok = vec_ok(*x, 8);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(x->dtype, vok, {
    ChannelSumF<T, V> f{make_view(*x), 1.f / (float)(x->h * x->w), 1, x->c};
    rc = launch_nc_reduce<V>(f, x->n, x->h, x->w, x->c, out, st);
  });
  return rc;
}

int otm_avgpool_bwd(const float* g, const otm_tensor* gx, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(g && gx && gx->ptr, "avgpool_bwd: null argument");
  bool vok = vec_ok(*gx, 8);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(gx->dtype, vok, {
    AvgPoolBwdF<T, V> f{make_view(*gx), g, 1.f / (float)(gx->h * gx->w), gx->c};
    rc = launch_ew<V>(f, gx->n, gx->h, gx->w, gx->c, st);
  });
  return rc;
}

int otm_cast(const otm_tensor* x, const otm_tensor* y, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(x && y && x->ptr && y->ptr && same_shape(*x, *y), "cast: bad arguments");
  View xv = make_view(*x), yv = make_view(*y);
  int rc;
  if (x->dtype == OTM_F32 && y->dtype == OTM_F32)
    rc = launch_ew<1>(CastF<float, float>{xv, yv}, x->n, x->h, x->w, x->c, st);
  else if (x->dtype == OTM_F32)
    rc = launch_ew<1>(CastF<float, __nv_bfloat16>{xv, yv}, x->n, x->h, x->w, x->c, st);
  else if (y->dtype == OTM_F32)
    rc = launch_ew<1>(CastF<__nv_bfloat16, float>{xv, yv}, x->n, x->h, x->w, x->c, st);
  else
    rc = launch_ew<1>(CastF<__nv_bfloat16, __nv_bfloat16>{xv, yv}, x->n, x->h, x->w, x->c, st);
  return rc;
}

int otm_add_inplace(const otm_tensor* dst, const otm_tensor* src, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(dst && src && dst->ptr && src->ptr && same_shape(*dst, *src) &&
                  dst->dtype == src->dtype,
              "add_inplace: bad arguments");
  bool vok = vec_ok(*dst, 8) && vec_ok(*src, 8);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(dst->dtype, vok, {
    AddF<T, V> f{make_view(*dst), make_view(*src)};
    rc = launch_ew<V>(f, dst->n, dst->h, dst->w, dst->c, st);
  });
  return rc;
}

}  // extern "C"
