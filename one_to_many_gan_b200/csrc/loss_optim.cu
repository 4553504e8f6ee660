// Loss reductions with fused backward seeds, multi-tensor Adam over flat arenas and the
// Philox synthetic-batch generator.  All HBM-bound: one pass over each operand, 128-bit
// access where the views allow, warp-shuffle + shared-memory block reduction, one atomic
// per CTA and quantity.
#include "common.cuh"

namespace otm {

// ---------------------------------------------------------------------------
// generic scalar reduction driver: F::NQ sums; F::operator()(n,h,w,c0, acc[NQ])
// ---------------------------------------------------------------------------
template <int V, typename F>
__global__ void __launch_bounds__(256) scalar_reduce_kernel(F f, int N, int H, int W, int C,
                                                            float* out) {
  constexpr int NQ = F::NQ;
  const int CV = C / V;
  const long long total = (long long)N * H * W * CV;
  float acc[NQ];
#pragma unroll
  for (int q = 0; q < NQ; ++q) acc[q] = 0.f;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(idx % CV);
    long long p = idx / CV;
    int w = (int)(p % W);
    p /= W;
    int h = (int)(p % H);
    int n = (int)(p / H);
    f(n, h, w, cv * V, acc);
  }
  __shared__ float red[NQ][8];
  const int lane = threadIdx.x % 32, warp = threadIdx.x / 32;
#pragma unroll
  for (int q = 0; q < NQ; ++q) {
    float v = warp_sum(acc[q]);
    if (lane == 0) red[q][warp] = v;
  }
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int q = 0; q < NQ; ++q) {
      float v = lane < 8 ? red[q][lane] : 0.f;
      v = warp_sum(v);
      if (lane == 0) atomicAdd(out + q, v * f.out_scale(q));
    }
  }
}

template <int V, typename F>
static int launch_scalar_reduce(F f, int N, int H, int W, int C, float* out, cudaStream_t st) {
  long long total = (long long)N * H * W * (C / V);
  if (total == 0) return OTM_OK;
  long long blocks = (total + 255) / 256;
  long long cap = (long long)num_sms() * 4;
  if (blocks > cap) blocks = cap;
  scalar_reduce_kernel<V, F><<<(int)blocks, 256, 0, st>>>(f, N, H, W, C, out);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

template <typename T, int V>
struct LsganF {
  static constexpr int NQ = 2;
  View x, grad;
  float target, gscale, inv_n;
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[2]) const {
    float v[V], g[V];
    load_vec<T, V>(vptr<T>(x, n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float d = v[i] - target;
      acc[0] += d * d;
      float sg = 2.f * v[i] - 1.f;
      acc[1] += (sg > 0.f) ? 1.f : (sg < 0.f ? -1.f : 0.f);
      g[i] = gscale * 2.f * d * inv_n;
    }
    if (grad.ptr) store_vec<T, V>(vptr_mut<T>(grad, n, h, w, c), g);
  }
  __device__ float out_scale(int) const { return inv_n; }
};

template <typename T, int V>
struct L1F {
  static constexpr int NQ = 1;
  View a, b, grad;
  float gscale, inv_n;
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[1]) const {
    float x[V], y[V], g[V];
    load_vec<T, V>(vptr<T>(a, n, h, w, c), x);
    load_vec<T, V>(vptr<T>(b, n, h, w, c), y);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float d = x[i] - y[i];
      acc[0] += fabsf(d);
      g[i] = gscale * inv_n * ((d > 0.f) ? 1.f : (d < 0.f ? -1.f : 0.f));
    }
    if (grad.ptr) store_vec<T, V>(vptr_mut<T>(grad, n, h, w, c), g);
  }
  __device__ float out_scale(int) const { return inv_n; }
};

template <typename T, int V>
struct MomentsF {
  static constexpr int NQ = 2;
  View x;
  __device__ void operator()(int n, int h, int w, int c, float (&acc)[2]) const {
    float v[V];
    load_vec<T, V>(vptr<T>(x, n, h, w, c), v);
#pragma unroll
    for (int i = 0; i < V; ++i) {
      acc[0] += v[i];
      acc[1] += v[i] * v[i];
    }
  }
  __device__ float out_scale(int) const { return 1.f; }
};

template <typename T, int V>
struct PathF {
  static constexpr int NQ = 1;
  View f1, f2, g1, g2;
  const float* h;
  float weight, gscale, inv_n;
  __device__ void operator()(int n, int hh, int w, int c, float (&acc)[1]) const {
    float a[V], b[V], ga[V], gb[V];
    load_vec<T, V>(vptr<T>(f1, n, hh, w, c), a);
    load_vec<T, V>(vptr<T>(f2, n, hh, w, c), b);
    const float ih = 1.f / h[n];
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float j = (a[i] - b[i]) * ih;
      acc[0] += j * j;
      ga[i] = gscale * weight * 2.f * j * ih * inv_n;
      gb[i] = -ga[i];
    }
    if (g1.ptr) store_vec<T, V>(vptr_mut<T>(g1, n, hh, w, c), ga);
    if (g2.ptr) store_vec<T, V>(vptr_mut<T>(g2, n, hh, w, c), gb);
  }
  __device__ float out_scale(int) const { return weight * inv_n; }
};

template <typename T, int V>
__global__ void __launch_bounds__(256) affine_grad_kernel(View x, View grad, const float* coef,
                                                          int accumulate, int N, int H, int W,
                                                          int C) {
  const int CV = C / V;
  const long long total = (long long)N * H * W * CV;
  const float c0 = coef[0], c1 = coef[1];
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int cv = (int)(idx % CV);
    long long p = idx / CV;
    int w = (int)(p % W);
    p /= W;
    int h = (int)(p % H);
    int n = (int)(p / H);
    float v[V], g[V];
    load_vec<T, V>(vptr<T>(x, n, h, w, cv * V), v);
    if (accumulate) load_vec_rw<T, V>(vptr<T>(grad, n, h, w, cv * V), g);
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = (accumulate ? g[i] : 0.f) + c0 + c1 * v[i];
    store_vec<T, V>(vptr_mut<T>(grad, n, h, w, cv * V), g);
  }
}

// ---------------------------------------------------------------------------
// Adam
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) adam_kernel(otm_adam_args a) {
  const int t = *a.step;
  const float bc1 = (float)(1.0 - pow((double)a.beta1, (double)t));
  const float bc2_sqrt = (float)sqrt(1.0 - pow((double)a.beta2, (double)t));
  const float step_size = a.lr / bc1;
  const long long n4 = a.n / 4;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 p = reinterpret_cast<float4*>(a.param)[i];
    float4 g = reinterpret_cast<const float4*>(a.grad)[i];
    float4 m = reinterpret_cast<float4*>(a.m)[i];
    float4 v = reinterpret_cast<float4*>(a.v)[i];
    float* pp = &p.x; float* gp = &g.x; float* mp = &m.x; float* vp = &v.x;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      float gg = gp[k] * a.grad_scale;
      mp[k] = a.beta1 * mp[k] + (1.f - a.beta1) * gg;
      vp[k] = a.beta2 * vp[k] + (1.f - a.beta2) * gg * gg;
      float denom = sqrtf(vp[k]) / bc2_sqrt + a.eps;
      pp[k] -= step_size * (mp[k] / denom);
    }
    reinterpret_cast<float4*>(a.param)[i] = p;
    reinterpret_cast<float4*>(a.m)[i] = m;
    reinterpret_cast<float4*>(a.v)[i] = v;
  }
  // tail
  for (long long i = n4 * 4 + (long long)blockIdx.x * blockDim.x + threadIdx.x; i < a.n;
       i += stride) {
    float gg = a.grad[i] * a.grad_scale;
    float m = a.beta1 * a.m[i] + (1.f - a.beta1) * gg;
    float v = a.beta2 * a.v[i] + (1.f - a.beta2) * gg * gg;
    a.m[i] = m; a.v[i] = v;
    a.param[i] -= step_size * (m / (sqrtf(v) / bc2_sqrt + a.eps));
  }
}

// ---------------------------------------------------------------------------
// Philox4x32-10
// ---------------------------------------------------------------------------
__device__ __forceinline__ void philox_round(uint32_t (&c)[4], uint32_t (&k)[2]) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u;
  uint32_t hi0 = __umulhi(M0, c[0]), lo0 = M0 * c[0];
  uint32_t hi1 = __umulhi(M1, c[2]), lo1 = M1 * c[2];
  uint32_t n0 = hi1 ^ c[1] ^ k[0], n1 = lo1, n2 = hi0 ^ c[3] ^ k[1], n3 = lo0;
  c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
  k[0] += 0x9E3779B9u; k[1] += 0xBB67AE85u;
}

__global__ void __launch_bounds__(256) synth_kernel(float* out, long long n, uint64_t seed,
                                                    uint64_t stream_id, uint64_t offset) {
  const long long n4 = (n + 3) / 4;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4;
       i += (long long)gridDim.x * blockDim.x) {
    uint64_t ctr = offset / 4 + (uint64_t)i;
    uint32_t c[4] = {(uint32_t)ctr, (uint32_t)(ctr >> 32), (uint32_t)stream_id,
                     (uint32_t)(stream_id >> 32)};
    uint32_t k[2] = {(uint32_t)seed, (uint32_t)(seed >> 32)};
#pragma unroll
    for (int r = 0; r < 10; ++r) philox_round(c, k);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      long long e = i * 4 + j;
      if (e < n) out[e] = (float)(c[j] >> 8) * (2.0f / 16777216.0f) - 1.0f;
    }
  }
}

// ---------------------------------------------------------------------------
// Device-resident real-data path: the whole (resized) image folder lives in HBM as uint8
// [N, C, H, W] -- what PIL decodes, 1 byte per pixel -- and a batch is ONE gather kernel:
//   out[b] = Normalize(0.5, 0.5)(ToTensor(data[idx[b]])), mirrored in w when flip[b] != 0
// = (v / 255 - 0.5) / 0.5, the transform chain of reference train.py:120-126 plus the random
// horizontal flip of ShoeDataset.__getitem__ (datasets.py:48-50).  HBM-bound: 1 B read + 4 B
// written per pixel; a thread converts 4 consecutive pixels (one 32-bit load, one 128-bit store).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
gather_batch_kernel(const uint8_t* __restrict__ data, const long long* __restrict__ idx,
                    const uint8_t* __restrict__ flip, float* __restrict__ out, int batch, int c,
                    int h, int w, int vec) {
  const long long plane = (long long)h * w, img = plane * c;
  if (vec) {
    const int wq = w / 4;
    const long long total = (long long)batch * c * h * wq;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
      const int xq = (int)(i % wq);
      long long r = i / wq;
      const int y = (int)(r % h);
      r /= h;
      const int ch = (int)(r % c), b = (int)(r / c);
      const bool fl = flip && flip[b];
      const uint8_t* row = data + idx[b] * img + ch * plane + (long long)y * w;
      const uchar4 v = *reinterpret_cast<const uchar4*>(row + (fl ? w - 4 - 4 * xq : 4 * xq));
      const float k = 2.f / 255.f;
      float4 o;
      if (fl) o = make_float4(v.w * k - 1.f, v.z * k - 1.f, v.y * k - 1.f, v.x * k - 1.f);
      else o = make_float4(v.x * k - 1.f, v.y * k - 1.f, v.z * k - 1.f, v.w * k - 1.f);
      *reinterpret_cast<float4*>(out + ((long long)b * c + ch) * plane + (long long)y * w + 4 * xq) = o;
    }
    return;
  }
  const long long total = (long long)batch * img;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total;
       i += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(i % w);
    const long long r = i / w;  // (b * c + ch) * h + y
    const int b = (int)(i / img);
    const bool fl = flip && flip[b];
    const long long src = idx[b] * img + (r - (long long)b * c * h) * w + (fl ? w - 1 - x : x);
    out[i] = data[src] * (2.f / 255.f) - 1.f;
  }
}

}  // namespace otm

using namespace otm;

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

static bool same_shape2(const otm_tensor& a, const otm_tensor& b) {
  return a.n == b.n && a.h == b.h && a.w == b.w && a.c == b.c && a.dtype == b.dtype;
}

extern "C" {

int otm_loss_lsgan(const otm_tensor* x, float target, float scale, float* out,
                   const otm_tensor* grad, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(x && x->ptr && out, "loss_lsgan: null");
  OTM_REQUIRE(!grad || !grad->ptr || same_shape2(*x, *grad), "loss_lsgan: grad mismatch");
  OTM_CHECK_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(float), st));
  const bool hg = grad && grad->ptr;
  bool vok = vec_ok(*x, 8) && (!hg || vec_ok(*grad, 8));
  const float inv_n = 1.f / ((float)x->n * x->h * x->w * x->c);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(x->dtype, vok, {
    LsganF<T, V> f{make_view(*x), hg ? make_view(*grad) : null_view(), target, scale, inv_n};
    rc = launch_scalar_reduce<V>(f, x->n, x->h, x->w, x->c, out, st);
  });
  return rc;
}

int otm_loss_l1(const otm_tensor* a, const otm_tensor* b, float scale, float* out,
                const otm_tensor* grad, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && b && a->ptr && b->ptr && out && same_shape2(*a, *b), "loss_l1: bad arguments");
  OTM_REQUIRE(!grad || !grad->ptr || same_shape2(*a, *grad), "loss_l1: grad mismatch");
  OTM_CHECK_CUDA(cudaMemsetAsync(out, 0, sizeof(float), st));
  const bool hg = grad && grad->ptr;
  bool vok = vec_ok(*a, 8) && vec_ok(*b, 8) && (!hg || vec_ok(*grad, 8));
  const float inv_n = 1.f / ((float)a->n * a->h * a->w * a->c);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(a->dtype, vok, {
    L1F<T, V> f{make_view(*a), make_view(*b), hg ? make_view(*grad) : null_view(), scale, inv_n};
    rc = launch_scalar_reduce<V>(f, a->n, a->h, a->w, a->c, out, st);
  });
  return rc;
}

int otm_moments(const otm_tensor* x, float* out, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(x && x->ptr && out, "moments: null");
  OTM_CHECK_CUDA(cudaMemsetAsync(out, 0, 2 * sizeof(float), st));
  bool vok = vec_ok(*x, 8);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(x->dtype, vok, {
    MomentsF<T, V> f{make_view(*x)};
    rc = launch_scalar_reduce<V>(f, x->n, x->h, x->w, x->c, out, st);
  });
  return rc;
}

int otm_affine_grad(const otm_tensor* x, const float* coef, const otm_tensor* grad,
                    int32_t accumulate, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(x && grad && x->ptr && grad->ptr && coef && same_shape2(*x, *grad),
              "affine_grad: bad arguments");
  bool vok = vec_ok(*x, 8) && vec_ok(*grad, 8);
  long long total = (long long)x->n * x->h * x->w * x->c;
  if (total == 0) return OTM_OK;
  OTM_DISPATCH_TV(x->dtype, vok, {
    long long blocks = (total / V + 255) / 256;
    long long cap = (long long)num_sms() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    affine_grad_kernel<T, V><<<(int)blocks, 256, 0, st>>>(make_view(*x), make_view(*grad), coef,
                                                         accumulate, x->n, x->h, x->w, x->c);
  });
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_loss_path(const otm_tensor* f1, const otm_tensor* f2, const float* h, float weight,
                  float scale, float* out, const otm_tensor* g1, const otm_tensor* g2,
                  otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(f1 && f2 && f1->ptr && f2->ptr && h && out && same_shape2(*f1, *f2),
              "loss_path: bad arguments");
  const bool h1 = g1 && g1->ptr, h2 = g2 && g2->ptr;
  OTM_REQUIRE((!h1 || same_shape2(*f1, *g1)) && (!h2 || same_shape2(*f1, *g2)),
              "loss_path: grad mismatch");
  bool vok = vec_ok(*f1, 8) && vec_ok(*f2, 8) && (!h1 || vec_ok(*g1, 8)) && (!h2 || vec_ok(*g2, 8));
  const float inv_n = 1.f / ((float)f1->n * f1->h * f1->w * f1->c);
  int rc = OTM_OK;
  OTM_DISPATCH_TV(f1->dtype, vok, {
    PathF<T, V> f{make_view(*f1), make_view(*f2), h1 ? make_view(*g1) : null_view(),
                  h2 ? make_view(*g2) : null_view(), h, weight, scale, inv_n};
    rc = launch_scalar_reduce<V>(f, f1->n, f1->h, f1->w, f1->c, out, st);
  });
  return rc;
}

int otm_adam(const otm_adam_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->param && a->grad && a->m && a->v && a->step, "adam: null");
  OTM_REQUIRE(((uintptr_t)a->param | (uintptr_t)a->grad | (uintptr_t)a->m | (uintptr_t)a->v) % 16 == 0,
              "adam: arenas must be 16-byte aligned");
  if (a->n == 0) return OTM_OK;
  long long blocks = (a->n / 4 + 255) / 256;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  adam_kernel<<<(int)blocks, 256, 0, st>>>(*a);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_gather_batch(const uint8_t* data, int64_t n_images, int32_t c, int32_t h, int32_t w,
                     const int64_t* idx, const uint8_t* flip, int32_t batch, float* out,
                     otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(data && idx && out && n_images >= 1 && c >= 1 && h >= 1 && w >= 1 && batch >= 1,
              "gather_batch: bad arguments");
  const int vec = (w % 4 == 0) && ((uintptr_t)data % 4 == 0) && ((uintptr_t)out % 16 == 0);
  const long long items = (long long)batch * c * h * (vec ? w / 4 : w);
  long long blocks = (items + 255) / 256;
  const long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  gather_batch_kernel<<<(int)blocks, 256, 0, st>>>(data, (const long long*)idx, flip, out, batch, c, h,
                                                   w, vec);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_synth_uniform(float* out, int64_t n, uint64_t seed, uint64_t stream_id, uint64_t offset,
                      otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(out && n >= 0 && offset % 4 == 0, "synth_uniform: bad arguments");
  if (n == 0) return OTM_OK;
  long long blocks = ((n + 3) / 4 + 255) / 256;
  long long cap = (long long)num_sms() * 8;
  if (blocks > cap) blocks = cap;
  synth_kernel<<<(int)blocks, 256, 0, st>>>(out, n, seed, stream_id, offset);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

}  // extern "C"
