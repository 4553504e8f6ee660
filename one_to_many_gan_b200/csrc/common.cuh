// Shared device/host helpers for the otm_b200 kernels (sm_100a only).
#pragma once

#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/otm_b200.h"

namespace otm {

// ---------------------------------------------------------------------------
// host-side error plumbing
// ---------------------------------------------------------------------------
extern thread_local char g_err[512];
extern std::atomic<int64_t> g_launches;

inline int fail(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

#define OTM_CHECK_CUDA(expr)                                                             \
  do {                                                                                   \
    cudaError_t _e = (expr);                                                             \
    if (_e != cudaSuccess)                                                               \
      return ::otm::fail(OTM_ERR_CUDA, "%s failed: %s (%s:%d)", #expr,                   \
                         cudaGetErrorString(_e), __FILE__, __LINE__);                    \
  } while (0)

#define OTM_REQUIRE(cond, ...)                                        \
  do {                                                                \
    if (!(cond)) return ::otm::fail(OTM_ERR_INVALID, __VA_ARGS__);    \
  } while (0)

#define OTM_LAUNCH_CHECK()                     \
  do {                                         \
    ::otm::g_launches.fetch_add(1);            \
    OTM_CHECK_CUDA(cudaGetLastError());        \
  } while (0)

// cudaFuncAttributeMaxDynamicSharedMemorySize is a per-DEVICE attribute of a kernel: remember it
// per (call site = kernel instantiation, device), thread-safe.  One bit per device ordinal.
#define OTM_ENSURE_SMEM(kern, bytes)                                                         \
  do {                                                                                        \
    static std::atomic<unsigned long long> done_{0};                                          \
    int dev_ = 0;                                                                             \
    OTM_CHECK_CUDA(cudaGetDevice(&dev_));                                                     \
    const unsigned long long bit_ = 1ull << (dev_ & 63);                                      \
    if (!(done_.load(std::memory_order_acquire) & bit_)) {                                    \
      OTM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,  \
                                          (int)(bytes)));                                     \
      done_.fetch_or(bit_, std::memory_order_release);                                        \
    }                                                                                         \
  } while (0)

inline int num_sms() {
  static std::atomic<int> cached[64];
  int dev = 0;
  cudaGetDevice(&dev);
  int n = cached[dev & 63].load(std::memory_order_relaxed);
  if (n == 0) {
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
    cached[dev & 63].store(n, std::memory_order_relaxed);
  }
  return n;
}

// ---------------------------------------------------------------------------
// device-side tensor view
// ---------------------------------------------------------------------------
struct View {
  char* ptr;
  int dtype;
  int n, h, w, c;
  long long sn, sh, sw;
};

inline View make_view(const otm_tensor& t) {
  View v;
  v.ptr = (char*)t.ptr;
  v.dtype = t.dtype;
  v.n = t.n; v.h = t.h; v.w = t.w; v.c = t.c;
  v.sn = t.sn; v.sh = t.sh; v.sw = t.sw;
  return v;
}
inline View null_view() {
  View v{};
  v.ptr = nullptr;
  return v;
}
__host__ __device__ inline View null_view_dev() {
  View v{};
  v.ptr = nullptr;
  return v;
}
inline size_t dtype_size(int dt) { return dt == OTM_BF16 ? 2 : 4; }

// can the view be accessed with V-wide channel vectors?
inline bool vec_ok(const otm_tensor& t, int V) {
  if (t.ptr == nullptr) return true;
  size_t es = dtype_size(t.dtype);
  size_t align = (V * es >= 16) ? 16 : V * es;
  return (t.c % V == 0) && (t.sn % V == 0) && (t.sh % V == 0) && (t.sw % V == 0) &&
         (((uintptr_t)t.ptr) % align == 0);
}

template <typename T>
__device__ __forceinline__ const T* vptr(const View& v, int n, int h, int w, int c) {
  return (const T*)v.ptr + (n * v.sn + h * v.sh + w * v.sw + c);
}
template <typename T>
__device__ __forceinline__ T* vptr_mut(const View& v, int n, int h, int w, int c) {
  return (T*)v.ptr + (n * v.sn + h * v.sh + w * v.sw + c);
}

__device__ __forceinline__ float to_f(float x) { return x; }
__device__ __forceinline__ float to_f(__nv_bfloat16 x) { return __bfloat162float(x); }
template <typename T> __device__ __forceinline__ T from_f(float x);
template <> __device__ __forceinline__ float from_f<float>(float x) { return x; }
template <> __device__ __forceinline__ __nv_bfloat16 from_f<__nv_bfloat16>(float x) {
  return __float2bfloat16_rn(x);
}

// V-wide (1 or 8) channel vector load/store with fp32 registers.
// (Measured: routing these through ld.global.nc (__ldg) or unrolling the row loops x4 made the
// stencil and reduce passes 10-80 % SLOWER on B200, so plain coherent loads are used.)
template <typename T, int V, bool RO>
__device__ __forceinline__ void load_vec_impl(const T* p, float (&v)[V]) {
  if constexpr (V == 1) {
    v[0] = to_f(RO ? __ldg(p) : *p);
  } else if constexpr (sizeof(T) == 2) {
    static_assert(V == 8, "V must be 1 or 8");
    const uint4* q = reinterpret_cast<const uint4*>(p);
    uint4 raw = RO ? __ldg(q) : *q;
    const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      float2 f = __bfloat1622float2(h2[i]);
      v[2 * i] = f.x;
      v[2 * i + 1] = f.y;
    }
  } else {
    const float4* q = reinterpret_cast<const float4*>(p);
    float4 a = RO ? __ldg(q) : q[0];
    float4 b = RO ? __ldg(q + 1) : q[1];
    v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w;
    v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w;
  }
}
template <typename T, int V>
__device__ __forceinline__ void load_vec(const T* p, float (&v)[V]) { load_vec_impl<T, V, false>(p, v); }
template <typename T, int V>
__device__ __forceinline__ void load_vec_rw(const T* p, float (&v)[V]) { load_vec_impl<T, V, false>(p, v); }
template <typename T, int V>
__device__ __forceinline__ void store_vec(T* p, const float (&v)[V]) {
  if constexpr (V == 1) {
    *p = from_f<T>(v[0]);
  } else if constexpr (sizeof(T) == 2) {
    uint4 raw;
    __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) h2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    *reinterpret_cast<uint4*>(p) = raw;
  } else {
    *reinterpret_cast<float4*>(p) = make_float4(v[0], v[1], v[2], v[3]);
    *reinterpret_cast<float4*>(p + 4) = make_float4(v[4], v[5], v[6], v[7]);
  }
}

// ---------------------------------------------------------------------------
// reflect-halo index helpers (nn.ReflectionPad2d semantics)
// ---------------------------------------------------------------------------
// Positions (in view coordinates, may be <0 or >=n) that mirror onto / from interior
// index i for a reflect halo of width p around an extent-n axis.  Returns the count;
// out[0] is always i itself.
__device__ __forceinline__ int mirror_set(int i, int n, int p, int (&out)[3]) {
  int k = 0;
  out[k++] = i;
  if (p > 0) {
    if (i >= 1 && i <= p) out[k++] = -i;
    if (i >= n - 1 - p && i <= n - 2) out[k++] = 2 * (n - 1) - i;
  }
  return k;
}

template <typename T, int V>
__device__ __forceinline__ void store_halo(const View& y, int halo, int n, int h, int w, int c,
                                           const float (&v)[V]) {
  if (halo == 0 || (h > halo && h < y.h - 1 - halo && w > halo && w < y.w - 1 - halo)) {
    store_vec<T, V>(vptr_mut<T>(y, n, h, w, c), v);
    return;
  }
  int hs[3], ws[3];
  int nh = mirror_set(h, y.h, halo, hs);
  int nw = mirror_set(w, y.w, halo, ws);
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b) store_vec<T, V>(vptr_mut<T>(y, n, hs[a], ws[b], c), v);
}

// gradient w.r.t. the un-padded tensor = sum over all padded positions that alias it
template <typename T, int V>
__device__ __forceinline__ void load_fold(const View& g, int halo, int n, int h, int w, int c,
                                          float (&v)[V]) {
  // interior pixels (all but a 2*halo+1 wide frame) alias nothing
  if (halo == 0 || (h > halo && h < g.h - 1 - halo && w > halo && w < g.w - 1 - halo)) {
    load_vec<T, V>(vptr<T>(g, n, h, w, c), v);
    return;
  }
  int hs[3], ws[3];
  int nh = mirror_set(h, g.h, halo, hs);
  int nw = mirror_set(w, g.w, halo, ws);
#pragma unroll
  for (int i = 0; i < V; ++i) v[i] = 0.f;
  for (int a = 0; a < nh; ++a)
    for (int b = 0; b < nw; ++b) {
      float t[V];
      load_vec<T, V>(vptr<T>(g, n, hs[a], ws[b], c), t);
#pragma unroll
      for (int i = 0; i < V; ++i) v[i] += t[i];
    }
}

// Activations are applied to whole register vectors with ONE (warp-uniform) branch per vector:
// a per-element `switch` inlines tanhf() at every call site, which bloated the tcgen05 epilogue
// past the instruction cache and cost ~100 us per launch (profiles/r1_epilogue_ablation.md).
// tanh as two MUFU ops (ex2 + rcp), fully unrollable: tanhf() is ~40 instructions with branches,
// and keeping it in a rolled loop indexes the value array dynamically, which moved the array of
// EVERY activation path to local memory (STL/LDL round trips in the tcgen05 epilogue).
// |error| <= ~2e-7 absolute.
__device__ __forceinline__ float tanh_fast(float x) {
  const float ax = fabsf(x);
  const float e = __expf(-2.f * ax);
  return copysignf(__fdividef(1.f - e, 1.f + e), x);
}
template <int V>
__device__ __forceinline__ void act_fwd_vec(float (&v)[V], int act) {
  if (act == OTM_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = fmaxf(v[i], 0.f);
  } else if (act == OTM_ACT_LRELU) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = v[i] > 0.f ? v[i] : 0.2f * v[i];
  } else if (act == OTM_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < V; ++i) v[i] = tanh_fast(v[i]);
  }
}
// g[i] *= act'(pre[i])  (derivative given the PRE-activation value)
template <int V>
__device__ __forceinline__ void act_bwd_vec(float (&g)[V], const float (&pre)[V], int act) {
  if (act == OTM_ACT_RELU) {
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = pre[i] > 0.f ? g[i] : 0.f;
  } else if (act == OTM_ACT_LRELU) {
#pragma unroll
    for (int i = 0; i < V; ++i) g[i] = pre[i] > 0.f ? g[i] : 0.2f * g[i];
  } else if (act == OTM_ACT_TANH) {
#pragma unroll
    for (int i = 0; i < V; ++i) {
      float t = tanh_fast(pre[i]);
      g[i] *= 1.f - t * t;
    }
  }
}
__device__ __forceinline__ float act_fwd(float x, int act) {
  float v[1] = {x};
  act_fwd_vec<1>(v, act);
  return v[0];
}

// Software prefetch into L2 for the streaming passes: a warp has only one or two 16-byte loads
// in flight per tensor, far too few bytes to cover HBM latency; prefetching the rows a few
// iterations ahead costs no registers.
template <typename T>
__device__ __forceinline__ void prefetch_l2(const View& v, int n, int h, int w, int c) {
  if (v.ptr) asm volatile("prefetch.global.L2 [%0];" ::"l"(vptr<T>(v, n, h, w, c)));
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// dispatch helpers ----------------------------------------------------------
#define OTM_DISPATCH_DTYPE(dt, T, ...)                     \
  do {                                                     \
    if ((dt) == OTM_BF16) { using T = __nv_bfloat16; __VA_ARGS__; } \
    else { using T = float; __VA_ARGS__; }                 \
  } while (0)

}  // namespace otm
