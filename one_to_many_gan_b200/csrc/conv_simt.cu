// CUDA-core implicit-GEMM convolution kernels (fp32 FFMA accumulation).
//  * the exact-fp32 "parity" mode of every conv (1e-4 tolerance cannot be met by a single
//    bf16/tf32 tensor-core pass), and
//  * the skinny layers in bf16 mode (Cin==1 first layers, Cout==1 last layers) whose GEMM
//    shapes (K=16/49, N=1) do not fill a tcgen05 tile and are HBM-bound anyway.
// The tcgen05 kernels for the dense layers live in conv_tc.cu.
#include "common.cuh"

namespace otm {

struct ConvP {
  View x, y, res;
  const void* w;
  long long w_bstride;
  int x_halo, y_halo, kh, kw, pad;
  int cin, cout, ktot;
  float alpha;
  const float* row_scale;
  const float* bias;
  int act;
  int tiles_w;
  const float* post_scale;  // generic tile kernel only (conv_fwd_simt routes there)
};

struct WgradP {
  View x, dy;
  int x_halo, kh, kw, pad, cin, cout, ktot;
  float* dw;
  float alpha;
  const float* rs;
  const float* cs;
  int splits, pix_per_split;
};

// ---------------------------------------------------------------------------
// generic tile kernel: 64 output pixels (8x8) x 64 output channels per CTA, K chunk 16
// ---------------------------------------------------------------------------
constexpr int TM = 64, TN = 64, TK = 16;

template <typename TI, typename TO>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvP p) {
  __shared__ float As[TK][TM + 4];
  __shared__ float Bs[TK][TN + 4];
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int tile = blockIdx.x;
  const int oh0 = (tile / p.tiles_w) * 8, ow0 = (tile % p.tiles_w) * 8;
  const int o0 = blockIdx.y * TN;
  const int tx = tid % 16, ty = tid / 16;  // 4 couts x 4 pixels per thread
  const TI* wbase = (const TI*)p.w + (long long)n * p.w_bstride;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  for (int k0 = 0; k0 < p.ktot; k0 += TK) {
    // A: 64 pixels x 16 k
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * 256;
      int k = e % TK, m = e / TK;
      int kk = k0 + k;
      float v = 0.f;
      if (kk < p.ktot) {
        int tap = kk / p.cin, c = kk - tap * p.cin;
        int r = tap / p.kw, s = tap - r * p.kw;
        int ih = oh0 + (m >> 3) + r - p.pad, iw = ow0 + (m & 7) + s - p.pad;
        if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo)
          v = to_f(*vptr<TI>(p.x, n, ih, iw, c));
      }
      As[k][m] = v;
    }
    // B: 64 couts x 16 k
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * 256;
      int k = e % TK, o = e / TK;
      int kk = k0 + k;
      float v = 0.f;
      if (kk < p.ktot && o0 + o < p.cout) v = to_f(wbase[(long long)(o0 + o) * p.ktot + kk]);
      Bs[k][o] = v;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }

  // epilogue
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int m = ty * 4 + i;
    int oh = oh0 + (m >> 3), ow = ow0 + (m & 7);
    if (oh >= p.y.h || ow >= p.y.w) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int o = o0 + tx * 4 + j;
      if (o >= p.cout) continue;
      float v = acc[i][j] * p.alpha;
      if (p.row_scale) v *= p.row_scale[(long long)n * p.cout + o];
      if (p.bias) v += p.bias[o];
      v = act_fwd(v, p.act);
      if (p.post_scale) v *= p.post_scale[(long long)n * p.cout + o];
      if (p.res.ptr) v += to_f(*vptr<TO>(p.res, n, oh, ow, o));
      float vv[1] = {v};
      store_halo<TO, 1>(p.y, p.y_halo, n, oh, ow, o, vv);
    }
  }
}

// ---------------------------------------------------------------------------
// skinny-N kernel (Cout <= 4): one thread per output pixel, weights in shared memory.
// Used for the generator's last 7x7 conv (64->1), the discriminator's last conv (512->1)
// and the dgrad of the first layers (-> 1 image channel).
// ---------------------------------------------------------------------------
template <typename TI, typename TO, int V>
__global__ void __launch_bounds__(128) conv_small_cout_kernel(ConvP p) {
  extern __shared__ float wsm[];  // [cout][ktot]
  const int n = blockIdx.z;
  const TI* wbase = (const TI*)p.w + (long long)n * p.w_bstride;
  for (int i = threadIdx.x; i < p.cout * p.ktot; i += blockDim.x) wsm[i] = to_f(wbase[i]);
  __syncthreads();
  const int HoWo = p.y.h * p.y.w;
  const int pix = blockIdx.x * blockDim.x + threadIdx.x;
  if (pix >= HoWo) return;
  const int oh = pix / p.y.w, ow = pix - oh * p.y.w;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  for (int r = 0; r < p.kh; ++r) {
    int ih = oh + r - p.pad;
    if (ih < -halo || ih >= H + halo) continue;
    for (int s = 0; s < p.kw; ++s) {
      int iw = ow + s - p.pad;
      if (iw < -halo || iw >= W + halo) continue;
      const TI* xp = vptr<TI>(p.x, n, ih, iw, 0);
      const float* wp = wsm + (r * p.kw + s) * p.cin;
      for (int c = 0; c < p.cin; c += V) {
        float xv[V];
        load_vec<TI, V>(xp + c, xv);
#pragma unroll
        for (int o = 0; o < 4; ++o) {
          if (o < p.cout) {
#pragma unroll
            for (int i = 0; i < V; ++i) acc[o] = fmaf(xv[i], wp[o * p.ktot + c + i], acc[o]);
          }
        }
      }
    }
  }
  for (int o = 0; o < p.cout; ++o) {
    float v = acc[o] * p.alpha;
    if (p.row_scale) v *= p.row_scale[(long long)n * p.cout + o];
    if (p.bias) v += p.bias[o];
    v = act_fwd(v, p.act);
    if (p.res.ptr) v += to_f(*vptr<TO>(p.res, n, oh, ow, o));
    float vv[1] = {v};
    store_halo<TO, 1>(p.y, p.y_halo, n, oh, ow, o, vv);
  }
}

// ---------------------------------------------------------------------------
// Cout == 1 with many input channels on a small grid (the discriminator head, 512 -> 1 4x4 at
// 13x13): one WARP per output pixel, lanes across the channels (coalesced 512-byte rows of x and
// of the [tap][cin] pack), shuffle reduction.  The pixel-per-thread kernel above walked K = 8192
// serially with 1 KB-strided loads and took 0.24 ms per launch.
// ---------------------------------------------------------------------------
template <typename TI, typename TO>
__global__ void __launch_bounds__(256) conv_cout1_warp_kernel(ConvP p, int total_px) {
  const int lane = threadIdx.x & 31;
  const int gw = (blockIdx.x * 256 + threadIdx.x) >> 5;
  if (gw >= total_px) return;
  const int HoWo = p.y.h * p.y.w;
  const int n = gw / HoWo, pix = gw - n * HoWo;
  const int oh = pix / p.y.w, ow = pix - oh * p.y.w;
  const TI* wbase = (const TI*)p.w + (long long)n * p.w_bstride;
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  float acc = 0.f;
  for (int r = 0; r < p.kh; ++r) {
    const int ih = oh + r - p.pad;
    if (ih < -halo || ih >= H + halo) continue;
    for (int s = 0; s < p.kw; ++s) {
      const int iw = ow + s - p.pad;
      if (iw < -halo || iw >= W + halo) continue;
      const TI* xp = vptr<TI>(p.x, n, ih, iw, 0);
      const TI* wp = wbase + (long long)(r * p.kw + s) * p.cin;
      for (int c = lane * 8; c < p.cin; c += 256) {
        float xv[8], wv[8];
        load_vec<TI, 8>(xp + c, xv);
        load_vec<TI, 8>(wp + c, wv);
#pragma unroll
        for (int i = 0; i < 8; ++i) acc = fmaf(xv[i], wv[i], acc);
      }
    }
  }
  acc = warp_sum(acc);
  if (lane == 0) {
    float v = acc * p.alpha;
    if (p.row_scale) v *= p.row_scale[n];
    if (p.bias) v += p.bias[0];
    v = act_fwd(v, p.act);
    if (p.res.ptr) v += to_f(*vptr<TO>(p.res, n, oh, ow, 0));
    float vv[1] = {v};
    store_halo<TO, 1>(p.y, p.y_halo, n, oh, ow, 0, vv);
  }
}

// ---------------------------------------------------------------------------
// Cout == 1, K x K (K = 4 or 7) with the input patch staged in shared memory:
// persistent CTAs over 16x16 output tiles; a thread owns 8 channels x 4 consecutive pixels and
// slides a 4-vector register window along the filter row, so each filter row costs K+3 patch
// loads + 2K weight loads for 32*K FMAs; the 8 channel-group lanes of a pixel are combined
// with three shuffles.  This is the generator's 64->1 7x7 output conv and the dgrad of the
// 1->64 4x4 input convs (HBM traffic = the input, once).
// ---------------------------------------------------------------------------
template <typename TI, typename TO, int KS>
__global__ void __launch_bounds__(256)
conv_cout1_tiled_kernel(ConvP p, int tiles_w, int tiles_per_img, int total_tiles) {
  constexpr int TT = 16, PW = TT + KS - 1, CC = 64;
  extern __shared__ __align__(16) unsigned char c1_smem[];
  TI* xs = reinterpret_cast<TI*>(c1_smem);                         // [PW*PW][CC]
  float* ws = reinterpret_cast<float*>(xs + (size_t)PW * PW * CC);  // [KS*KS][CC]
  const int cg = threadIdx.x % 8, pg = threadIdx.x / 8;
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  const int nchunks = p.cin / CC;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * TT, ow0 = (tile % tiles_w) * TT;
    const TI* wbase = (const TI*)p.w + (long long)n * p.w_bstride;
    float acc[2][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) acc[a][b] = 0.f;
    for (int ch = 0; ch < nchunks; ++ch) {
      const int c0 = ch * CC;
      __syncthreads();
      for (int e = threadIdx.x; e < KS * KS * CC; e += 256) {
        int c = e % CC, tap = e / CC;
        ws[e] = to_f(wbase[(long long)tap * p.cin + c0 + c]);
      }
      for (int e = threadIdx.x; e < PW * PW * 8; e += 256) {
        int v = e % 8, q = e / 8;
        int pw = q % PW, ph = q / PW;
        int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
        float tmp[8];
        if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo) {
          load_vec<TI, 8>(vptr<TI>(p.x, n, ih, iw, c0 + v * 8), tmp);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) tmp[i] = 0.f;
        }
        store_vec<TI, 8>(xs + (size_t)q * CC + v * 8, tmp);
      }
      __syncthreads();
#pragma unroll
      for (int pass = 0; pass < 2; ++pass) {
        const int g = pass * 32 + pg;
        const int row = g / 4, col4 = (g % 4) * 4;
#pragma unroll 1
        for (int r = 0; r < KS; ++r) {
          const TI* xrow = xs + ((size_t)(row + r) * PW + col4) * CC + cg * 8;
          float xw[4][8];
#pragma unroll
          for (int j = 0; j < 4; ++j) load_vec_rw<TI, 8>(xrow + j * CC, xw[j]);
#pragma unroll
          for (int s2 = 0; s2 < KS; ++s2) {
            float wv[8];
            load_vec_rw<float, 8>(ws + (r * KS + s2) * CC + cg * 8, wv);
#pragma unroll
            for (int px = 0; px < 4; ++px) {
              const float* xv = xw[(s2 + px) % 4];
#pragma unroll
              for (int i = 0; i < 8; ++i) acc[pass][px] = fmaf(xv[i], wv[i], acc[pass][px]);
            }
            if (s2 + 1 < KS) load_vec_rw<TI, 8>(xrow + (s2 + 4) * CC, xw[s2 % 4]);
          }
        }
      }
    }
    // combine the 8 channel-group lanes of each pixel group, then lanes 0..3 write one pixel each
#pragma unroll
    for (int pass = 0; pass < 2; ++pass) {
#pragma unroll
      for (int px = 0; px < 4; ++px) {
        float v = acc[pass][px];
        v += __shfl_xor_sync(0xffffffffu, v, 1);
        v += __shfl_xor_sync(0xffffffffu, v, 2);
        v += __shfl_xor_sync(0xffffffffu, v, 4);
        acc[pass][px] = v;
      }
      if (cg < 4) {
        const int g = pass * 32 + pg;
        const int oh = oh0 + g / 4, ow = ow0 + (g % 4) * 4 + cg;
        if (oh < p.y.h && ow < p.y.w) {
          float v = (cg == 0 ? acc[pass][0] : cg == 1 ? acc[pass][1] : cg == 2 ? acc[pass][2] : acc[pass][3]);
          v *= p.alpha;
          if (p.row_scale) v *= p.row_scale[n];
          if (p.bias) v += p.bias[0];
          v = act_fwd(v, p.act);
          if (p.res.ptr) v += to_f(*vptr<TO>(p.res, n, oh, ow, 0));
          float vv[1] = {v};
          store_halo<TO, 1>(p.y, p.y_halo, n, oh, ow, 0, vv);
        }
      }
    }
  }
}

template <typename TI, typename TO, int KS>
static int launch_cout1_tiled(const ConvP& p, int n, cudaStream_t st) {
  constexpr int TT = 16, PW = TT + KS - 1, CC = 64;
  const size_t smem = (size_t)PW * PW * CC * sizeof(TI) + (size_t)KS * KS * CC * sizeof(float);
  const int tiles_w = (p.y.w + TT - 1) / TT, tiles_h = (p.y.h + TT - 1) / TT;
  const int per_img = tiles_w * tiles_h, total = per_img * n;
  auto kern = conv_cout1_tiled_kernel<TI, TO, KS>;
  OTM_ENSURE_SMEM(kern, (int)smem);
  int ctas = num_sms() * (smem > 110 * 1024 ? 1 : 2);
  if (ctas > total) ctas = total;
  kern<<<ctas, 256, smem, st>>>(p, tiles_w, per_img, total);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

// ---------------------------------------------------------------------------
// Cin == 1 (the image-side 7x7 / 4x4 convs, Cout = 64): FFMA-bound, K = 49 / 16 is far below a
// tensor-core tile.  One thread per output pixel holds all 64 accumulators; the image patch and
// the weights sit in shared memory (weights read as broadcast 128-bit loads).
// ---------------------------------------------------------------------------
template <typename TO, int KS>
__global__ void __launch_bounds__(256)
conv_cin1_kernel(ConvP p, int tiles_w, int tiles_per_img, int total_tiles) {
  constexpr int TT = 16, PW = TT + KS - 1, CO = 64;
  __shared__ float xs[PW * PW];
  __shared__ __align__(16) float ws[KS * KS * CO];  // [tap][cout]
  const int tx = threadIdx.x % TT, ty = threadIdx.x / TT;
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  const float* wbase = (const float*)p.w;
  for (int e = threadIdx.x; e < KS * KS * CO; e += 256) {
    int o = e % CO, tap = e / CO;
    ws[e] = wbase[o * KS * KS + tap];
  }
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * TT, ow0 = (tile % tiles_w) * TT;
    __syncthreads();
    for (int e = threadIdx.x; e < PW * PW; e += 256) {
      int pw = e % PW, ph = e / PW;
      int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
      float v = 0.f;
      if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo)
        v = *vptr<float>(p.x, n, ih, iw, 0);
      xs[e] = v;
    }
    __syncthreads();
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; ++o) acc[o] = 0.f;
#pragma unroll 1
    for (int r = 0; r < KS; ++r) {
#pragma unroll
      for (int s2 = 0; s2 < KS; ++s2) {
        const float xv = xs[(ty + r) * PW + tx + s2];
        const float4* w4 = reinterpret_cast<const float4*>(ws + (r * KS + s2) * CO);
#pragma unroll
        for (int o4 = 0; o4 < CO / 4; ++o4) {
          float4 w = w4[o4];
          acc[4 * o4] = fmaf(xv, w.x, acc[4 * o4]);
          acc[4 * o4 + 1] = fmaf(xv, w.y, acc[4 * o4 + 1]);
          acc[4 * o4 + 2] = fmaf(xv, w.z, acc[4 * o4 + 2]);
          acc[4 * o4 + 3] = fmaf(xv, w.w, acc[4 * o4 + 3]);
        }
      }
    }
    const int oh = oh0 + ty, ow = ow0 + tx;
    if (oh < p.y.h && ow < p.y.w) {
#pragma unroll
      for (int g = 0; g < CO / 8; ++g) {
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          float a = acc[g * 8 + j] * p.alpha;
          if (p.row_scale) a *= p.row_scale[(long long)n * CO + g * 8 + j];
          if (p.bias) a += p.bias[g * 8 + j];
          v[j] = a;
        }
        act_fwd_vec<8>(v, p.act);
        if (p.res.ptr) {
          float rr[8];
          load_vec<TO, 8>(vptr<TO>(p.res, n, oh, ow, g * 8), rr);
#pragma unroll
          for (int j = 0; j < 8; ++j) v[j] += rr[j];
        }
        store_halo<TO, 8>(p.y, p.y_halo, n, oh, ow, g * 8, v);
      }
    }
  }
}

// wgrad of the same layers: dw[o][tap] = sum_px dy[px][o] * x[px + tap].  A thread owns one tap
// and 8 output channels; per pixel it needs one image value (broadcast) and one 128-bit dy load.
template <typename TDY, int KS>
__global__ void __launch_bounds__(KS * KS * 8 <= 128 ? 128 : 416)
wgrad_cin1_kernel(WgradP p, int tiles_w, int tiles_per_img, int total_tiles) {
  constexpr int TT = 16, PW = TT + KS - 1, CO = 64;
  constexpr int NT = KS * KS * 8 <= 128 ? 128 : 416;
  extern __shared__ __align__(16) unsigned char wc1_smem[];
  TDY* dys = reinterpret_cast<TDY*>(wc1_smem);                                  // [256 px][64]
  float* xs = reinterpret_cast<float*>(wc1_smem + sizeof(TDY) * TT * TT * CO);  // [PW*PW]
  const int tap = threadIdx.x / 8, og = threadIdx.x % 8;
  const bool active = tap < KS * KS;
  const int r = tap / KS, s2 = tap % KS;
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * TT, ow0 = (tile % tiles_w) * TT;
    __syncthreads();
    for (int e = threadIdx.x; e < PW * PW; e += NT) {
      int pw = e % PW, ph = e / PW;
      int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
      float v = 0.f;
      if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo)
        v = *vptr<float>(p.x, n, ih, iw, 0);
      xs[e] = v;
    }
    for (int e = threadIdx.x; e < TT * TT * 8; e += NT) {
      int v8 = e % 8, q = e / 8;
      int oh = oh0 + q / TT, ow = ow0 + q % TT;
      float tmp[8];
      if (oh < p.dy.h && ow < p.dy.w) {
        load_vec<TDY, 8>(vptr<TDY>(p.dy, n, oh, ow, v8 * 8), tmp);
      } else {
#pragma unroll
        for (int j = 0; j < 8; ++j) tmp[j] = 0.f;
      }
      store_vec<TDY, 8>(dys + q * CO + v8 * 8, tmp);
    }
    __syncthreads();
    if (active) {
      for (int qh = 0; qh < TT; ++qh) {
#pragma unroll 4
        for (int qw = 0; qw < TT; ++qw) {
          const float xv = xs[(qh + r) * PW + qw + s2];
          float d[8];
          load_vec_rw<TDY, 8>(dys + (qh * TT + qw) * CO + og * 8, d);
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(xv, d[j], acc[j]);
        }
      }
    }
  }
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; ++j) atomicAdd(p.dw + (og * 8 + j) * KS * KS + tap, acc[j] * p.alpha);
  }
}

// ---------------------------------------------------------------------------
// Cin == 1 layers in bf16 mode on the warp-level tensor-core path (mma.sync m16n8k16, bf16 x
// bf16 -> fp32).  K = 49 / 16 taps is below a tcgen05 tile and these layers are bound by the
// 64-channel tensor they write (fwd) or read (wgrad), so the point of the MMA is only to take
// the 64 x taps FMAs per pixel off the FFMA pipe: the FFMA kernels above ran at 19 TFLOP/s
// (0.35 ms per 7x7 launch, 16x the time of the HBM traffic).
//   forward : M = 16 pixels of one tile row, N = 64 couts (8 n-tiles), K = taps (padded to 16s).
//             A fragments are gathered from the bf16 image patch in shared memory, the weight
//             (B) fragments live in registers for the whole kernel.
//   wgrad   : M = taps, N = 64 couts, K = pixels.  A gathered from the image patch, B =
//             ldmatrix.trans of the [pixel][cout] dy tile (row pitch 144 B: conflict-free).
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mma_16816(float (&d)[4], const uint32_t (&a)[4],
                                          const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}
__device__ __forceinline__ uint32_t pack_bf16(__nv_bfloat16 lo, __nv_bfloat16 hi) {
  return (uint32_t)__bfloat16_as_ushort(lo) | ((uint32_t)__bfloat16_as_ushort(hi) << 16);
}
__device__ __forceinline__ uint32_t pack_bf16f(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int KS>
__global__ void __launch_bounds__(256, 2)
conv_cin1_mma_kernel(ConvP p, int tiles_w, int tiles_per_img, int total_tiles) {
  constexpr int TT = 16, PW = TT + KS - 1, CO = 64, TAPS = KS * KS, KSTEPS = (TAPS + 15) / 16;
  constexpr int SP = CO + 8;  // staging row pitch (elements): 144 B, conflict-free
  __shared__ __nv_bfloat16 xs[(PW * PW + 7) & ~7];
  __shared__ __align__(16) __nv_bfloat16 stage[8][16][SP];
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int g = lane >> 2, q = lane & 3;
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  // weight fragments: B[k = tap][n = cout] = w[cout][tap], zero for the padded taps
  uint32_t bf[KSTEPS][8][2];
  {
    const float* wbase = (const float*)p.w;
#pragma unroll
    for (int kk = 0; kk < KSTEPS; ++kk)
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int k0 = kk * 16 + q * 2 + h * 8, o = j * 8 + g;
          const float lo = k0 < TAPS ? wbase[o * TAPS + k0] : 0.f;
          const float hi = k0 + 1 < TAPS ? wbase[o * TAPS + k0 + 1] : 0.f;
          bf[kk][j][h] = pack_bf16f(lo, hi);
        }
  }
  // patch offsets of this lane's four k columns per k-step (padded taps read offset 0: their
  // weights are zero and the patch holds finite values)
  int aoff[KSTEPS][4];
#pragma unroll
  for (int kk = 0; kk < KSTEPS; ++kk)
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      const int k = kk * 16 + q * 2 + (c & 1) + (c >> 1) * 8;
      aoff[kk][c] = k < TAPS ? (k / KS) * PW + (k % KS) : 0;
    }
  __shared__ float s_bias[CO];
  if (threadIdx.x < CO) s_bias[threadIdx.x] = p.bias ? p.bias[threadIdx.x] : 0.f;

  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * TT, ow0 = (tile % tiles_w) * TT;
    __syncthreads();
    for (int e = threadIdx.x; e < PW * PW; e += 256) {
      const int pw = e % PW, ph = e / PW;
      const int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
      float v = 0.f;
      if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo)
        v = *vptr<float>(p.x, n, ih, iw, 0);
      xs[e] = __float2bfloat16_rn(v);
    }
    __syncthreads();
#pragma unroll 1
    for (int mt = 0; mt < 2; ++mt) {
      const int ty = warp * 2 + mt;  // tile row = the 16 pixels (M) of this m-tile
      float acc[8][4];
#pragma unroll
      for (int j = 0; j < 8; ++j)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;
      const int b0 = ty * PW + g, b1 = b0 + 8;
#pragma unroll
      for (int kk = 0; kk < KSTEPS; ++kk) {
        uint32_t a[4];
        a[0] = pack_bf16(xs[b0 + aoff[kk][0]], xs[b0 + aoff[kk][1]]);
        a[1] = pack_bf16(xs[b1 + aoff[kk][0]], xs[b1 + aoff[kk][1]]);
        a[2] = pack_bf16(xs[b0 + aoff[kk][2]], xs[b0 + aoff[kk][3]]);
        a[3] = pack_bf16(xs[b1 + aoff[kk][2]], xs[b1 + aoff[kk][3]]);
#pragma unroll
        for (int j = 0; j < 8; ++j) mma_16816(acc[j], a, bf[kk][j]);
      }
      // epilogue: alpha, bias, activation -> bf16 -> per-warp staging tile [16 px][64 ch]
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = fmaf(acc[j][e], p.alpha, s_bias[j * 8 + q * 2 + (e & 1)]);
        act_fwd_vec<4>(v, p.act);
        *reinterpret_cast<uint32_t*>(&stage[warp][g][j * 8 + q * 2]) = pack_bf16f(v[0], v[1]);
        *reinterpret_cast<uint32_t*>(&stage[warp][g + 8][j * 8 + q * 2]) = pack_bf16f(v[2], v[3]);
      }
      __syncwarp();
      const int oh = oh0 + ty;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int chunk = lane + 32 * i;  // 16 px x 8 chunks of 8 channels
        const int px = chunk >> 3, c8 = chunk & 7;
        const int ow = ow0 + px;
        if (oh < p.y.h && ow < p.y.w) {
          const uint4 val = *reinterpret_cast<const uint4*>(&stage[warp][px][c8 * 8]);
          if (p.y_halo == 0) {
            *reinterpret_cast<uint4*>(vptr_mut<__nv_bfloat16>(p.y, n, oh, ow, c8 * 8)) = val;
          } else {
            float f[8];
            const __nv_bfloat162* h2 = reinterpret_cast<const __nv_bfloat162*>(&val);
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 ff = __bfloat1622float2(h2[e]);
              f[2 * e] = ff.x; f[2 * e + 1] = ff.y;
            }
            store_halo<__nv_bfloat16, 8>(p.y, p.y_halo, n, oh, ow, c8 * 8, f);
          }
        }
      }
      __syncwarp();
    }
  }
}

__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <int KS>
__global__ void __launch_bounds__(256, 2)
wgrad_cin1_mma_kernel(WgradP p, int tiles_w, int tiles_per_img, int total_tiles) {
  constexpr int TT = 16, PW = TT + KS - 1, CO = 64, TAPS = KS * KS, MT = (TAPS + 15) / 16;
  constexpr int KSL = MT == 4 ? 1 : 4;   // K slices (tile rows) across warps
  constexpr int NG = 8 / (MT * KSL);     // n-groups of 4 n-tiles across warps
  static_assert(MT * KSL * NG == 8 && NG == 2, "warp layout");
  constexpr int DP = CO + 8;             // dy row pitch (elements): 144 B
  extern __shared__ __align__(16) unsigned char wm_smem[];
  __nv_bfloat16* dys = reinterpret_cast<__nv_bfloat16*>(wm_smem);  // [256 px][DP]
  __nv_bfloat16* xs = dys + TT * TT * DP;                          // [PW*PW]
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int g = lane >> 2, q = lane & 3;
  const int mi = warp % MT, ng = (warp / MT) % NG, ks = warp / (MT * NG);
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  int toff[2];
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    const int tap = mi * 16 + g + h * 8;
    toff[h] = tap < TAPS ? (tap / KS) * PW + (tap % KS) : 0;
  }
  float acc[4][4];
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[j][e] = 0.f;

  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * TT, ow0 = (tile % tiles_w) * TT;
    __syncthreads();
    for (int e = threadIdx.x; e < PW * PW; e += 256) {
      const int pw = e % PW, ph = e / PW;
      const int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
      float v = 0.f;
      if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo)
        v = *vptr<float>(p.x, n, ih, iw, 0);
      xs[e] = __float2bfloat16_rn(v);
    }
    for (int e = threadIdx.x; e < TT * TT * 8; e += 256) {
      const int c8 = e & 7, px = e >> 3;
      const int oh = oh0 + px / TT, ow = ow0 + px % TT;
      uint4 val = make_uint4(0, 0, 0, 0);
      if (oh < p.dy.h && ow < p.dy.w)
        val = *reinterpret_cast<const uint4*>(vptr<__nv_bfloat16>(p.dy, n, oh, ow, c8 * 8));
      *reinterpret_cast<uint4*>(dys + px * DP + c8 * 8) = val;
    }
    __syncthreads();
#pragma unroll 4
    for (int kq = 0; kq < TT / KSL; ++kq) {
      const int kk = ks * (TT / KSL) + kq;  // tile row = 16 pixels of K
      uint32_t a[4];
      const int xb = kk * PW + q * 2;
      a[0] = pack_bf16(xs[xb + toff[0]], xs[xb + toff[0] + 1]);
      a[1] = pack_bf16(xs[xb + toff[1]], xs[xb + toff[1] + 1]);
      a[2] = pack_bf16(xs[xb + toff[0] + 8], xs[xb + toff[0] + 9]);
      a[3] = pack_bf16(xs[xb + toff[1] + 8], xs[xb + toff[1] + 9]);
#pragma unroll
      for (int jp = 0; jp < 2; ++jp) {  // two n-tiles per ldmatrix.x4
        uint32_t b[4];
        const int krow = kk * 16 + ((lane >> 3) & 1) * 8 + (lane & 7);
        const int ncol = (ng * 4 + jp * 2 + (lane >> 4)) * 8;
        ldmatrix_x4_trans(b, dys + krow * DP + ncol);
        const uint32_t b0[2] = {b[0], b[1]}, b1[2] = {b[2], b[3]};
        mma_16816(acc[jp * 2], a, b0);
        mma_16816(acc[jp * 2 + 1], a, b1);
      }
    }
  }
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int tap = mi * 16 + g + (e >> 1) * 8;
      const int o = (ng * 4 + j) * 8 + q * 2 + (e & 1);
      if (tap < TAPS) atomicAdd(p.dw + o * TAPS + tap, acc[j][e] * p.alpha);
    }
}

// ---------------------------------------------------------------------------
// Cout == 1, Cin == 64 layers (the generator's 7x7 output conv, the dgrad of the 4x4 image-side
// convs) in bf16 mode on mma.sync.  A single output channel leaves N = 1, so the filter's
// HORIZONTAL taps become the N dimension:
//     z[q][s] = sum_{r, ch} X[row + r][q][ch] * w[r][s][ch]      (M = 16 patch columns q,
//     y[px]   = sum_s z[px + s][s]                                 N = 8 >= KS, K = KS * 64)
// A fragments = ldmatrix of the [pixel][channel] bf16 patch (row pitch 144 B, conflict-free),
// each patch-row fragment is reused by up to KS output rows; the weight fragments stay in
// registers.  wgrad is the transpose: D[ch][s] += sum_q X[row + r][q][ch] * dy[row][q - s]
// (A = ldmatrix.trans of the same patch, B gathered from a zero-padded dy row).
// Tile: 16 output rows x (16*MW - KS + 1) output columns, patch (16 + KS - 1) x 16*MW pixels.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}

template <int KS, int MW>
struct Cout1Geom {
  static constexpr int TH = 16, PH = TH + KS - 1, PWC = 16 * MW, TWO = PWC - KS + 1, CP = 72;
  static constexpr int PATCH_ELEMS = (PH * PWC + 16) * CP;  // + 16 zero rows: ldmatrix overrun
};

// stage the (PH x PWC) x 64-channel patch of tile (oh0, ow0) into shared memory (zero outside)
template <int KS, int MW>
__device__ __forceinline__ void cout1_load_patch(const View& x, int x_halo, int pad, int n, int oh0,
                                                 int ow0, __nv_bfloat16* xs) {
  using G = Cout1Geom<KS, MW>;
  const int H = x.h, W = x.w;
  for (int e = threadIdx.x; e < G::PH * G::PWC * 8; e += 256) {
    const int c8 = e & 7, q = e >> 3;
    const int pw = q % G::PWC, ph = q / G::PWC;
    const int ih = oh0 + ph - pad, iw = ow0 + pw - pad;
    uint4 val = make_uint4(0, 0, 0, 0);
    if (ih >= -x_halo && ih < H + x_halo && iw >= -x_halo && iw < W + x_halo)
      val = *reinterpret_cast<const uint4*>(vptr<__nv_bfloat16>(x, n, ih, iw, c8 * 8));
    *reinterpret_cast<uint4*>(xs + q * G::CP + c8 * 8) = val;
  }
}

template <typename TO, int KS, int MW>
__global__ void __launch_bounds__(256, MW == 1 ? 3 : 1)
conv_cout1_mma_kernel(ConvP p, int tiles_w, int tiles_per_img, int total_tiles) {
  using G = Cout1Geom<KS, MW>;
  extern __shared__ __align__(16) unsigned char c1m_smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(c1m_smem);
  float* zs = reinterpret_cast<float*>(xs + G::PATCH_ELEMS);  // [8 warps][2 rows][PWC][8]
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int g = lane >> 2, q4 = lane & 3;
  for (int e = threadIdx.x; e < 16 * G::CP; e += 256) xs[G::PH * G::PWC * G::CP + e] = __float2bfloat16_rn(0.f);
  // weight fragments: B[k = (r, ch)][n = s] = w[(r, s)][ch]; columns s >= KS are zero
  uint32_t bf[KS][4][2];
  {
    const __nv_bfloat16* wbase = (const __nv_bfloat16*)p.w;  // [tap][cin]
#pragma unroll
    for (int r = 0; r < KS; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c)
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          const int ch = c * 16 + q4 * 2 + h * 8;
          uint32_t v = 0;
          if (g < KS) v = *reinterpret_cast<const uint32_t*>(wbase + (r * KS + g) * 64 + ch);
          bf[r][c][h] = v;
        }
  }
  float* zw = zs + warp * (2 * G::PWC * 8);
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * G::TH, ow0 = (tile % tiles_w) * G::TWO;
    __syncthreads();
    cout1_load_patch<KS, MW>(p.x, p.x_halo, p.pad, n, oh0, ow0, xs);
    __syncthreads();
    float acc[2][MW][4];
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int b = 0; b < MW; ++b)
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;
    const int py0 = warp * 2;  // this warp's two output rows
#pragma unroll
    for (int rr = 0; rr < KS + 1; ++rr) {  // patch rows py0 .. py0 + KS
      const int R = py0 + rr;
#pragma unroll
      for (int mt = 0; mt < MW; ++mt)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          uint32_t a[4];
          const int row = R * G::PWC + mt * 16 + (lane & 7) + ((lane >> 3) & 1) * 8;
          ldmatrix_x4(a, xs + row * G::CP + c * 16 + (lane >> 4) * 8);
          if (rr < KS) mma_16816(acc[0][mt], a, bf[rr < KS ? rr : 0][c]);          // row py0,   r = rr
          if (rr >= 1) mma_16816(acc[1][mt], a, bf[rr >= 1 ? rr - 1 : 0][c]);      // row py0+1, r = rr-1
        }
    }
    // z -> shared, then y[px] = sum_s z[px + s][s]
#pragma unroll
    for (int a = 0; a < 2; ++a)
#pragma unroll
      for (int mt = 0; mt < MW; ++mt) {
        float* zr = zw + (a * G::PWC + mt * 16) * 8;
        *reinterpret_cast<float2*>(zr + g * 8 + q4 * 2) = make_float2(acc[a][mt][0], acc[a][mt][1]);
        *reinterpret_cast<float2*>(zr + (g + 8) * 8 + q4 * 2) = make_float2(acc[a][mt][2], acc[a][mt][3]);
      }
    __syncwarp();
#pragma unroll
    for (int a = 0; a < 2; ++a) {
      const int oh = oh0 + py0 + a;
      for (int px = lane; px < G::TWO; px += 32) {
        const int ow = ow0 + px;
        if (oh < p.y.h && ow < p.y.w) {
          float v = 0.f;
#pragma unroll
          for (int sx = 0; sx < KS; ++sx) v += zw[(a * G::PWC + px + sx) * 8 + sx];
          v *= p.alpha;
          if (p.row_scale) v *= p.row_scale[n];
          if (p.bias) v += p.bias[0];
          v = act_fwd(v, p.act);
          if (p.res.ptr) v += to_f(*vptr<TO>(p.res, n, oh, ow, 0));
          float vv[1] = {v};
          store_halo<TO, 1>(p.y, p.y_halo, n, oh, ow, 0, vv);
        }
      }
    }
    __syncwarp();
  }
}

template <typename TDY, int KS, int MW>
__global__ void __launch_bounds__(256, MW == 1 ? 3 : 1)
wgrad_cout1_mma_kernel(WgradP p, int tiles_w, int tiles_per_img, int total_tiles) {
  using G = Cout1Geom<KS, MW>;
  constexpr int DYP = G::PWC + 16;  // dy row: 8 zeros | TWO values | zeros
  constexpr int RH = (KS + 1) / 2;  // filter rows per warp half
  extern __shared__ __align__(16) unsigned char w1m_smem[];
  __nv_bfloat16* xs = reinterpret_cast<__nv_bfloat16*>(w1m_smem);
  __nv_bfloat16* dys = xs + G::PATCH_ELEMS;  // [16 rows][DYP]
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int g = lane >> 2, q4 = lane & 3;
  const int ct = warp & 3;         // 16-channel tile (M)
  const int r0 = (warp >> 2) * RH;  // this warp's filter rows r0 .. r0 + RH - 1
  for (int e = threadIdx.x; e < 16 * G::CP; e += 256) xs[G::PH * G::PWC * G::CP + e] = __float2bfloat16_rn(0.f);
  float acc[RH][4];
#pragma unroll
  for (int a = 0; a < RH; ++a)
#pragma unroll
    for (int e = 0; e < 4; ++e) acc[a][e] = 0.f;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * G::TH, ow0 = (tile % tiles_w) * G::TWO;
    __syncthreads();
    cout1_load_patch<KS, MW>(p.x, p.x_halo, p.pad, n, oh0, ow0, xs);
    for (int e = threadIdx.x; e < 16 * DYP; e += 256) {
      const int py = e / DYP, c = e % DYP - 8;
      float v = 0.f;
      const int oh = oh0 + py, ow = ow0 + c;
      if (c >= 0 && c < G::TWO && oh < p.dy.h && ow < p.dy.w) v = to_f(*vptr<TDY>(p.dy, n, oh, ow, 0));
      dys[e] = __float2bfloat16_rn(v);
    }
    __syncthreads();
#pragma unroll 1
    for (int R = 0; R < G::PH; ++R) {
#pragma unroll
      for (int mt = 0; mt < MW; ++mt) {
        // A[m = ch][k = q]: transposed load of patch rows q (k) x channels (m)
        uint32_t a[4];
        const int row = R * G::PWC + mt * 16 + (lane & 7) + (lane >> 4) * 8;
        ldmatrix_x4_trans(a, xs + row * G::CP + ct * 16 + ((lane >> 3) & 1) * 8);
#pragma unroll
        for (int rl = 0; rl < RH; ++rl) {
          const int r = r0 + rl, py = R - r;
          if (r < KS && py >= 0 && py < G::TH) {
            // B[k = q][n = s] = dy[py][q - s]
            const __nv_bfloat16* d = dys + py * DYP + 8 + mt * 16 + q4 * 2 - g;
            uint32_t b[2];
            b[0] = pack_bf16(d[0], d[1]);
            b[1] = pack_bf16(d[8], d[9]);
            mma_16816(acc[rl], a, b);
          }
        }
      }
    }
  }
  const int taps = KS * KS;
#pragma unroll
  for (int rl = 0; rl < RH; ++rl)
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int r = r0 + rl, sx = q4 * 2 + (e & 1), ch = ct * 16 + g + (e >> 1) * 8;
      if (r < KS && sx < KS) atomicAdd(p.dw + ch * taps + r * KS + sx, acc[rl][e] * p.alpha);
    }
}

// ---------------------------------------------------------------------------
// wgrad, generic tile: M = 64 couts, N = 64 flattened (r,s,i), K = pixels of one sample
// chunk.  grid = (ceil(ktot/64), ceil(cout/64), n * splits); atomicAdd into dw.
// ---------------------------------------------------------------------------

template <typename T, typename TDY>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradP p) {
  __shared__ float As[TK][TM + 4];  // dy: [pixel][cout]
  __shared__ float Bs[TK][TN + 4];  // x : [pixel][k]
  __shared__ int koff_r[TN], koff_s[TN], koff_c[TN];
  const int tid = threadIdx.x;
  const int n = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int o0 = blockIdx.y * TM, kbase = blockIdx.x * TN;
  const int tx = tid % 16, ty = tid / 16;
  if (tid < TN) {
    int kk = kbase + tid;
    int tap = kk / p.cin, c = kk - tap * p.cin;
    int r = tap / p.kw, s = tap - r * p.kw;
    koff_r[tid] = r - p.pad; koff_s[tid] = s - p.pad; koff_c[tid] = (kk < p.ktot) ? c : -1;
  }
  __syncthreads();
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
  const int Ho = p.dy.h, Wo = p.dy.w, HW = Ho * Wo;
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  const int pbeg = split * p.pix_per_split, pend = min(HW, pbeg + p.pix_per_split);
  for (int p0 = pbeg; p0 < pend; p0 += TK) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int e = tid + j * 256;
      int m = e % TM, k = e / TM;  // consecutive threads -> consecutive channels
      int pix = p0 + k;
      float a = 0.f, b = 0.f;
      if (pix < pend) {
        int oh = pix / Wo, ow = pix - oh * Wo;
        if (o0 + m < p.cout) a = to_f(*vptr<TDY>(p.dy, n, oh, ow, o0 + m));
        int c = koff_c[m];
        if (c >= 0) {
          int ih = oh + koff_r[m], iw = ow + koff_s[m];
          if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo)
            b = to_f(*vptr<T>(p.x, n, ih, iw, c));
        }
      }
      As[k][m] = a;
      Bs[k][m] = b;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < TK; ++k) {
      float a[4], b[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
      for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
  const int taps = p.kh * p.kw;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    int o = o0 + ty * 4 + i;
    if (o >= p.cout) continue;
    float rsv = p.rs ? p.rs[(long long)n * p.cout + o] : 1.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      int kk = kbase + tx * 4 + j;
      if (kk >= p.ktot) continue;
      int tap = kk / p.cin, c = kk - tap * p.cin;
      float v = acc[i][j] * p.alpha * rsv;
      if (p.cs) v *= p.cs[(long long)n * p.cin + c];
      atomicAdd(p.dw + ((long long)o * p.cin + c) * taps + tap, v);
    }
  }
}

// ---------------------------------------------------------------------------
// wgrad for skinny outputs (Cout <= 4): one CTA per (sample, TT x TT output tile).  The input
// patch and the dy tile are staged in shared memory once; each thread owns up to WG_MAXI
// (tap, channel, cout) accumulators, reads the patch conflict-free (consecutive threads ->
// consecutive channels) and finishes with one atomicAdd per accumulator.
// ---------------------------------------------------------------------------
constexpr int WG_MAXI = 32;

// Persistent: gridDim.x CTAs stride over all (sample, tile) pairs and keep their partial sums in
// registers, so the final reduction is gridDim.x atomics per weight instead of one per tile.
// Per-sample factors (rs/cs) are not supported here (the skinny layers are never modulated).
template <typename T, typename TDY>
__global__ void __launch_bounds__(256) wgrad_small_cout_kernel(WgradP p, int TT, int tiles_w,
                                                               int tiles_per_img, int total_tiles,
                                                               int vec8) {
  extern __shared__ __align__(16) unsigned char wg_smem[];
  const int PW = TT + p.kw - 1, PH = TT + p.kh - 1;
  T* xs = reinterpret_cast<T*>(wg_smem);                                   // [PH][PW][cin]
  float* dys = reinterpret_cast<float*>(xs + (size_t)PH * PW * p.cin + 8);  // [TT*TT][cout]
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  const int patch = PH * PW * p.cin;
  const int total = p.ktot * p.cout;  // items: (o, tap, c), c fastest
  float acc[WG_MAXI];
  int base[WG_MAXI], oidx[WG_MAXI];
#pragma unroll
  for (int i = 0; i < WG_MAXI; ++i) {
    acc[i] = 0.f;
    int item = threadIdx.x + i * 256;
    if (item < total) {
      int kk = item % p.ktot;
      int tap = kk / p.cin, c = kk - tap * p.cin;
      int r = tap / p.kw, s2 = tap - r * p.kw;
      base[i] = (r * PW + s2) * p.cin + c;
      oidx[i] = item / p.ktot;
    } else {
      base[i] = -1;
      oidx[i] = 0;
    }
  }
  const int n_items = (total + 255) / 256;  // block-uniform trip count
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * TT, ow0 = (tile % tiles_w) * TT;
    __syncthreads();  // previous tile fully consumed
    if (vec8) {  // 128-bit staging: 8 channels per load
      const int cv = p.cin / 8;
      for (int e = threadIdx.x; e < PH * PW * cv; e += 256) {
        int v = e % cv, q = e / cv;
        int pw = q % PW, ph = q / PW;
        int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
        float tmp[8];
        if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo) {
          load_vec<T, 8>(vptr<T>(p.x, n, ih, iw, v * 8), tmp);
        } else {
#pragma unroll
          for (int i = 0; i < 8; ++i) tmp[i] = 0.f;
        }
        store_vec<T, 8>(xs + (size_t)q * p.cin + v * 8, tmp);
      }
    } else {
      for (int e = threadIdx.x; e < patch; e += 256) {
        int c = e % p.cin, q = e / p.cin;
        int pw = q % PW, ph = q / PW;
        int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
        T v = from_f<T>(0.f);
        if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo) v = *vptr<T>(p.x, n, ih, iw, c);
        xs[e] = v;
      }
    }
    for (int e = threadIdx.x; e < TT * TT * p.cout; e += 256) {
      int o = e % p.cout, q = e / p.cout;
      int oh = oh0 + q / TT, ow = ow0 + q % TT;
      float v = 0.f;
      if (oh < p.dy.h && ow < p.dy.w) v = to_f(*vptr<TDY>(p.dy, n, oh, ow, o));
      dys[e] = v;
    }
    __syncthreads();
    for (int qh = 0; qh < TT; ++qh) {
      for (int qw = 0; qw < TT; ++qw) {
        const int pixoff = (qh * PW + qw) * p.cin;
        const float* dq = dys + (qh * TT + qw) * p.cout;
#pragma unroll
        for (int i = 0; i < WG_MAXI; ++i) {
          if (i < n_items && base[i] >= 0)
            acc[i] = fmaf(dq[oidx[i]], to_f(xs[base[i] + pixoff]), acc[i]);
        }
      }
    }
  }
  const int taps = p.kh * p.kw;
#pragma unroll
  for (int i = 0; i < WG_MAXI; ++i) {
    if (base[i] < 0) continue;
    int item = threadIdx.x + i * 256;
    int o = item / p.ktot, kk = item % p.ktot;
    int tap = kk / p.cin, c = kk - tap * p.cin;
    atomicAdd(p.dw + ((long long)o * p.cin + c) * taps + tap, acc[i] * p.alpha);
  }
}

// ---------------------------------------------------------------------------
// wgrad for Cout == 1 (generator output conv 64->1 7x7, discriminator output conv 512->1 4x4):
// dw[tap][c] = sum_px dy[px] * x[px + tap][c].  A thread owns one tap and 8 channels: per pixel
// one broadcast dy value and one 128-bit patch load feed 8 FMAs.  Persistent over tiles.
// blockDim = taps * cin/8 (<= 1024).
// ---------------------------------------------------------------------------
template <typename T, typename TDY>
__global__ void __launch_bounds__(1024)
wgrad_cout1_kernel(WgradP p, int TT, int tiles_w, int tiles_per_img, int total_tiles) {
  extern __shared__ __align__(16) unsigned char wo1_smem[];
  const int PW = TT + p.kw - 1, PH = TT + p.kh - 1;
  const int cv = p.cin / 8;
  T* xs = reinterpret_cast<T*>(wo1_smem);                                   // [PH*PW][cin]
  float* dys = reinterpret_cast<float*>(xs + (size_t)PH * PW * p.cin);      // [TT*TT]
  const int nthreads = blockDim.x;
  const int tap = threadIdx.x / cv, cg = threadIdx.x % cv;
  const int r = tap / p.kw, s2 = tap % p.kw;
  const bool active = tap < p.kh * p.kw;
  const int H = p.x.h, W = p.x.w, halo = p.x_halo;
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  for (int t = blockIdx.x; t < total_tiles; t += gridDim.x) {
    const int n = t / tiles_per_img, tile = t - n * tiles_per_img;
    const int oh0 = (tile / tiles_w) * TT, ow0 = (tile % tiles_w) * TT;
    __syncthreads();
    for (int e = threadIdx.x; e < PH * PW * cv; e += nthreads) {
      int v = e % cv, q = e / cv;
      int pw = q % PW, ph = q / PW;
      int ih = oh0 + ph - p.pad, iw = ow0 + pw - p.pad;
      float tmp[8];
      if (ih >= -halo && ih < H + halo && iw >= -halo && iw < W + halo) {
        load_vec<T, 8>(vptr<T>(p.x, n, ih, iw, v * 8), tmp);
      } else {
#pragma unroll
        for (int i = 0; i < 8; ++i) tmp[i] = 0.f;
      }
      store_vec<T, 8>(xs + (size_t)q * p.cin + v * 8, tmp);
    }
    for (int e = threadIdx.x; e < TT * TT; e += nthreads) {
      int oh = oh0 + e / TT, ow = ow0 + e % TT;
      float v = 0.f;
      if (oh < p.dy.h && ow < p.dy.w) v = to_f(*vptr<TDY>(p.dy, n, oh, ow, 0));
      dys[e] = v;
    }
    __syncthreads();
    if (active) {
      for (int qh = 0; qh < TT; ++qh) {
        const T* xrow = xs + ((size_t)(qh + r) * PW + s2) * p.cin + cg * 8;
        const float* drow = dys + qh * TT;
#pragma unroll 4
        for (int qw = 0; qw < TT; ++qw) {
          float xv[8];
          load_vec_rw<T, 8>(xrow + (size_t)qw * p.cin, xv);
          const float d = drow[qw];
#pragma unroll
          for (int j = 0; j < 8; ++j) acc[j] = fmaf(d, xv[j], acc[j]);
        }
      }
    }
  }
  if (active) {
    const int taps = p.kh * p.kw;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      atomicAdd(p.dw + (long long)(cg * 8 + j) * taps + tap, acc[j] * p.alpha);
  }
}

// ---------------------------------------------------------------------------
// weight staging / modulation coefficients
// ---------------------------------------------------------------------------
// One thread stages 8 consecutive output elements (one 128-bit store for bf16) for ALL nb
// per-sample packs: the 8 weights are gathered from the [O,I,kh,kw] parameter once, then scaled
// by the per-sample factors and streamed out.
template <typename TO>
__global__ void __launch_bounds__(256)
weight_pack_kernel(const float* __restrict__ w, int cout, int cin, int kh, int kw, float alpha,
                   const float* __restrict__ cs, const float* __restrict__ rs, int nb, int transpose,
                   TO* out) {
  const int taps = kh * kw;
  const long long per = (long long)cout * cin * taps;
  const long long groups = per / 8;
  for (long long gidx = (long long)blockIdx.x * blockDim.x + threadIdx.x; gidx < groups;
       gidx += (long long)gridDim.x * blockDim.x) {
    const long long e = gidx * 8;
    float wv[8];
    int o0, i0, r, s2;
    if (!transpose) {  // [o][r][s][i], vector along i
      i0 = (int)(e % cin); long long t = e / cin;
      s2 = (int)(t % kw); t /= kw;
      r = (int)(t % kh); o0 = (int)(t / kh);
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = alpha * w[((long long)o0 * cin + i0 + j) * taps + r * kw + s2];
    } else {  // [i][kh-1-r][kw-1-s][o], vector along o
      o0 = (int)(e % cout); long long t = e / cout;
      int sf = (int)(t % kw); t /= kw;
      int rf = (int)(t % kh); i0 = (int)(t / kh);
      r = kh - 1 - rf; s2 = kw - 1 - sf;
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = alpha * w[((long long)(o0 + j) * cin + i0) * taps + r * kw + s2];
    }
    // grid.y slices the samples: a 128x128x3x3 pack is only 72 CTAs' worth of 8-element groups,
    // and walking all nb (up to 96) samples serially per thread kept the 19-28 MB per-sample
    // packs at ~0.7 TB/s with half the SMs idle
    const int b_lo = (int)((long long)nb * blockIdx.y / gridDim.y);
    const int b_hi = (int)((long long)nb * (blockIdx.y + 1) / gridDim.y);
    for (int b = b_lo; b < b_hi; ++b) {
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float f = 1.f;
        if (cs) f *= cs[(long long)b * cin + (transpose ? i0 : i0 + j)];
        if (rs) f *= rs[(long long)b * cout + (transpose ? o0 + j : o0)];
        v[j] = wv[j] * f;
      }
      store_vec<TO, 8>(out + (long long)b * per + e, v);
    }
  }
}

// Many shared packs in ONE launch (otm_weight_pack_multi): after an optimiser step every staged
// pack of that network is stale, and rebuilding ~50 packs of <= 0.6 MB each as separate launches
// cost 5 us apiece (0.25 ms per iteration).  The job table travels as a kernel parameter, so the
// launch is graph-capturable without a device-side table; a block serves one job.
constexpr int PACK_MAX_JOBS = 64;
struct PackJob {
  const float* w;
  void* out;
  int cout, cin, kh, kw;
  float alpha;
  int transpose;
  int block0;  // first block of this job
};
struct PackJobs {
  int n, total_blocks;
  PackJob j[PACK_MAX_JOBS];
};

template <typename TO>
__global__ void __launch_bounds__(256) weight_pack_multi_kernel(const __grid_constant__ PackJobs jobs) {
  int jb = 0;
  while (jb + 1 < jobs.n && (int)blockIdx.x >= jobs.j[jb + 1].block0) ++jb;
  const PackJob& J = jobs.j[jb];
  const int nblk = (jb + 1 < jobs.n ? jobs.j[jb + 1].block0 : jobs.total_blocks) - J.block0;
  const int cout = J.cout, cin = J.cin, kh = J.kh, kw = J.kw, taps = kh * kw;
  const float alpha = J.alpha;
  const float* __restrict__ w = J.w;
  TO* out = reinterpret_cast<TO*>(J.out);
  const long long groups = (long long)cout * cin * taps / 8;
  for (long long gidx = (long long)(blockIdx.x - J.block0) * blockDim.x + threadIdx.x; gidx < groups;
       gidx += (long long)nblk * blockDim.x) {
    const long long e = gidx * 8;
    float wv[8];
    if (!J.transpose) {  // [o][r][s][i], vector along i
      const int i0 = (int)(e % cin); long long t = e / cin;
      const int s2 = (int)(t % kw); t /= kw;
      const int r = (int)(t % kh), o0 = (int)(t / kh);
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = alpha * w[((long long)o0 * cin + i0 + j) * taps + r * kw + s2];
    } else {  // [i][kh-1-r][kw-1-s][o], vector along o
      const int o0 = (int)(e % cout); long long t = e / cout;
      const int sf = (int)(t % kw); t /= kw;
      const int rf = (int)(t % kh), i0 = (int)(t / kh);
      const int r = kh - 1 - rf, s2 = kw - 1 - sf;
#pragma unroll
      for (int j = 0; j < 8; ++j) wv[j] = alpha * w[((long long)(o0 + j) * cin + i0) * taps + r * kw + s2];
    }
    store_vec<TO, 8>(out + e, wv);
  }
}

// scalar fallback for shapes whose innermost packed dimension is not a multiple of 8
template <typename TO>
__global__ void weight_pack_scalar_kernel(const float* __restrict__ w, int cout, int cin, int kh,
                                          int kw, float alpha, const float* cs, const float* rs,
                                          int nb, int transpose, TO* out) {
  const int taps = kh * kw;
  const long long per = (long long)cout * cin * taps;
  const long long total = per * nb;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    int b = (int)(idx / per);
    long long e = idx - (long long)b * per;
    int o, i, r, s;
    if (!transpose) {
      i = (int)(e % cin); long long t = e / cin;
      s = (int)(t % kw); t /= kw;
      r = (int)(t % kh); o = (int)(t / kh);
    } else {
      o = (int)(e % cout); long long t = e / cout;
      int s2 = (int)(t % kw); t /= kw;
      int r2 = (int)(t % kh); i = (int)(t / kh);
      r = kh - 1 - r2; s = kw - 1 - s2;
    }
    float v = alpha * w[((long long)o * cin + i) * taps + r * kw + s];
    if (cs) v *= cs[(long long)b * cin + i];
    if (rs) v *= rs[(long long)b * cout + o];
    out[idx] = from_f<TO>(v);
  }
}

__global__ void weight_sqsum_kernel(const float* __restrict__ w, long long count, int taps,
                                    float alpha, float* q) {
  long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= count) return;
  float s = 0.f;
  for (int k = 0; k < taps; ++k) {
    float v = alpha * w[i * taps + k];
    s = fmaf(v, v, s);
  }
  q[i] = s;
}

// one warp per (b,o)
__global__ void demod_kernel(const float* __restrict__ s, const float* __restrict__ q, int nb,
                             int cout, int cin, float eps, float* sigma_inv) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) / 32, lane = threadIdx.x % 32;
  if (warp >= nb * cout) return;
  int b = warp / cout, o = warp - b * cout;
  float acc = 0.f;
  for (int i = lane; i < cin; i += 32) {
    float sv = s[(long long)b * cin + i];
    acc = fmaf(sv * sv, q[(long long)o * cin + i], acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) sigma_inv[warp] = rsqrtf(acc + eps);
}

// ds[b,i] = Q[b,i] + 2 s[b,i] sum_o dd[b,o] q[o,i],  dd = -0.5 sigma_inv^2 P
// A CTA (4 warps) owns 32 consecutive i of one sample: the warps split the o loop (a 128-term
// dependent chain per thread took 20 us per launch), partial sums meet in shared memory.
__global__ void __launch_bounds__(128) mod_bwd_ds_kernel(otm_mod_bwd_args a) {
  __shared__ float part[4][32];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const int blocks_per_b = (a.cin + 31) / 32;
  const int b = blockIdx.x / blocks_per_b, i = (blockIdx.x % blocks_per_b) * 32 + lane;
  float acc = 0.f;
  if (i < a.cin) {
    const int o0 = a.cout * wq / 4, o1 = a.cout * (wq + 1) / 4;
#pragma unroll 4
    for (int o = o0; o < o1; ++o) {
      const float si = a.sigma_inv[b * a.cout + o];
      const float dd = -0.5f * si * si * a.P[b * a.cout + o];
      acc = fmaf(dd, a.q[(long long)o * a.cin + i], acc);
    }
  }
  part[wq][lane] = acc;
  __syncthreads();
  if (wq == 0 && i < a.cin) {
    const int idx = b * a.cin + i;
    float qv = a.Q[idx];
    if (a.q_scaled) qv = a.s[idx] != 0.f ? qv / a.s[idx] : 0.f;  // Q was reduced against s * x
    a.ds[idx] = qv + 2.f * a.s[idx] * (part[0][lane] + part[1][lane] + part[2][lane] + part[3][lane]);
  }
}
// dw[o,i,k] += 2 alpha^2 w[o,i,k] sum_b dd[b,o] s[b,i]^2
// A CTA (4 warps) owns 32 consecutive (o,i) pairs: the warps split the sample loop, the partial
// sums meet in shared memory and all 128 threads then update the 32*taps contiguous weights.
__global__ void __launch_bounds__(128) mod_bwd_dw_kernel(otm_mod_bwd_args a) {
  __shared__ float part[4][32];
  __shared__ float fs[32];
  const int lane = threadIdx.x & 31, wq = threadIdx.x >> 5;
  const long long idx0 = (long long)blockIdx.x * 32;
  const long long total = (long long)a.cout * a.cin;
  const long long idx = idx0 + lane;
  float dq = 0.f;
  if (idx < total) {
    const int o = (int)(idx / a.cin), i = (int)(idx - (long long)o * a.cin);
    const int b0 = a.nb * wq / 4, b1 = a.nb * (wq + 1) / 4;
#pragma unroll 4
    for (int b = b0; b < b1; ++b) {
      const float si = a.sigma_inv[b * a.cout + o];
      const float dd = -0.5f * si * si * a.P[b * a.cout + o];
      const float sv = a.s[b * a.cin + i];
      dq = fmaf(dd, sv * sv, dq);
    }
  }
  part[wq][lane] = dq;
  __syncthreads();
  if (wq == 0)
    fs[lane] = 2.f * a.alpha * a.alpha * (part[0][lane] + part[1][lane] + part[2][lane] + part[3][lane]);
  __syncthreads();
  const int cnt = 32 * a.taps;
  for (int e = threadIdx.x; e < cnt; e += 128) {
    const long long g = idx0 * a.taps + e;
    if (g < total * a.taps) a.dw[g] += fs[e / a.taps] * a.w[g];
  }
}

int conv_fwd_simt(const otm_conv_fwd_args* a, cudaStream_t st) {
  ConvP p;
  p.x = make_view(a->x); p.y = make_view(a->y);
  p.res = a->residual.ptr ? make_view(a->residual) : null_view();
  p.w = a->wpack; p.w_bstride = a->w_batch_stride;
  p.x_halo = a->x_halo; p.y_halo = a->y_halo; p.kh = a->kh; p.kw = a->kw; p.pad = a->pad;
  p.cin = a->x.c; p.cout = a->y.c; p.ktot = a->kh * a->kw * a->x.c;
  p.alpha = a->alpha; p.row_scale = a->row_scale; p.bias = a->bias; p.act = a->act;
  p.post_scale = a->post_scale;
  p.tiles_w = (a->y.w + 7) / 8;
  const int tiles_h = (a->y.h + 7) / 8;
  const bool in_bf = a->x.dtype == OTM_BF16, out_bf = a->y.dtype == OTM_BF16;
  if (a->post_scale) {  // only the generic tile kernel applies the post-activation scale
    dim3 grid_ps(p.tiles_w * tiles_h, (p.cout + TN - 1) / TN, a->y.n);
    if (in_bf && out_bf) conv_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid_ps, 256, 0, st>>>(p);
    else if (in_bf) conv_simt_kernel<__nv_bfloat16, float><<<grid_ps, 256, 0, st>>>(p);
    else if (out_bf) conv_simt_kernel<float, __nv_bfloat16><<<grid_ps, 256, 0, st>>>(p);
    else conv_simt_kernel<float, float><<<grid_ps, 256, 0, st>>>(p);
    OTM_LAUNCH_CHECK();
    return OTM_OK;
  }
  if (p.cin == 1 && p.cout == 64 && a->kh == a->kw && (a->kh == 4 || a->kh == 7) && !in_bf &&
      a->w_batch_stride == 0 && vec_ok(a->y, 8) && vec_ok(a->residual, 8)) {
    constexpr int TT = 16;
    const int tiles_w = (a->y.w + TT - 1) / TT, tiles_h = (a->y.h + TT - 1) / TT;
    const int per_img = tiles_w * tiles_h, total = per_img * a->y.n;
    int ctas = num_sms() * 2;
    if (ctas > total) ctas = total;
    const bool mma = out_bf && !a->row_scale && !a->residual.ptr;
    if (a->kh == 7) {
      if (mma) conv_cin1_mma_kernel<7><<<ctas, 256, 0, st>>>(p, tiles_w, per_img, total);
      else if (out_bf) conv_cin1_kernel<__nv_bfloat16, 7><<<ctas, 256, 0, st>>>(p, tiles_w, per_img, total);
      else conv_cin1_kernel<float, 7><<<ctas, 256, 0, st>>>(p, tiles_w, per_img, total);
    } else {
      if (mma) conv_cin1_mma_kernel<4><<<ctas, 256, 0, st>>>(p, tiles_w, per_img, total);
      else if (out_bf) conv_cin1_kernel<__nv_bfloat16, 4><<<ctas, 256, 0, st>>>(p, tiles_w, per_img, total);
      else conv_cin1_kernel<float, 4><<<ctas, 256, 0, st>>>(p, tiles_w, per_img, total);
    }
    OTM_LAUNCH_CHECK();
    return OTM_OK;
  }
  if (p.cout == 1 && p.cin == 64 && in_bf && a->kh == a->kw && (a->kh == 4 || a->kh == 7) &&
      a->w_batch_stride == 0 && vec_ok(a->x, 8) && a->y.h * a->y.w >= 256) {
    constexpr int mw = 1;  // one M-warp per CTA measured best (3 CTAs / SM)
#define OTM_C1M(TO, KS, MW)                                                                      \
  do {                                                                                            \
    using G = Cout1Geom<KS, MW>;                                                                  \
    const size_t smem = (size_t)G::PATCH_ELEMS * 2 + (size_t)8 * 2 * G::PWC * 8 * sizeof(float);  \
    const int tiles_w = (a->y.w + G::TWO - 1) / G::TWO, tiles_h = (a->y.h + G::TH - 1) / G::TH;   \
    const int per_img = tiles_w * tiles_h, total = per_img * a->y.n;                              \
    auto kern = conv_cout1_mma_kernel<TO, KS, MW>;                                                \
    OTM_ENSURE_SMEM(kern,      \
                                          (int)smem);                                                                                             \
    int ctas = num_sms() * (MW == 1 ? 3 : 1);                                                     \
    if (ctas > total) ctas = total;                                                               \
    kern<<<ctas, 256, smem, st>>>(p, tiles_w, per_img, total);                                    \
  } while (0)
#define OTM_C1M_KS(TO, MW) do { if (a->kh == 7) OTM_C1M(TO, 7, MW); else OTM_C1M(TO, 4, MW); } while (0)
    if (out_bf) { if (mw == 1) OTM_C1M_KS(__nv_bfloat16, 1); else OTM_C1M_KS(__nv_bfloat16, 2); }
    else { if (mw == 1) OTM_C1M_KS(float, 1); else OTM_C1M_KS(float, 2); }
#undef OTM_C1M_KS
#undef OTM_C1M
    OTM_LAUNCH_CHECK();
    return OTM_OK;
  }
  if (p.cout == 1 && p.cin % 64 == 0 && a->kh == a->kw && (a->kh == 4 || a->kh == 7) &&
      vec_ok(a->x, 8) && a->y.h * a->y.w >= 256) {
#define OTM_C1(TI, TO) \
  (a->kh == 7 ? launch_cout1_tiled<TI, TO, 7>(p, a->y.n, st) : launch_cout1_tiled<TI, TO, 4>(p, a->y.n, st))
    if (in_bf && out_bf) return OTM_C1(__nv_bfloat16, __nv_bfloat16);
    if (in_bf) return OTM_C1(__nv_bfloat16, float);
    if (out_bf) return OTM_C1(float, __nv_bfloat16);
    return OTM_C1(float, float);
#undef OTM_C1
  }
  if (p.cout == 1 && p.cin % 256 == 0 && vec_ok(a->x, 8) && ((uintptr_t)a->wpack % 16 == 0) &&
      (a->w_batch_stride % 8 == 0)) {
    const int total_px = a->y.n * a->y.h * a->y.w;
    const int blocks = (total_px * 32 + 255) / 256;
    if (in_bf && out_bf) conv_cout1_warp_kernel<__nv_bfloat16, __nv_bfloat16><<<blocks, 256, 0, st>>>(p, total_px);
    else if (in_bf) conv_cout1_warp_kernel<__nv_bfloat16, float><<<blocks, 256, 0, st>>>(p, total_px);
    else if (out_bf) conv_cout1_warp_kernel<float, __nv_bfloat16><<<blocks, 256, 0, st>>>(p, total_px);
    else conv_cout1_warp_kernel<float, float><<<blocks, 256, 0, st>>>(p, total_px);
    OTM_LAUNCH_CHECK();
    return OTM_OK;
  }
  if (p.cout <= 4 && (size_t)p.cout * p.ktot * 4 <= 160 * 1024) {
    size_t smem = (size_t)p.cout * p.ktot * sizeof(float);
    bool v8 = vec_ok(a->x, 8);
    dim3 grid((a->y.h * a->y.w + 127) / 128, 1, a->y.n);
#define OTM_LAUNCH_SMALL(TI, TO, V)                                                          \
  do {                                                                                        \
    auto kern = conv_small_cout_kernel<TI, TO, V>;                                            \
    OTM_ENSURE_SMEM(kern,  \
                                          160 * 1024);                                                                                         \
    kern<<<grid, 128, smem, st>>>(p);                                                         \
  } while (0)
    if (in_bf && out_bf) { if (v8) OTM_LAUNCH_SMALL(__nv_bfloat16, __nv_bfloat16, 8); else OTM_LAUNCH_SMALL(__nv_bfloat16, __nv_bfloat16, 1); }
    else if (in_bf) { if (v8) OTM_LAUNCH_SMALL(__nv_bfloat16, float, 8); else OTM_LAUNCH_SMALL(__nv_bfloat16, float, 1); }
    else if (out_bf) { if (v8) OTM_LAUNCH_SMALL(float, __nv_bfloat16, 8); else OTM_LAUNCH_SMALL(float, __nv_bfloat16, 1); }
    else { if (v8) OTM_LAUNCH_SMALL(float, float, 8); else OTM_LAUNCH_SMALL(float, float, 1); }
#undef OTM_LAUNCH_SMALL
    OTM_LAUNCH_CHECK();
    return OTM_OK;
  }
  dim3 grid(p.tiles_w * tiles_h, (p.cout + TN - 1) / TN, a->y.n);
  if (in_bf && out_bf) conv_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else if (in_bf) conv_simt_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p);
  else if (out_bf) conv_simt_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else conv_simt_kernel<float, float><<<grid, 256, 0, st>>>(p);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int conv_wgrad_simt(const otm_conv_wgrad_args* a, cudaStream_t st) {
  WgradP p;
  p.x = make_view(a->x); p.dy = make_view(a->dy);
  p.x_halo = a->x_halo; p.kh = a->kh; p.kw = a->kw; p.pad = a->pad;
  p.cin = a->x.c; p.cout = a->dy.c; p.ktot = a->kh * a->kw * a->x.c;
  p.dw = a->dw; p.alpha = a->alpha; p.rs = a->rs; p.cs = a->cs;
  // single input channel (image-side convs)
  if (p.cin == 1 && p.cout == 64 && a->kh == a->kw && (a->kh == 4 || a->kh == 7) &&
      a->x.dtype == OTM_F32 && !a->rs && !a->cs && vec_ok(a->dy, 8)) {
    constexpr int TT = 16;
    const int tiles_w = (a->dy.w + TT - 1) / TT, tiles_h = (a->dy.h + TT - 1) / TT;
    const int per_img = tiles_w * tiles_h, total = per_img * a->dy.n;
    int ctas = num_sms() * 2;
    if (ctas > total) ctas = total;
    const bool yb = a->dy.dtype == OTM_BF16;
#define OTM_WC1(TDY, KS, NT)                                                                      \
  do {                                                                                            \
    const size_t smem = sizeof(TDY) * TT * TT * 64 + sizeof(float) * (TT + KS - 1) * (TT + KS - 1); \
    auto kern = wgrad_cin1_kernel<TDY, KS>;                                                       \
    OTM_ENSURE_SMEM(kern,      \
                                          (int)smem);                                                                                             \
    kern<<<ctas, NT, smem, st>>>(p, tiles_w, per_img, total);                                     \
  } while (0)
    if (yb) {
#define OTM_WM1(KS)                                                                              \
  do {                                                                                            \
    const size_t smem = sizeof(__nv_bfloat16) * (TT * TT * 72 + (TT + KS - 1) * (TT + KS - 1) + 8); \
    auto kern = wgrad_cin1_mma_kernel<KS>;                                                        \
    OTM_ENSURE_SMEM(kern,      \
                                          (int)smem);                                                                                             \
    kern<<<ctas, 256, smem, st>>>(p, tiles_w, per_img, total);                                    \
  } while (0)
      if (a->kh == 7) OTM_WM1(7); else OTM_WM1(4);
#undef OTM_WM1
      OTM_LAUNCH_CHECK();
      return OTM_OK;
    }
    if (a->kh == 7) {
      if (yb) OTM_WC1(__nv_bfloat16, 7, 416); else OTM_WC1(float, 7, 416);
    } else {
      if (yb) OTM_WC1(__nv_bfloat16, 4, 128); else OTM_WC1(float, 4, 128);
    }
#undef OTM_WC1
    OTM_LAUNCH_CHECK();
    return OTM_OK;
  }
  // single output channel, 64 input channels, bf16 activations: mma.sync kernel
  if (p.cout == 1 && p.cin == 64 && a->x.dtype == OTM_BF16 && a->kh == a->kw &&
      (a->kh == 4 || a->kh == 7) && !a->rs && !a->cs && vec_ok(a->x, 8) && a->dy.h * a->dy.w >= 256) {
    constexpr int mw = 1;  // one M-warp per CTA measured best (3 CTAs / SM)
#define OTM_W1M(TDY, KS, MW)                                                                     \
  do {                                                                                            \
    using G = Cout1Geom<KS, MW>;                                                                  \
    const size_t smem = (size_t)G::PATCH_ELEMS * 2 + (size_t)16 * (G::PWC + 16) * 2;              \
    const int tiles_w = (a->dy.w + G::TWO - 1) / G::TWO, tiles_h = (a->dy.h + G::TH - 1) / G::TH; \
    const int per_img = tiles_w * tiles_h, total = per_img * a->dy.n;                             \
    auto kern = wgrad_cout1_mma_kernel<TDY, KS, MW>;                                              \
    OTM_ENSURE_SMEM(kern,      \
                                          (int)smem);                                                                                             \
    int ctas = num_sms() * (MW == 1 ? 3 : 1);                                                     \
    if (ctas > total) ctas = total;                                                               \
    kern<<<ctas, 256, smem, st>>>(p, tiles_w, per_img, total);                                    \
  } while (0)
#define OTM_W1M_KS(TDY, MW) do { if (a->kh == 7) OTM_W1M(TDY, 7, MW); else OTM_W1M(TDY, 4, MW); } while (0)
    if (a->dy.dtype == OTM_BF16) { if (mw == 1) OTM_W1M_KS(__nv_bfloat16, 1); else OTM_W1M_KS(__nv_bfloat16, 2); }
    else { if (mw == 1) OTM_W1M_KS(float, 1); else OTM_W1M_KS(float, 2); }
#undef OTM_W1M_KS
#undef OTM_W1M
    OTM_LAUNCH_CHECK();
    return OTM_OK;
  }
  // single output channel: (tap, 8-channel) threads over a shared-memory patch
  if (p.cout == 1 && p.cin % 8 == 0 && a->kh * a->kw * (p.cin / 8) <= 1024 && !a->rs && !a->cs &&
      vec_ok(a->x, 8)) {
    const size_t es = dtype_size(a->x.dtype);
    int TT = 16;
    auto smem_for = [&](int tt) {
      return (size_t)(tt + p.kh - 1) * (tt + p.kw - 1) * p.cin * es + (size_t)tt * tt * sizeof(float);
    };
    while (TT > 2 && smem_for(TT) > 100 * 1024) TT /= 2;
    if (smem_for(TT) <= 100 * 1024) {
      const int tiles_w = (a->dy.w + TT - 1) / TT, tiles_h = (a->dy.h + TT - 1) / TT;
      const int per_img = tiles_w * tiles_h, total_tiles = per_img * a->dy.n;
      const size_t smem = smem_for(TT);
      int nthreads = (a->kh * a->kw * (p.cin / 8) + 31) / 32 * 32;
      int ctas = num_sms() * 2;
      if (ctas > total_tiles) ctas = total_tiles;
#define OTM_WO1(TX, TDY)                                                                        \
  do {                                                                                          \
    auto kern = wgrad_cout1_kernel<TX, TDY>;                                                    \
    OTM_ENSURE_SMEM(kern,    \
                                          100 * 1024);                                                                                           \
    kern<<<ctas, nthreads, smem, st>>>(p, TT, tiles_w, per_img, total_tiles);                   \
  } while (0)
      const bool xb = a->x.dtype == OTM_BF16, yb = a->dy.dtype == OTM_BF16;
      if (xb && yb) OTM_WO1(__nv_bfloat16, __nv_bfloat16);
      else if (xb) OTM_WO1(__nv_bfloat16, float);
      else if (yb) OTM_WO1(float, __nv_bfloat16);
      else OTM_WO1(float, float);
#undef OTM_WO1
      OTM_LAUNCH_CHECK();
      return OTM_OK;
    }
  }
  // skinny output: persistent smem-tiled kernel
  if (p.cout <= 4 && (long long)p.ktot * p.cout <= 256LL * WG_MAXI && !a->rs && !a->cs) {
    const size_t es = dtype_size(a->x.dtype);
    int TT = 16;
    auto smem_for = [&](int tt) {
      return (size_t)(tt + p.kh - 1) * (tt + p.kw - 1) * p.cin * es + 8 * es +
             (size_t)tt * tt * p.cout * sizeof(float) + 16;
    };
    while (TT > 2 && smem_for(TT) > 100 * 1024) TT /= 2;
    if (smem_for(TT) <= 200 * 1024) {
      const int tiles_w = (a->dy.w + TT - 1) / TT, tiles_h = (a->dy.h + TT - 1) / TT;
      const int per_img = tiles_w * tiles_h, total_tiles = per_img * a->dy.n;
      const size_t smem = smem_for(TT);
      int ctas = num_sms() * (smem > 110 * 1024 ? 1 : 2);
      if (ctas > total_tiles) ctas = total_tiles;
#define OTM_WG_SMALL(TX, TDY)                                                                   \
  do {                                                                                          \
    auto kern = wgrad_small_cout_kernel<TX, TDY>;                                               \
    OTM_ENSURE_SMEM(kern,    \
                                          200 * 1024);                                                                                           \
    kern<<<ctas, 256, smem, st>>>(p, TT, tiles_w, per_img, total_tiles, vec8);                  \
  } while (0)
      const int vec8 = vec_ok(a->x, 8) ? 1 : 0;
      const bool xb = a->x.dtype == OTM_BF16, yb = a->dy.dtype == OTM_BF16;
      if (xb && yb) OTM_WG_SMALL(__nv_bfloat16, __nv_bfloat16);
      else if (xb) OTM_WG_SMALL(__nv_bfloat16, float);
      else if (yb) OTM_WG_SMALL(float, __nv_bfloat16);
      else OTM_WG_SMALL(float, float);
#undef OTM_WG_SMALL
      OTM_LAUNCH_CHECK();
      return OTM_OK;
    }
  }
  const int HW = a->dy.h * a->dy.w;
  const int base_ctas = ((p.ktot + TN - 1) / TN) * ((p.cout + TM - 1) / TM) * a->dy.n;
  int splits = (num_sms() * 4 + base_ctas - 1) / base_ctas;
  int max_splits = (HW + 255) / 256;
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = splits;
  p.pix_per_split = ((HW + splits - 1) / splits + TK - 1) / TK * TK;
  dim3 grid((p.ktot + TN - 1) / TN, (p.cout + TM - 1) / TM, a->dy.n * splits);
  OTM_REQUIRE(a->dy.n * splits <= 65535, "wgrad: grid.z too large");
  const bool xb = a->x.dtype == OTM_BF16, yb = a->dy.dtype == OTM_BF16;
  if (xb && yb) wgrad_simt_kernel<__nv_bfloat16, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else if (xb) wgrad_simt_kernel<__nv_bfloat16, float><<<grid, 256, 0, st>>>(p);
  else if (yb) wgrad_simt_kernel<float, __nv_bfloat16><<<grid, 256, 0, st>>>(p);
  else wgrad_simt_kernel<float, float><<<grid, 256, 0, st>>>(p);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

}  // namespace otm

using namespace otm;

extern "C" {

int otm_weight_pack(const otm_weight_pack_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->w && a->out && a->nb >= 1, "weight_pack: bad arguments");
  const long long per = (long long)a->cout * a->cin * a->kh * a->kw;
  const int inner = a->transpose ? a->cout : a->cin;
  const int cap = num_sms() * 16;
  if (inner % 8 == 0 && ((uintptr_t)a->out % 16 == 0)) {
    int blocks = (int)((per / 8 + 255) / 256);
    if (blocks > cap) blocks = cap;
    // ~8 CTAs per SM in total, at least 2 samples per thread (amortises the weight gather)
    int ysplit = (num_sms() * 8 + blocks - 1) / blocks;
    if (ysplit > (a->nb + 1) / 2) ysplit = (a->nb + 1) / 2;
    if (ysplit < 1) ysplit = 1;
    const dim3 pgrid(blocks, ysplit);
    if (a->out_dtype == OTM_BF16)
      weight_pack_kernel<__nv_bfloat16><<<pgrid, 256, 0, st>>>(
          a->w, a->cout, a->cin, a->kh, a->kw, a->alpha, a->cs, a->rs, a->nb, a->transpose,
          (__nv_bfloat16*)a->out);
    else
      weight_pack_kernel<float><<<pgrid, 256, 0, st>>>(a->w, a->cout, a->cin, a->kh, a->kw,
                                                        a->alpha, a->cs, a->rs, a->nb,
                                                        a->transpose, (float*)a->out);
  } else {
    long long total = per * a->nb;
    int blocks = (int)((total + 255) / 256);
    if (blocks > cap) blocks = cap;
    if (a->out_dtype == OTM_BF16)
      weight_pack_scalar_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(
          a->w, a->cout, a->cin, a->kh, a->kw, a->alpha, a->cs, a->rs, a->nb, a->transpose,
          (__nv_bfloat16*)a->out);
    else
      weight_pack_scalar_kernel<float><<<blocks, 256, 0, st>>>(
          a->w, a->cout, a->cin, a->kh, a->kw, a->alpha, a->cs, a->rs, a->nb, a->transpose,
          (float*)a->out);
  }
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_weight_pack_multi(const otm_weight_pack_args* jobs, int32_t njobs, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(jobs && njobs >= 1, "weight_pack_multi: bad arguments");
  for (int base = 0; base < njobs; base += PACK_MAX_JOBS) {
    PackJobs pj;
    pj.n = njobs - base < PACK_MAX_JOBS ? njobs - base : PACK_MAX_JOBS;
    int blocks = 0;
    for (int k = 0; k < pj.n; ++k) {
      const otm_weight_pack_args& a = jobs[base + k];
      OTM_REQUIRE(a.w && a.out && a.nb == 1 && !a.cs && !a.rs,
                  "weight_pack_multi: job %d is not a shared pack", base + k);
      OTM_REQUIRE(a.out_dtype == jobs[0].out_dtype, "weight_pack_multi: mixed output dtypes");
      const int inner = a.transpose ? a.cout : a.cin;
      OTM_REQUIRE(inner % 8 == 0 && (uintptr_t)a.out % 16 == 0,
                  "weight_pack_multi: job %d needs an inner dimension that is a multiple of 8 and a "
                  "16-byte aligned output", base + k);
      PackJob& J = pj.j[k];
      J.w = a.w; J.out = a.out; J.cout = a.cout; J.cin = a.cin; J.kh = a.kh; J.kw = a.kw;
      J.alpha = a.alpha; J.transpose = a.transpose; J.block0 = blocks;
      const long long groups = (long long)a.cout * a.cin * a.kh * a.kw / 8;
      long long nb = (groups + 511) / 512;  // two 8-element groups per thread
      if (nb > 64) nb = 64;
      blocks += (int)nb;
    }
    pj.total_blocks = blocks;
    if (jobs[0].out_dtype == OTM_BF16)
      weight_pack_multi_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(pj);
    else
      weight_pack_multi_kernel<float><<<blocks, 256, 0, st>>>(pj);
    OTM_LAUNCH_CHECK();
  }
  return OTM_OK;
}

int otm_weight_sqsum(const float* w, int32_t cout, int32_t cin, int32_t taps, float alpha,
                     float* q, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(w && q, "weight_sqsum: null");
  long long count = (long long)cout * cin;
  weight_sqsum_kernel<<<(int)((count + 255) / 256), 256, 0, st>>>(w, count, taps, alpha, q);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_demod(const float* s, const float* q, int32_t nb, int32_t cout, int32_t cin, float eps,
              float* sigma_inv, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(s && q && sigma_inv, "demod: null");
  int warps = nb * cout;
  demod_kernel<<<(warps * 32 + 255) / 256, 256, 0, st>>>(s, q, nb, cout, cin, eps, sigma_inv);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

int otm_mod_bwd(const otm_mod_bwd_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->w && a->s && a->sigma_inv && a->q && a->P && a->Q && a->ds && a->dw,
              "mod_bwd: null");
  mod_bwd_ds_kernel<<<a->nb * ((a->cin + 31) / 32), 128, 0, st>>>(*a);
  OTM_LAUNCH_CHECK();
  long long cnt = (long long)a->cout * a->cin;
  mod_bwd_dw_kernel<<<(int)((cnt + 31) / 32), 128, 0, st>>>(*a);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

}  // extern "C"
