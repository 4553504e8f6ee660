// Reflect-border correction of a 3x3 dgrad (otm_conv_reflect_border, include/otm_b200.h).
//
// The reference pads every ResnetBlock / ModulatedResnetBlock convolution with
// nn.ReflectionPad2d(1) (reference src/model/blocks.py:21-27,49-56).  Autograd's backward of that
// pair is conv_transpose onto the (H+2) x (W+2) padded tensor followed by the pad's backward, which
// folds the halo ring onto rows / columns 1 and H-2 / W-2.  Rounds 1-2 did the same: a tcgen05
// dgrad launch with pad 2 onto 66 x 66 pixels (45 half tiles of 128 pixels where the 64 x 64
// interior needs 32: +41 % MMA time, 124.6 us against 87 us per n = 96 launch) and a fold in the
// next HBM pass.  Here the dgrad is the plain "same" correlation on H x W and the ring
// (2 (H + W) + 4 positions, 6 % of the pixels) is four thin GEMMs
//     ring value[pos][o] = sum_{t, i} dy_line[pos + t][i] * W_tap(line, t)[o][i]
// (a ring position only sees ONE filter row or column: the other taps read zero padding), run on
// warp-level tensor cores (mma.sync m16n8k16 bf16 -> fp32: M = 66 or 64 is far below a tcgen05
// tile) and added into y with packed-bf16 atomics under the same per-channel epilogue as the main
// launch (row / post scale, ReLU gate, dot reduction): every term is linear in the conv result.
//
// CTA = (border line, 80-position chunk, 64 output channels, sample), 4 warps x 16 channels.
// The dy line (K channels per position, zero-extended by 2) sits in shared memory once, filter tap
// t is a row offset into it (ldmatrix A fragments); the tap's 64 x K weight rows are staged next to
// it (ldmatrix B fragments), both by cp.async into dense XOR-swizzled rows.
#include "common.cuh"

namespace otm {

struct BorderP {
  View x;     // dy [n, K, H, W] bf16
  View y;     // [n, C, H, W] bf16, accumulated into
  View gate;  // optional
  const __nv_bfloat16* wp;
  long long w_batch_stride;
  const float* row_scale;
  const float* post_scale;
  float* dot_sums;
  int K, C, H, W;
  int mchunks;
  int line_rows;  // staged rows of the dy line (positions + taps + shift actually needed)
};

constexpr int BD_MT = 5;               // m16 tiles per CTA
constexpr int BD_M = BD_MT * 16;       // positions per CTA
constexpr int BD_N = 64;               // output channels per CTA

__device__ __forceinline__ void bd_ldmatrix_x4(uint32_t (&r)[4], const void* smem_row) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(smem_row);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void bd_mma(float (&d)[4], const uint32_t (&a)[4], const uint32_t (&b)[2]) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, "
      "{%8,%9}, {%0,%1,%2,%3};"
      : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
}

// reflected target pixel of ring position `pos` on `line`
__device__ __forceinline__ void ring_target(int line, int pos, int H, int W, int& ta, int& tb) {
  if (line < 2) {
    const int bcol = pos - 1;
    ta = line == 0 ? 1 : H - 2;
    tb = bcol < 0 ? 1 : (bcol >= W ? W - 2 : bcol);
  } else {
    ta = pos < H ? pos : H - 1;  // (pos >= L is masked by the caller; keep the address in range)
    tb = line == 2 ? 1 : W - 2;
  }
}

// line 0: top halo row (a = -1), 1: bottom (a = H), 2: left halo column (b = -1), 3: right (b = W)
__global__ void __launch_bounds__(128, 6) conv_reflect_border_kernel(BorderP p) {
  extern __shared__ __align__(16) uint8_t bd_smem[];
  const int line = blockIdx.x / p.mchunks, mc = blockIdx.x % p.mchunks;
  const int n = blockIdx.z;
  const bool horiz = line < 2;
  const int L = horiz ? p.W + 2 : p.H;    // ring positions on this line
  const int ext = horiz ? p.W : p.H;      // extent of the dy line
  const int p0 = mc * BD_M;
  if (p0 >= L) return;
  const int pitch = p.K * 2;              // bytes per staged row; 16-byte chunks XOR-swizzled by row & 7
  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;

  // Shared memory: the dy line, then ONE weight buffer [64 output channels][K] holding the current
  // filter tap.  Everything arrives by 16-byte cp.async, rows are dense (pitch 2K bytes) with the
  // 16-byte chunks XOR-swizzled by (row & 7) for conflict-free ldmatrix: 37 KB per CTA at K = 128,
  // six CTAs per SM, so the 768 CTAs of an n = 96 launch are one wave.  (The first version read the
  // weight fragments with 32-bit global loads -- 8 rows x 16 B per instruction -- and ncu showed
  // half of all stall samples on the first MMA / ldmatrix after each batch of them; its 5 CTAs
  // per SM also left a 28-CTA second wave.)
  uint8_t* const bsm = bd_smem + (size_t)p.line_rows * pitch;
  const __nv_bfloat16* wsample = p.wp + (long long)n * p.w_batch_stride;
  const long long wrow = 9LL * p.K;  // elements per output channel in the pack
  const int chunks = p.K / 8;        // 16-byte chunks per row
  auto tap_of = [&](int t) {
    // pack tap (r', s') a ring position sees: top r' = 2, bottom r' = 0, left s' = 2, right s' = 0
    return line == 0 ? 6 + t : line == 1 ? t : line == 2 ? 3 * t + 2 : 3 * t;
  };
  auto stage_weights = [&](int t) {
    const __nv_bfloat16* src = wsample + (long long)(blockIdx.y * BD_N) * wrow + (long long)tap_of(t) * p.K;
    for (int idx = threadIdx.x; idx < BD_N * chunks; idx += blockDim.x) {
      const int r = idx / chunks, c8 = idx - r * chunks;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(bsm + (size_t)r * pitch + ((c8 ^ (r & 7)) << 4))),
                   "l"(src + (long long)r * wrow + c8 * 8)
                   : "memory");
    }
    asm volatile("cp.async.commit_group;" ::: "memory");
  };

  // ---- stage the dy line: row j <-> line coordinate q = p0 + j - 2, zero outside [0, ext) ----
  for (int idx = threadIdx.x; idx < p.line_rows * chunks; idx += blockDim.x) {
    const int j = idx / chunks, c8 = idx - j * chunks;
    const int q = p0 + j - 2;
    uint8_t* dst = bd_smem + (size_t)j * pitch + ((c8 ^ (j & 7)) << 4);
    if (q >= 0 && q < ext) {
      const int hh = line == 0 ? 0 : line == 1 ? p.H - 1 : q;
      const int ww = line == 2 ? 0 : line == 3 ? p.W - 1 : q;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(
                       (uint32_t)__cvta_generic_to_shared(dst)),
                   "l"(vptr<__nv_bfloat16>(p.x, n, hh, ww, c8 * 8))
                   : "memory");
    } else {
      *reinterpret_cast<uint4*>(dst) = make_uint4(0u, 0u, 0u, 0u);
    }
  }
  stage_weights(0);  // one group: line + tap 0

  // the gate pixels of the epilogue: towards L2 now, under the copies and the MMAs
  if (p.gate.ptr) {
    const int pos = p0 + (int)threadIdx.x;
    if (threadIdx.x < BD_M && pos < L) {
      int ta, tb;
      ring_target(line, pos, p.H, p.W, ta, tb);
      asm volatile("prefetch.global.L2 [%0];" ::"l"(
          vptr<__nv_bfloat16>(p.gate, n, ta, tb, blockIdx.y * BD_N)));
    }
  }

  // ring position pos (line-local, pos = p0 + m) reads line coordinates pos + t - 2 + shift,
  // t = 0..2: horizontal lines have pos = b + 1 (b = -1 .. W), vertical ones pos = a (a = 0 .. H-1)
  const int shift = horiz ? 0 : 1;
  const int o0 = blockIdx.y * BD_N + warp * 16;
  float acc[BD_MT][2][4];
#pragma unroll
  for (int a = 0; a < BD_MT; ++a)
#pragma unroll
    for (int b = 0; b < 2; ++b)
#pragma unroll
      for (int e = 0; e < 4; ++e) acc[a][b][e] = 0.f;

  const int mt_n = min(BD_MT, (L - p0 + 15) / 16);  // m16 tiles that hold ring positions
#pragma unroll 1
  for (int t = 0; t < 3; ++t) {
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    __syncthreads();  // tap t's weights (and, for t = 0, the line) have landed
    // A fragment rows: t + shift + mt * 16 + (lane & 7) + 8 * ((lane >> 3) & 1); row & 7 below
    const int a_sw = (t + shift + (lane & 7)) & 7;
    const uint8_t* arow = bd_smem + (size_t)(t + shift + (lane & 7) + ((lane >> 3) & 1) * 8) * pitch;
    // B fragments of this warp's two n8 tiles: matrices (rows +0..7, k 0..7), (+0..7, k 8..15),
    // (+8..15, k 0..7), (+8..15, k 8..15) = b0[0], b0[1], b1[0], b1[1]; row & 7 == lane & 7
    const uint8_t* brow = bsm + (size_t)(warp * 16 + (lane & 7) + ((lane >> 4) & 1) * 8) * pitch;
#pragma unroll 2
    for (int kc = 0; kc < p.K / 16; ++kc) {
      uint32_t bq[4];
      bd_ldmatrix_x4(bq, brow + (((kc * 2 + ((lane >> 3) & 1)) ^ (lane & 7)) << 4));
      const uint32_t b0[2] = {bq[0], bq[1]}, b1[2] = {bq[2], bq[3]};
      const int a_off = ((kc * 2 + (lane >> 4)) ^ a_sw) << 4;
#pragma unroll
      for (int mt = 0; mt < BD_MT; ++mt) {
        if (mt < mt_n) {  // (warp-uniform) a 64-position vertical line has no fifth tile
          uint32_t a[4];
          bd_ldmatrix_x4(a, arow + (size_t)(mt * 16) * pitch + a_off);
          bd_mma(acc[mt][0], a, b0);
          bd_mma(acc[mt][1], a, b1);
        }
      }
    }
    if (t < 2) {
      __syncthreads();  // every warp is done with tap t's weights
      stage_weights(t + 1);
    }
  }

  // ---- epilogue: scale / gate / dot, packed-bf16 atomic add onto the reflected target ----
  float f[2][2], dot[2][2];
#pragma unroll
  for (int nt = 0; nt < 2; ++nt)
#pragma unroll
    for (int e = 0; e < 2; ++e) {
      const long long ci = (long long)n * p.C + o0 + nt * 8 + (lane & 3) * 2 + e;
      f[nt][e] = (p.row_scale ? p.row_scale[ci] : 1.f) * (p.post_scale ? p.post_scale[ci] : 1.f);
      dot[nt][e] = 0.f;
    }
  // all gate values first (independent loads in flight together), then the math and the reds
  uint32_t gv[BD_MT][2][2];
  if (p.gate.ptr) {
#pragma unroll
    for (int mt = 0; mt < BD_MT; ++mt)
#pragma unroll
      for (int hf = 0; hf < 2; ++hf) {
        const int pos = p0 + mt * 16 + (lane >> 2) + 8 * hf;
        int ta, tb;
        ring_target(line, pos, p.H, p.W, ta, tb);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt)
          gv[mt][hf][nt] = pos < L ? *reinterpret_cast<const uint32_t*>(vptr<__nv_bfloat16>(
                                         p.gate, n, ta, tb, o0 + nt * 8 + (lane & 3) * 2))
                                   : 0u;
      }
  }
#pragma unroll
  for (int mt = 0; mt < BD_MT; ++mt)
#pragma unroll
    for (int hf = 0; hf < 2; ++hf) {
      const int pos = p0 + mt * 16 + (lane >> 2) + 8 * hf;
      if (pos < L) {
        int ta, tb;
        ring_target(line, pos, p.H, p.W, ta, tb);
#pragma unroll
        for (int nt = 0; nt < 2; ++nt) {
          const int ch = o0 + nt * 8 + (lane & 3) * 2;
          float v0 = acc[mt][nt][2 * hf], v1 = acc[mt][nt][2 * hf + 1];
          if (p.gate.ptr) {
            const float2 g = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(&gv[mt][hf][nt]));
            dot[nt][0] = fmaf(v0, g.x, dot[nt][0]);
            dot[nt][1] = fmaf(v1, g.y, dot[nt][1]);
            v0 = g.x != 0.f ? v0 : 0.f;
            v1 = g.y != 0.f ? v1 : 0.f;
          }
          // fire-and-forget packed reduction (atomicAdd on a generic pointer compiled to a
          // returning ATOM plus a shared-memory CAS fallback)
          const __nv_bfloat162 hv = __floats2bfloat162_rn(v0 * f[nt][0], v1 * f[nt][1]);
          asm volatile("red.global.add.noftz.bf16x2 [%0], %1;" ::"l"(
                           vptr_mut<__nv_bfloat16>(p.y, n, ta, tb, ch)),
                       "r"(*reinterpret_cast<const uint32_t*>(&hv))
                       : "memory");
        }
      }
    }
  if (p.dot_sums) {
#pragma unroll
    for (int nt = 0; nt < 2; ++nt)
#pragma unroll
      for (int e = 0; e < 2; ++e) {
        float d = dot[nt][e];
        d += __shfl_xor_sync(0xffffffffu, d, 4);
        d += __shfl_xor_sync(0xffffffffu, d, 8);
        d += __shfl_xor_sync(0xffffffffu, d, 16);
        if (lane < 4)
          atomicAdd(p.dot_sums + (long long)n * p.C + o0 + nt * 8 + (lane & 3) * 2 + e, d);
      }
  }
}

static bool bd_tensor_ok(const otm_tensor& t) {
  return t.dtype == OTM_BF16 && t.c % 8 == 0 && t.sw % 8 == 0 && t.sh % 8 == 0 && t.sn % 8 == 0 &&
         ((uintptr_t)t.ptr % 16 == 0);
}

}  // namespace otm

using namespace otm;

extern "C" int otm_conv_reflect_border(const otm_conv_reflect_border_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->dy.ptr && a->y.ptr && a->wpack, "conv_reflect_border: null argument");
  OTM_REQUIRE(a->dy.n == a->y.n && a->dy.h == a->y.h && a->dy.w == a->y.w,
              "conv_reflect_border: dy %dx%dx%d does not match y %dx%dx%d", a->dy.n, a->dy.h, a->dy.w,
              a->y.n, a->y.h, a->y.w);
  OTM_REQUIRE(a->y.h >= 3 && a->y.w >= 3, "conv_reflect_border: image smaller than 3x3");
  OTM_REQUIRE(bd_tensor_ok(a->dy) && bd_tensor_ok(a->y), "conv_reflect_border: bf16 NHWC tensors with "
              "16-byte aligned pixels required");
  OTM_REQUIRE(a->dy.c % 64 == 0 && a->y.c % BD_N == 0,
              "conv_reflect_border: K (%d) must be a multiple of 64 and Cout (%d) of %d", a->dy.c,
              a->y.c, BD_N);
  OTM_REQUIRE((uintptr_t)a->wpack % 4 == 0, "conv_reflect_border: unaligned pack");
  OTM_REQUIRE(a->w_batch_stride == 0 || a->w_batch_stride == 9LL * a->dy.c * a->y.c,
              "conv_reflect_border: bad w_batch_stride");
  if (a->gate.ptr)
    OTM_REQUIRE(bd_tensor_ok(a->gate) && a->gate.n == a->y.n && a->gate.h == a->y.h &&
                    a->gate.w == a->y.w && a->gate.c == a->y.c,
                "conv_reflect_border: gate mismatch");
  OTM_REQUIRE(!a->dot_sums || a->gate.ptr, "conv_reflect_border: dot_sums needs a gate tensor");
  BorderP p;
  p.x = make_view(a->dy);
  p.y = make_view(a->y);
  p.gate = a->gate.ptr ? make_view(a->gate) : null_view();
  p.wp = (const __nv_bfloat16*)a->wpack;
  p.w_batch_stride = a->w_batch_stride;
  p.row_scale = a->row_scale;
  p.post_scale = a->post_scale;
  p.dot_sums = a->dot_sums;
  p.K = a->dy.c; p.C = a->y.c; p.H = a->y.h; p.W = a->y.w;
  const int lmax = (p.W + 2 > p.H) ? p.W + 2 : p.H;
  p.mchunks = (lmax + BD_M - 1) / BD_M;
  // staged line rows: positions of the fullest chunk + 2 tap rows (+ 1 for vertical lines)
  auto rows_for = [](int L, int extra) { int m = ((L < BD_M ? L : BD_M) + 15) / 16 * 16; return m + 2 + extra; };
  const int rh = rows_for(p.W + 2, 0), rv = rows_for(p.H, 1);
  p.line_rows = rh > rv ? rh : rv;
  const int smem = (p.line_rows + BD_N) * (p.K * 2);  // dy line + one weight tap
  OTM_REQUIRE(smem <= 224 * 1024, "conv_reflect_border: K = %d too large", p.K);
  auto kern = conv_reflect_border_kernel;
  if (smem > 48 * 1024) OTM_ENSURE_SMEM(kern, 224 * 1024);  // one ceiling: the attribute is not additive
  {
    // 768 CTAs (n = 96) must be ONE wave: 6 CTAs per SM need the large shared-memory carve-out
    // (ncu showed 5 per SM by shared memory AND by registers: a 28-CTA second wave doubled the time)
    static std::atomic<unsigned long long> carve_done{0};
    int dev = 0;
    OTM_CHECK_CUDA(cudaGetDevice(&dev));
    const unsigned long long bit = 1ull << (dev & 63);
    if (!(carve_done.load(std::memory_order_acquire) & bit)) {
      OTM_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout,
                                          cudaSharedmemCarveoutMaxShared));
      carve_done.fetch_or(bit, std::memory_order_release);
    }
  }
  dim3 grid(4 * p.mchunks, p.C / BD_N, p.y.n);
  kern<<<grid, 128, smem, st>>>(p);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}
