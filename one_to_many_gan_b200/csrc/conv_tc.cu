// tcgen05 / TMEM / TMA implicit-GEMM convolution for sm_100a (bf16 operands, fp32 accumulate).
//
// Forward and dgrad (same kernel, dgrad = flipped/transposed weight pack):
//   GEMM  M = 128 output pixels (a TH x TW rectangle of one image), N = BN output channels,
//         K = kh*kw*Cin walked as (filter tap, 64-channel chunk).
//   A operand: one 4-D TMA box {64 ch, TW, TH, 1} per (tap, chunk) straight out of the NHWC
//         activation (halo included in the tensor map; TMA zero-fills out-of-bounds = the
//         conv's zero padding) -> 128 rows x 128 B, SWIZZLE_128B = UMMA K-major canonical.
//   B operand: 2-D TMA box {64 k, BN rows} of the packed weights [Cout][kh][kw][Cin].
//   D: fp32 accumulator in TMEM (BN columns x 128 lanes), read back with tcgen05.ld by four
//      epilogue warps which apply alpha, demodulation row-scale, bias, activation, residual
//      and write bf16 NHWC (plus the reflect halo the consumer conv needs).
// Wgrad:
//   GEMM  M = 128 (Cout or Cin), N = BN (the other), K = output pixels of one sample chunk.
//   Both operands are MN-major (channels contiguous, pixel = K strided): the same TMA boxes
//   ({64 ch, 8, 8, 1} = 64 pixels x 128 B) read through MN-major SWIZZLE_128B descriptors.
//   Split-K over (sample, pixel chunk); the epilogue applies the per-sample modulation
//   factors rs[n,o]*cs[n,i] and reduces into the fp32 weight gradient with red.global.add.
//
// Warp roles (256 threads): warp0 = TMA producer, warp1 = MMA issuer (one elected lane),
// warp2 = TMEM allocator, warps4-7 = epilogue (TMEM lane quarter = warp_idx % 4).
#include <cuda.h>
#include <cstdlib>

#include "common.cuh"

namespace otm {

// ---------------------------------------------------------------------------
// PTX wrappers
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
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
__device__ __forceinline__ void fence_barrier_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_before() {
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tc_fence_after() {
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1, int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar,
                                            int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes"
      " [%0], [%1, {%3, %4}], [%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, uint32_t src, int c0, int c1,
                                             int c2, int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%2, %3, %4, %5}], [%1];" ::"l"(map),
      "r"(src), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_store_commit() {
  asm volatile("cp.async.bulk.commit_group;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read() {  // staged smem may be overwritten
  asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_read1() {  // all but the newest group were read
  asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
}
__device__ __forceinline__ void tma_store_wait_all() {
  asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
  asm volatile("prefetch.tensormap [%0];" ::"l"(map) : "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_alloc(uint32_t dst_smem) {
  asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem),
               "n"(NCOLS)
               : "memory");
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
template <int NCOLS>
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr) {
  asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "n"(NCOLS)
               : "memory");
}
__device__ __forceinline__ void umma_bf16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b,
                                          uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n"
      ".reg .pred p;\n"
      "setp.ne.b32 p, %4, 0;\n"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
      "}\n" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                   bar)
               : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}

// 32 columns per instruction: halves the number of (serialising) TMEM round trips of the epilogue
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]),
        "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
        "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}

// 16 lanes x 32 columns in the mma C-fragment layout: thread t receives, for every 8-column
// block k, r[4k+0..1] = (lane t/4, columns 8k + 2(t%4) + {0,1}) and r[4k+2..3] = the same
// columns of lane t/4 + 8 -- exactly the register layout stmatrix stores, so an accumulator
// that is TRANSPOSED in TMEM (lane = channel, column = pixel) goes to a pixel-major shared
// tile with one stmatrix.trans per 4 (8 channel x 8 pixel) blocks instead of 2-byte stores.
__device__ __forceinline__ void tmem_ld_16x256b_x4(uint32_t taddr, float (&v)[16]) {
  uint32_t r[16];
  asm volatile(
      "tcgen05.ld.sync.aligned.16x256b.x4.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
        "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
        "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
#pragma unroll
  for (int i = 0; i < 16; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_ld_wait() {
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 t = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float2 bf16x2_to_float2(uint32_t v) {
  return __bfloat1622float2(*reinterpret_cast<__nv_bfloat162*>(&v));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t saddr, uint32_t (&r)[4]) {
  asm volatile("ldmatrix.sync.aligned.x4.trans.m8n8.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(saddr)
               : "memory");
}
__device__ __forceinline__ uint32_t hadd2_bf16(uint32_t a, uint32_t b) {
  __nv_bfloat162 r = __hadd2(*reinterpret_cast<__nv_bfloat162*>(&a), *reinterpret_cast<__nv_bfloat162*>(&b));
  return *reinterpret_cast<uint32_t*>(&r);
}
// four transposed 8x8 b16 matrices; lanes 8j .. 8j+7 give the row addresses of matrix j
__device__ __forceinline__ void stmatrix_x4_trans(uint32_t saddr, uint32_t a, uint32_t b, uint32_t c,
                                                  uint32_t d) {
  asm volatile("stmatrix.sync.aligned.x4.trans.m8n8.shared.b16 [%0], {%1, %2, %3, %4};" ::"r"(saddr),
               "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}

// shared-memory matrix descriptor, SWIZZLE_128B, sm_100 version field = 1
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes,
                                              uint32_t sbo_bytes) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
  d |= (uint64_t)((lbo_bytes >> 4) & 0x3FFF) << 16;
  d |= (uint64_t)((sbo_bytes >> 4) & 0x3FFF) << 32;
  d |= (uint64_t)1 << 46;  // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;  // SWIZZLE_128B
  return d;
}
// instruction descriptor: bf16 x bf16 -> fp32
__host__ __device__ constexpr uint32_t make_idesc(int M, int N, int a_mn_major, int b_mn_major) {
  return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)a_mn_major << 15) |
         ((uint32_t)b_mn_major << 16) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}

// ---------------------------------------------------------------------------
// forward / dgrad kernel
// ---------------------------------------------------------------------------
struct TcFwdP {
  View y, res;
  const float* row_scale;
  const float* post_scale;  // [n, cout] applied AFTER the activation (NULL = 1)
  float* stat_sums;         // [n, cout, 2] sum / sum of squares of the stored values, or NULL
  float* dot_sums;          // [n, cout] sum_hw v * residual (gate mode of the transposed pair kernel)
  int res_mode;             // 0: add the residual; 1: gate by it; 2: dot only (otm_conv_fwd_args.residual_mode)
  const float* bias;
  float alpha;
  int act, y_halo;
  int cin, cout, kh, kw;
  int coord_off;  // x_halo - pad : added to (h0 + r), (w0 + s) to index the halo'd tensor map
  int TW, TH, tiles_w;
  int w_rows_per_sample;  // cout if per-sample packs else 0
};

constexpr int A_STAGE_BYTES = 128 * 128;  // 128 pixels x 64 bf16

// Epilogue math for 16 consecutive output channels of one pixel, result packed to bf16 and
// written to the shared-memory staging tile (row = pixel, pitch BN*2+16 bytes: conflict-free
// for the 16-byte per-lane stores of a quarter warp).
template <int BN>
__device__ __forceinline__ void epilogue_math16(const TcFwdP& p, int n, int oh, int ow, int o,
                                                bool valid, float (&v)[16], uint8_t* dst,
                                                const float* sc /* smem: alpha*row_scale */,
                                                const float* bi /* smem: bias */,
                                                const float* po /* smem: post-activation scale */,
                                                uint8_t* dst_hi = nullptr) {
  {
    const float4* rs = reinterpret_cast<const float4*>(sc);
    const float4* bs = reinterpret_cast<const float4*>(bi);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      float4 t = rs[q], u = bs[q];
      v[4 * q] = fmaf(v[4 * q], t.x, u.x);
      v[4 * q + 1] = fmaf(v[4 * q + 1], t.y, u.y);
      v[4 * q + 2] = fmaf(v[4 * q + 2], t.z, u.z);
      v[4 * q + 3] = fmaf(v[4 * q + 3], t.w, u.w);
    }
  }
  act_fwd_vec<16>(v, p.act);
  if (p.post_scale) {
    const float4* ps = reinterpret_cast<const float4*>(po);
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const float4 t = ps[q];
      v[4 * q] *= t.x; v[4 * q + 1] *= t.y; v[4 * q + 2] *= t.z; v[4 * q + 3] *= t.w;
    }
  }
  if (p.res.ptr && valid) {
    const __nv_bfloat16* rp = vptr<__nv_bfloat16>(p.res, n, oh, ow, o);
    float r0[8], r1[8];
    load_vec<__nv_bfloat16, 8>(rp, r0);
    load_vec<__nv_bfloat16, 8>(rp + 8, r1);
#pragma unroll
    for (int i = 0; i < 8; ++i) { v[i] += r0[i]; v[8 + i] += r1[i]; }
  }
  uint4 lo, hi;
  __nv_bfloat162* l2 = reinterpret_cast<__nv_bfloat162*>(&lo);
  __nv_bfloat162* h2 = reinterpret_cast<__nv_bfloat162*>(&hi);
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    l2[i] = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
    h2[i] = __floats2bfloat162_rn(v[8 + 2 * i], v[8 + 2 * i + 1]);
  }
  *reinterpret_cast<uint4*>(dst) = lo;
  *reinterpret_cast<uint4*>(dst_hi ? dst_hi : dst + 16) = hi;
}

// ---------------------------------------------------------------------------
// Persistent forward / dgrad kernel: one CTA per SM loops over output tiles; the fp32
// accumulator is double-buffered in TMEM (2 x BN columns), so the epilogue of tile i
// (TMEM read-out, epilogue math, staged + coalesced write-out) overlaps the TMA/MMA main
// loop of tile i+1, and barrier init / TMEM allocation / descriptor prefetch happen once.
// Measured motivation (profiles/r1_epilogue_ablation.md): the main loop alone sustains
// ~1.05 PFLOP/s on the 128-channel layers, the serialised epilogue brought that to 0.64.
// ---------------------------------------------------------------------------
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

struct TcFwdPP {
  TcFwdP p;
  int total_tiles;     // batch * pixel tiles * cout tiles
  int tiles_per_img;   // pixel tiles per image
  int cout_tiles;
};

// ---------------------------------------------------------------------------
// Row-reuse variant of the persistent kernel (square KxK filters, K = 3 or 4).
// The main loop is bound by L2->SM operand traffic (16 KB of A + 16 KB of B per 256 MMA
// cycles), so the A operand is loaded ONCE per (filter column s, channel chunk) as a
// (16+K-1) x 8-pixel box and reused for the K filter rows: with an 8-pixel-wide tile one image
// row of the box is exactly one 1024-byte swizzle atom, so the A descriptor of filter row r is
// the same box advanced by r*1024 bytes.  A traffic drops K-fold (9 -> 3 boxes of 18 KB per
// chunk for 3x3).  A and B live in separate rings with their own full/empty barriers.
// ---------------------------------------------------------------------------
template <int BN, int KS, int NA, int NB>
__global__ void __launch_bounds__(256, 1)
conv_tc_fwd_rr_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmY, TcFwdPP pp) {
  const TcFwdP& p = pp.p;
  constexpr int TW = 8, TH = 16;
  constexpr int A_SLOT = (TH + KS - 1) * TW * 128;  // bytes
  constexpr int B_SLOT = BN * 128;
  constexpr int SUB_BYTES = 128 * 128;
  constexpr int TILE_BYTES = (BN / 64) * SUB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ringA = smem;
  uint8_t* ringB = ringA + NA * A_SLOT;
  uint8_t* stage_out = ringB + NB * B_SLOT;
  uint64_t* bars = (uint64_t*)(stage_out + TILE_BYTES);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + NA;
  uint64_t* fullB = emptyA + NA;
  uint64_t* emptyB = fullB + NB;
  uint64_t* tmem_full = emptyB + NB;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);
  float* s_scale = (float*)(((uintptr_t)(tmem_ptr + 2) + 15) & ~(uintptr_t)15);
  float* s_bias = s_scale + BN;
  float* s_post = s_bias + BN;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int cin_chunks = p.cin / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(smem_u32(&fullA[i]), 1); mbar_init(smem_u32(&emptyA[i]), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(smem_u32(&fullB[i]), 1); mbar_init(smem_u32(&emptyB[i]), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<2 * BN>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto decode = [&](int t, int& n, int& h0, int& w0, int& o0) {
    const int ct = t % pp.cout_tiles;
    const int rest = t / pp.cout_tiles;
    const int pt = rest % pp.tiles_per_img;
    n = rest / pp.tiles_per_img;
    h0 = (pt / p.tiles_w) * TH;
    w0 = (pt % p.tiles_w) * TW;
    o0 = ct * BN;
  };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int ga = 0, gb = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x) {
      int n, h0, w0, o0;
      decode(t, n, h0, w0, o0);
      const int wrow = n * p.w_rows_per_sample + o0;
      for (int cc = 0; cc < cin_chunks; ++cc) {
        for (int s = 0; s < KS; ++s, ++ga) {
          const int sa = ga % NA;
          mbar_wait(smem_u32(&emptyA[sa]), ((ga / NA) & 1) ^ 1);
          if (lane == 0) {
            const uint32_t bar = smem_u32(&fullA[sa]);
            mbar_expect_tx(bar, A_SLOT);
            tma_load_4d(smem_u32(ringA + sa * A_SLOT), &tmA, bar, cc * 64, w0 + s + p.coord_off,
                        h0 + p.coord_off, n);
          }
          __syncwarp();
          for (int r = 0; r < KS; ++r, ++gb) {
            const int sb = gb % NB;
            mbar_wait(smem_u32(&emptyB[sb]), ((gb / NB) & 1) ^ 1);
            if (lane == 0) {
              const uint32_t bar = smem_u32(&fullB[sb]);
              mbar_expect_tx(bar, B_SLOT);
              tma_load_2d(smem_u32(ringB + sb * B_SLOT), &tmB, bar, (r * KS + s) * p.cin + cc * 64, wrow);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    int ga = 0, gb = 0, lt = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      mbar_wait(smem_u32(&tmem_empty[buf]), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * BN);
      uint32_t first = 1;
      for (int cc = 0; cc < cin_chunks; ++cc) {
        for (int s = 0; s < KS; ++s, ++ga) {
          const int sa = ga % NA;
          mbar_wait(smem_u32(&fullA[sa]), (ga / NA) & 1);
          for (int r = 0; r < KS; ++r, ++gb) {
            const int sb = gb % NB;
            mbar_wait(smem_u32(&fullB[sb]), (gb / NB) & 1);
            tc_fence_after();
            if (lane == 0) {
              const uint64_t da = make_desc(smem_u32(ringA + sa * A_SLOT + r * (TW * 128)), 16, 1024);
              const uint64_t db = make_desc(smem_u32(ringB + sb * B_SLOT), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(tacc, da + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                first = 0;
              }
              umma_commit(smem_u32(&emptyB[sb]));
              if (r == KS - 1) umma_commit(smem_u32(&emptyA[sa]));
              if (r == KS - 1 && s == KS - 1 && cc == cin_chunks - 1)
                umma_commit(smem_u32(&tmem_full[buf]));
            }
            first = 0;
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const int wq = warp - 4;
    const int te = threadIdx.x - 128;
    const int m = wq * 32 + lane;
    int lt = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x, ++lt) {
      int n, h0, w0, o0;
      decode(t, n, h0, w0, o0);
      const int buf = lt & 1;
      for (int c = te; c < BN; c += 128) {
        s_scale[c] = p.alpha * (p.row_scale ? p.row_scale[(long long)n * p.cout + o0 + c] : 1.f);
        s_bias[c] = p.bias ? p.bias[o0 + c] : 0.f;
        s_post[c] = p.post_scale ? p.post_scale[(long long)n * p.cout + o0 + c] : 1.f;
      }
      if (te == 0) tma_store_wait_read();
      asm volatile("bar.sync 2, 128;" ::: "memory");
      mbar_wait(smem_u32(&tmem_full[buf]), (lt >> 1) & 1);
      tc_fence_after();
      const int oh = h0 + m / TW, ow = w0 + m % TW;
      const bool valid = (oh < p.y.h) && (ow < p.y.w);
      const uint32_t tacc = tmem_base + (uint32_t)(buf * BN) + ((uint32_t)(wq * 32) << 16);
      uint8_t* myrow = stage_out + m * 128;
      const int sw = m & 7;
#pragma unroll 1
      for (int j = 0; j < BN / 16; ++j) {
        float v[16];
        tmem_ld16(tacc + (uint32_t)(j * 16), v);
        uint8_t* sub = myrow + (j >> 2) * SUB_BYTES;
        const int c = (j & 3) * 2;
        epilogue_math16<BN>(p, n, oh, ow, o0 + j * 16, valid, v, sub + ((c ^ sw) << 4),
                            s_scale + j * 16, s_bias + j * 16, s_post + j * 16,
                            sub + (((c + 1) ^ sw) << 4));
      }
      tc_fence_before();
      mbar_arrive(smem_u32(&tmem_empty[buf]));
      fence_proxy_async();
      asm volatile("bar.sync 1, 128;" ::: "memory");
      if (te == 0) {
#pragma unroll
        for (int sb = 0; sb < BN / 64; ++sb)
          tma_store_4d(&tmY, smem_u32(stage_out + sb * SUB_BYTES), o0 + sb * 64, w0, h0, n);
        tma_store_commit();
      }
      if (p.y_halo > 0 && valid) {
        int hs[3], ws[3];
        const int nh = mirror_set(oh, p.y.h, p.y_halo, hs);
        const int nw = mirror_set(ow, p.y.w, p.y_halo, ws);
        if (nh * nw > 1) {
          for (int piece = 0; piece < BN / 8; ++piece) {
            const uint4 val = *reinterpret_cast<const uint4*>(
                myrow + (piece >> 3) * SUB_BYTES + (((piece & 7) ^ sw) << 4));
            for (int a = 0; a < nh; ++a)
              for (int b = 0; b < nw; ++b)
                if (a + b > 0)
                  *reinterpret_cast<uint4*>(
                      vptr_mut<__nv_bfloat16>(p.y, n, hs[a], ws[b], o0 + piece * 8)) = val;
          }
        }
      }
    }
    if (te == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<2 * BN>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// Pair variant of the row-reuse kernel: one CTA works on TWO vertically adjacent 16x8 pixel
// tiles at once (two fp32 accumulators in TMEM, double-buffered: 4*BN columns), so every B
// (weight) tile fetched from L2 feeds two MMAs.  Operand traffic per 2 tiles and 64-channel
// chunk: 6 A boxes (108 KB) + 9 B tiles (144 KB) instead of 2 x (54 + 144) KB.
// ---------------------------------------------------------------------------
template <int BN, int KS, int NA, int NB>
__global__ void __launch_bounds__(256, 1)
conv_tc_fwd_rr2_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                      const __grid_constant__ CUtensorMap tmY, TcFwdPP pp) {
  const TcFwdP& p = pp.p;
  constexpr int TW = 8, TH = 16;
  constexpr int A_BOX = (TH + KS - 1) * TW * 128;  // bytes, one pixel tile
  constexpr int A_SLOT = 2 * A_BOX;                // the pair of vertically adjacent tiles
  constexpr int B_SLOT = BN * 128;
  constexpr int SUB_BYTES = 128 * 128;
  constexpr int TILE_BYTES = (BN / 64) * SUB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ringA = smem;
  uint8_t* ringB = ringA + NA * A_SLOT;
  uint8_t* stage_out = ringB + NB * B_SLOT;
  uint64_t* bars = (uint64_t*)(stage_out + TILE_BYTES);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + NA;
  uint64_t* fullB = emptyA + NA;
  uint64_t* emptyB = fullB + NB;
  uint64_t* tmem_full = emptyB + NB;   // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint32_t* tmem_ptr = (uint32_t*)(tmem_empty + 2);
  float* s_scale = (float*)(((uintptr_t)(tmem_ptr + 2) + 15) & ~(uintptr_t)15);
  float* s_bias = s_scale + BN;
  float* s_post = s_bias + BN;

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int cin_chunks = p.cin / 64;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(smem_u32(&fullA[i]), 1); mbar_init(smem_u32(&emptyA[i]), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(smem_u32(&fullB[i]), 1); mbar_init(smem_u32(&emptyB[i]), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<4 * BN>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto decode = [&](int t, int& n, int& h0, int& w0, int& o0) {
    const int ct = t % pp.cout_tiles;
    const int rest = t / pp.cout_tiles;
    const int pt = rest % pp.tiles_per_img;
    n = rest / pp.tiles_per_img;
    h0 = (pt / p.tiles_w) * (2 * TH);  // pt indexes PAIRS of tile rows
    w0 = (pt % p.tiles_w) * TW;
    o0 = ct * BN;
  };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int ga = 0, gb = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x) {
      int n, h0, w0, o0;
      decode(t, n, h0, w0, o0);
      const int wrow = n * p.w_rows_per_sample + o0;
      const bool two = (h0 + TH) < p.y.h;  // the lower tile of the pair has output rows
      for (int cc = 0; cc < cin_chunks; ++cc) {
        for (int s = 0; s < KS; ++s, ++ga) {
          const int sa = ga % NA;
          mbar_wait(smem_u32(&emptyA[sa]), ((ga / NA) & 1) ^ 1);
          if (lane == 0) {
            const uint32_t bar = smem_u32(&fullA[sa]);
            mbar_expect_tx(bar, two ? A_SLOT : A_BOX);
            tma_load_4d(smem_u32(ringA + sa * A_SLOT), &tmA, bar, cc * 64, w0 + s + p.coord_off,
                        h0 + p.coord_off, n);
            if (two)
              tma_load_4d(smem_u32(ringA + sa * A_SLOT + A_BOX), &tmA, bar, cc * 64,
                          w0 + s + p.coord_off, h0 + TH + p.coord_off, n);
          }
          __syncwarp();
          for (int r = 0; r < KS; ++r, ++gb) {
            const int sb = gb % NB;
            mbar_wait(smem_u32(&emptyB[sb]), ((gb / NB) & 1) ^ 1);
            if (lane == 0) {
              const uint32_t bar = smem_u32(&fullB[sb]);
              mbar_expect_tx(bar, B_SLOT);
              tma_load_2d(smem_u32(ringB + sb * B_SLOT), &tmB, bar, (r * KS + s) * p.cin + cc * 64, wrow);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc(128, BN, 0, 0);
    int ga = 0, gb = 0, lt = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      mbar_wait(smem_u32(&tmem_empty[buf]), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * 2 * BN);  // [tile0 | tile1]
      uint32_t first = 1;
      bool two;
      {
        int n, h0, w0, o0;
        decode(t, n, h0, w0, o0);
        two = (h0 + TH) < p.y.h;
      }
      for (int cc = 0; cc < cin_chunks; ++cc) {
        for (int s = 0; s < KS; ++s, ++ga) {
          const int sa = ga % NA;
          mbar_wait(smem_u32(&fullA[sa]), (ga / NA) & 1);
          for (int r = 0; r < KS; ++r, ++gb) {
            const int sb = gb % NB;
            mbar_wait(smem_u32(&fullB[sb]), (gb / NB) & 1);
            tc_fence_after();
            if (lane == 0) {
              const uint64_t da0 = make_desc(smem_u32(ringA + sa * A_SLOT + r * (TW * 128)), 16, 1024);
              const uint64_t da1 =
                  make_desc(smem_u32(ringA + sa * A_SLOT + A_BOX + r * (TW * 128)), 16, 1024);
              const uint64_t db = make_desc(smem_u32(ringB + sb * B_SLOT), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {  // the B tile is read once for both pixel tiles
                umma_bf16(tacc, da0 + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                if (two)
                  umma_bf16(tacc + BN, da1 + (uint64_t)(k * 2), db + (uint64_t)(k * 2), idesc,
                            first ? 0u : 1u);
                first = 0;
              }
              umma_commit(smem_u32(&emptyB[sb]));
              if (r == KS - 1) umma_commit(smem_u32(&emptyA[sa]));
              if (r == KS - 1 && s == KS - 1 && cc == cin_chunks - 1)
                umma_commit(smem_u32(&tmem_full[buf]));
            }
            first = 0;
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue ----------------
    const int wq = warp - 4;
    const int te = threadIdx.x - 128;
    const int m = wq * 32 + lane;
    int lt = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x, ++lt) {
      int n, h0, w0, o0;
      decode(t, n, h0, w0, o0);
      const int buf = lt & 1;
      for (int c = te; c < BN; c += 128) {
        s_scale[c] = p.alpha * (p.row_scale ? p.row_scale[(long long)n * p.cout + o0 + c] : 1.f);
        s_bias[c] = p.bias ? p.bias[o0 + c] : 0.f;
        s_post[c] = p.post_scale ? p.post_scale[(long long)n * p.cout + o0 + c] : 1.f;
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");  // s_scale / s_bias visible
      mbar_wait(smem_u32(&tmem_full[buf]), (lt >> 1) & 1);
      tc_fence_after();
      uint8_t* myrow = stage_out + m * 128;
      const int sw = m & 7;
      const int nhalf = (h0 + TH) < p.y.h ? 2 : 1;  // a fully out-of-range lower tile is skipped
#pragma unroll 1
      for (int half = 0; half < nhalf; ++half) {
        const int hh0 = h0 + half * TH;
        const int oh = hh0 + m / TW, ow = w0 + m % TW;
        const bool valid = (oh < p.y.h) && (ow < p.y.w);
        const uint32_t tacc =
            tmem_base + (uint32_t)(buf * 2 * BN + half * BN) + ((uint32_t)(wq * 32) << 16);
        // the previous TMA store must have finished reading the staging tile
        if (te == 0) tma_store_wait_read();
        asm volatile("bar.sync 3, 128;" ::: "memory");
#pragma unroll 1
        for (int j2 = 0; j2 < BN / 32; ++j2) {
          float v32[32];
          tmem_ld32(tacc + (uint32_t)(j2 * 32), v32);
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const int j = j2 * 2 + q;
            float v[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) v[i] = v32[q * 16 + i];
            uint8_t* sub = myrow + (j >> 2) * SUB_BYTES;
            const int c = (j & 3) * 2;
            epilogue_math16<BN>(p, n, oh, ow, o0 + j * 16, valid, v, sub + ((c ^ sw) << 4),
                                s_scale + j * 16, s_bias + j * 16, s_post + j * 16,
                            sub + (((c + 1) ^ sw) << 4));
          }
        }
        tc_fence_before();
        if (half == nhalf - 1) mbar_arrive(smem_u32(&tmem_empty[buf]));  // accumulators drained
        fence_proxy_async();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (te == 0 && hh0 < p.y.h) {
#pragma unroll
          for (int sb = 0; sb < BN / 64; ++sb)
            tma_store_4d(&tmY, smem_u32(stage_out + sb * SUB_BYTES), o0 + sb * 64, w0, hh0, n);
          tma_store_commit();
        }
        if (p.y_halo > 0 && valid) {
          int hs[3], ws[3];
          const int nh = mirror_set(oh, p.y.h, p.y_halo, hs);
          const int nw = mirror_set(ow, p.y.w, p.y_halo, ws);
          if (nh * nw > 1) {
            for (int a = 0; a < nh; ++a)
              for (int b = 0; b < nw; ++b)
                if (a + b > 0) {
                  uint4* dst = reinterpret_cast<uint4*>(
                      vptr_mut<__nv_bfloat16>(p.y, n, hs[a], ws[b], o0));
#pragma unroll 4
                  for (int piece = 0; piece < BN / 8; ++piece)
                    dst[piece] = *reinterpret_cast<const uint4*>(
                        myrow + (piece >> 3) * SUB_BYTES + (((piece & 7) ^ sw) << 4));
                }
          }
        }
      }
    }
    if (te == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<4 * BN>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// Transposed pair kernel ("weights as the M operand") for BN = 128 output channels.
// The pair kernel above issues two 128x128x16 MMAs per weight tile; each reads 4 KB + 4 KB of
// operands from shared memory per 64 tensor-pipe cycles = 128 B/clk, which IS the SM's
// shared-memory bandwidth: its main loop saturates at ~1.1 PFLOP/s (the 256-channel layers, N =
// 256, reach 1.2-1.4).  Here D^T = W * X^T: M = 128 output channels (the weight tile, K-major),
// N = 256 pixels = the two vertically adjacent 16x8 tiles loaded as ONE (32 + KS - 1)-row box,
// so one 128x256x16 MMA reads 4 KB + 8 KB per 128 cycles = 96 B/clk.
// The accumulator is transposed (TMEM lane = output channel, column = pixel): the epilogue
// applies the per-channel scale/bias as per-thread scalars and transposes through the
// SWIZZLE_128B staging tile with 2-byte stores; residual add, TMA store and the reflect halo
// then work on staged pixel rows.
// ---------------------------------------------------------------------------
template <int KS, int NA, int NB>
__global__ void __launch_bounds__(256, 1)
conv_tc_fwd_rr2t_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                        const __grid_constant__ CUtensorMap tmY, const __grid_constant__ CUtensorMap tmR,
                        TcFwdPP pp) {
  const TcFwdP& p = pp.p;
  constexpr int BN = 128;  // output channels per tile (the MMA's M)
  constexpr int TW = 8, TH = 16;
  constexpr int A_SLOT = (2 * TH + KS - 1) * TW * 128;  // bytes: the pair's pixel rows, one box
  constexpr int B_SLOT = BN * 128;
  constexpr int SUB_BYTES = 128 * 128;
  constexpr int TILE_BYTES = (BN / 64) * SUB_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* ringA = smem;
  uint8_t* ringB = ringA + NA * A_SLOT;
  uint8_t* stage_base = ringB + NB * B_SLOT;  // TWO staging tiles: half-tiles alternate
  uint64_t* bars = (uint64_t*)(stage_base + 2 * TILE_BYTES);
  uint64_t* fullA = bars;
  uint64_t* emptyA = fullA + NA;
  uint64_t* fullB = emptyA + NA;
  uint64_t* emptyB = fullB + NB;
  uint64_t* tmem_full = emptyB + NB;     // [2]
  uint64_t* tmem_empty = tmem_full + 2;  // [2]
  uint64_t* res_full = tmem_empty + 2;   // [2] residual half-tile landed in staging tile x
  uint32_t* tmem_ptr = (uint32_t*)(res_full + 2);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int cin_chunks = p.cin / 64;
  const bool has_res = p.res.ptr != nullptr;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    tma_prefetch_desc(&tmY);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < NA; ++i) { mbar_init(smem_u32(&fullA[i]), 1); mbar_init(smem_u32(&emptyA[i]), 1); }
    for (int i = 0; i < NB; ++i) { mbar_init(smem_u32(&fullB[i]), 1); mbar_init(smem_u32(&emptyB[i]), 1); }
    for (int b = 0; b < 2; ++b) {
      mbar_init(smem_u32(&tmem_full[b]), 1);
      mbar_init(smem_u32(&tmem_empty[b]), 128);
      mbar_init(smem_u32(&res_full[b]), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  auto decode = [&](int t, int& n, int& h0, int& w0, int& o0) {
    const int ct = t % pp.cout_tiles;
    const int rest = t / pp.cout_tiles;
    const int pt = rest % pp.tiles_per_img;
    n = rest / pp.tiles_per_img;
    h0 = (pt / p.tiles_w) * (2 * TH);  // pt indexes PAIRS of tile rows
    w0 = (pt % p.tiles_w) * TW;
    o0 = ct * BN;
  };

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int ga = 0, gb = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x) {
      int n, h0, w0, o0;
      decode(t, n, h0, w0, o0);
      const int wrow = n * p.w_rows_per_sample + o0;
      for (int cc = 0; cc < cin_chunks; ++cc) {
        for (int s = 0; s < KS; ++s, ++ga) {
          const int sa = ga % NA;
          mbar_wait(smem_u32(&emptyA[sa]), ((ga / NA) & 1) ^ 1);
          if (lane == 0) {
            const uint32_t bar = smem_u32(&fullA[sa]);
            mbar_expect_tx(bar, A_SLOT);
            tma_load_4d(smem_u32(ringA + sa * A_SLOT), &tmA, bar, cc * 64, w0 + s + p.coord_off,
                        h0 + p.coord_off, n);
          }
          __syncwarp();
          for (int r = 0; r < KS; ++r, ++gb) {
            const int sb = gb % NB;
            mbar_wait(smem_u32(&emptyB[sb]), ((gb / NB) & 1) ^ 1);
            if (lane == 0) {
              const uint32_t bar = smem_u32(&fullB[sb]);
              mbar_expect_tx(bar, B_SLOT);
              tma_load_2d(smem_u32(ringB + sb * B_SLOT), &tmB, bar, (r * KS + s) * p.cin + cc * 64, wrow);
            }
            __syncwarp();
          }
        }
      }
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc2 = make_idesc(128, 256, 0, 0);  // both 16x8 tiles
    constexpr uint32_t idesc1 = make_idesc(128, 128, 0, 0);  // lower tile out of range
    int ga = 0, gb = 0, lt = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x, ++lt) {
      const int buf = lt & 1;
      uint32_t idesc;
      {
        int n, h0, w0, o0;
        decode(t, n, h0, w0, o0);
        idesc = (h0 + TH) < p.y.h ? idesc2 : idesc1;
      }
      mbar_wait(smem_u32(&tmem_empty[buf]), ((lt >> 1) & 1) ^ 1);
      tc_fence_after();
      const uint32_t tacc = tmem_base + (uint32_t)(buf * 256);
      uint32_t first = 1;
      for (int cc = 0; cc < cin_chunks; ++cc) {
        for (int s = 0; s < KS; ++s, ++ga) {
          const int sa = ga % NA;
          mbar_wait(smem_u32(&fullA[sa]), (ga / NA) & 1);
          for (int r = 0; r < KS; ++r, ++gb) {
            const int sb = gb % NB;
            mbar_wait(smem_u32(&fullB[sb]), (gb / NB) & 1);
            tc_fence_after();
            if (lane == 0) {
              const uint64_t dpx = make_desc(smem_u32(ringA + sa * A_SLOT + r * (TW * 128)), 16, 1024);
              const uint64_t dw = make_desc(smem_u32(ringB + sb * B_SLOT), 16, 1024);
#pragma unroll
              for (int k = 0; k < 4; ++k) {
                umma_bf16(tacc, dw + (uint64_t)(k * 2), dpx + (uint64_t)(k * 2), idesc, first ? 0u : 1u);
                first = 0;
              }
              umma_commit(smem_u32(&emptyB[sb]));
              if (r == KS - 1) umma_commit(smem_u32(&emptyA[sa]));
              if (r == KS - 1 && s == KS - 1 && cc == cin_chunks - 1)
                umma_commit(smem_u32(&tmem_full[buf]));
            }
            first = 0;
            __syncwarp();
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------- epilogue warps: TMEM -> scale/bias/activation -> staged bf16 tile -----
    // Named barriers (count 192 = 4 epilogue warps + the 2 finishing warps):
    //   2 + x   stage x is free again (finishing warps arrive, epilogue warps wait)
    //   4 + x   stage x holds a finished half-tile (epilogue warps arrive, finishing warps wait)
    const int wq = warp - 4;
    const bool gate_mode = has_res && p.res_mode != 0;
    const bool nomask = p.res_mode == 2;  // dot only
    int lt = 0, hc = 0;  // tiles / half-tiles done by this CTA
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x, ++lt) {
      int n, h0, w0, o0;
      decode(t, n, h0, w0, o0);
      const int buf = lt & 1;
      // fragment layout: this thread holds channels wq*32 + 8j + lane/4, j = 0..3
      float scale[4], bias[4], post[4];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int ch = o0 + wq * 32 + 8 * j + (lane >> 2);
        scale[j] = p.alpha * (p.row_scale ? p.row_scale[(long long)n * p.cout + ch] : 1.f);
        bias[j] = p.bias ? p.bias[ch] : 0.f;
        post[j] = p.post_scale ? p.post_scale[(long long)n * p.cout + ch] : 1.f;
      }
      const float gsc[4] = {scale[0] * post[0], scale[1] * post[1], scale[2] * post[2],
                            scale[3] * post[3]};
      mbar_wait(smem_u32(&tmem_full[buf]), (lt >> 1) & 1);
      tc_fence_after();
      const int nhalf = (h0 + TH) < p.y.h ? 2 : 1;
      // InstanceNorm statistics of the consumer: per-thread sums of its 4 channels over the tile
      float ssum[4] = {0.f, 0.f, 0.f, 0.f}, ssq[4] = {0.f, 0.f, 0.f, 0.f};
      const bool col_ok0 = w0 + 2 * (lane & 3) < p.y.w, col_ok1 = w0 + 2 * (lane & 3) + 1 < p.y.w;
#pragma unroll 1
      for (int half = 0; half < nhalf; ++half, ++hc) {
        const int x = hc & 1;
        uint8_t* const stage_out = stage_base + x * TILE_BYTES;
        // staging tile [sub-tile of 64 ch][pixel][128 B, 16-byte chunks XOR-swizzled by pixel & 7]:
        // lane = 8j + i addresses row i of matrix j = channels wq*32 + 8j .. + 7 of pixel 8k + i
        const uint32_t st_base = smem_u32(stage_out) + (uint32_t)((wq >> 1) * SUB_BYTES) +
                                 (uint32_t)((lane & 7) * 128) +
                                 (uint32_t)(((((wq & 1) * 4 + (lane >> 3)) ^ (lane & 7))) << 4);
        const uint32_t tacc =
            tmem_base + (uint32_t)(buf * 256 + half * 128) + ((uint32_t)(wq * 32) << 16);
        // 32 pixels (TMEM columns) x this warp's 32 channels per step; the next step's TMEM
        // read is in flight while this one is scaled, packed and stored
        float va[2][16], vb[2][16];  // lanes +0..15 / +16..31 of the warp's TMEM window
        tmem_ld_16x256b_x4(tacc, va[0]);
        tmem_ld_16x256b_x4(tacc + (16u << 16), vb[0]);
        // the finishing warps are done with this staging tile (its TMA store has read it) ...
        asm volatile("bar.sync %0, 192;" ::"r"(2 + x) : "memory");
        // ... and this half-tile's residual has been bulk-copied into it
        if (has_res) mbar_wait(smem_u32(&res_full[x]), (hc >> 1) & 1);
#pragma unroll
        for (int j2 = 0; j2 < 4; ++j2) {
          float (&xa)[16] = va[j2 & 1];
          float (&xb)[16] = vb[j2 & 1];
          tmem_ld_wait();
          if (j2 < 3) {
            tmem_ld_16x256b_x4(tacc + (uint32_t)((j2 + 1) * 32), va[(j2 + 1) & 1]);
            tmem_ld_16x256b_x4(tacc + (uint32_t)((j2 + 1) * 32) + (16u << 16), vb[(j2 + 1) & 1]);
          }
          if (gate_mode) {
            // residual_mode 1: the staged tile gates the result (ReLU backward through h~ = s * ReLU(u),
            // s of either sign: the mask is h~ != 0, as in otm_mod_in) and
            // is reduced against the raw accumulator; same fragment layout as the store below
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const uint32_t sa = st_base + (uint32_t)((j2 * 32 + k * 8) * 128);
              uint32_t rr[4];
              ldmatrix_x4_trans(sa, rr);
              const float2 g0 = bf16x2_to_float2(rr[0]), g1 = bf16x2_to_float2(rr[1]);
              const float2 g2 = bf16x2_to_float2(rr[2]), g3 = bf16x2_to_float2(rr[3]);
              ssum[0] = fmaf(xa[4 * k], g0.x, fmaf(xa[4 * k + 1], g0.y, ssum[0]));
              ssum[1] = fmaf(xa[4 * k + 2], g1.x, fmaf(xa[4 * k + 3], g1.y, ssum[1]));
              ssum[2] = fmaf(xb[4 * k], g2.x, fmaf(xb[4 * k + 1], g2.y, ssum[2]));
              ssum[3] = fmaf(xb[4 * k + 2], g3.x, fmaf(xb[4 * k + 3], g3.y, ssum[3]));
              const uint32_t o0 = pack_bf16x2((nomask || g0.x != 0.f) ? xa[4 * k] * gsc[0] : 0.f,
                                              (nomask || g0.y != 0.f) ? xa[4 * k + 1] * gsc[0] : 0.f);
              const uint32_t o1 = pack_bf16x2((nomask || g1.x != 0.f) ? xa[4 * k + 2] * gsc[1] : 0.f,
                                              (nomask || g1.y != 0.f) ? xa[4 * k + 3] * gsc[1] : 0.f);
              const uint32_t o2 = pack_bf16x2((nomask || g2.x != 0.f) ? xb[4 * k] * gsc[2] : 0.f,
                                              (nomask || g2.y != 0.f) ? xb[4 * k + 1] * gsc[2] : 0.f);
              const uint32_t o3 = pack_bf16x2((nomask || g3.x != 0.f) ? xb[4 * k + 2] * gsc[3] : 0.f,
                                              (nomask || g3.y != 0.f) ? xb[4 * k + 3] * gsc[3] : 0.f);
              stmatrix_x4_trans(sa, o0, o1, o2, o3);
            }
            continue;
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
#pragma unroll
            for (int e = 0; e < 2; ++e) {
              xa[4 * k + e] = fmaf(xa[4 * k + e], scale[0], bias[0]);
              xa[4 * k + 2 + e] = fmaf(xa[4 * k + 2 + e], scale[1], bias[1]);
              xb[4 * k + e] = fmaf(xb[4 * k + e], scale[2], bias[2]);
              xb[4 * k + 2 + e] = fmaf(xb[4 * k + 2 + e], scale[3], bias[3]);
            }
          }
          act_fwd_vec<16>(xa, p.act);
          act_fwd_vec<16>(xb, p.act);
          if (p.post_scale) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                xa[4 * k + e] *= post[0];
                xa[4 * k + 2 + e] *= post[1];
                xb[4 * k + e] *= post[2];
                xb[4 * k + 2 + e] *= post[3];
              }
            }
          }
          if (p.stat_sums) {
            // fragment element (k, e): pixel row hh0 + 4 j2 + k, pixel column w0 + 2 (lane % 4) + e
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const bool rok = h0 + half * TH + 4 * j2 + k < p.y.h;
#pragma unroll
              for (int e = 0; e < 2; ++e) {
                const bool ok = rok && (e ? col_ok1 : col_ok0);
                const float a0 = ok ? xa[4 * k + e] : 0.f, a1 = ok ? xa[4 * k + 2 + e] : 0.f;
                const float b0 = ok ? xb[4 * k + e] : 0.f, b1 = ok ? xb[4 * k + 2 + e] : 0.f;
                ssum[0] += a0; ssq[0] = fmaf(a0, a0, ssq[0]);
                ssum[1] += a1; ssq[1] = fmaf(a1, a1, ssq[1]);
                ssum[2] += b0; ssq[2] = fmaf(b0, b0, ssq[2]);
                ssum[3] += b1; ssq[3] = fmaf(b1, b1, ssq[3]);
              }
            }
          }
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t sa = st_base + (uint32_t)((j2 * 32 + k * 8) * 128);
            uint32_t o[4] = {pack_bf16x2(xa[4 * k], xa[4 * k + 1]), pack_bf16x2(xa[4 * k + 2], xa[4 * k + 3]),
                             pack_bf16x2(xb[4 * k], xb[4 * k + 1]), pack_bf16x2(xb[4 * k + 2], xb[4 * k + 3])};
            if (has_res) {
              // the staged residual comes back in the very fragment layout that is stored:
              // same warp, same addresses, read before written
              uint32_t rr[4];
              ldmatrix_x4_trans(sa, rr);
#pragma unroll
              for (int e = 0; e < 4; ++e) o[e] = hadd2_bf16(o[e], rr[e]);  // one rounding
            }
            stmatrix_x4_trans(sa, o[0], o[1], o[2], o[3]);
          }
        }
        // the TMA store (async proxy) is issued by a finishing warp after this barrier
        fence_proxy_async();
        asm volatile("bar.arrive %0, 192;" ::"r"(4 + x) : "memory");
        tc_fence_before();
        if (half == nhalf - 1) mbar_arrive(smem_u32(&tmem_empty[buf]));  // accumulator drained
      }
      if (gate_mode && p.dot_sums) {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 1);
          ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 2);
        }
        if ((lane & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            atomicAdd(p.dot_sums + (long long)n * p.cout + o0 + wq * 32 + 8 * j + (lane >> 2),
                      ssum[j] * p.alpha);
        }
      } else if (p.stat_sums) {
        // the 4 lanes of a quad hold the same channels (different pixel columns)
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 1);
          ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], 1);
          ssum[j] += __shfl_xor_sync(0xffffffffu, ssum[j], 2);
          ssq[j] += __shfl_xor_sync(0xffffffffu, ssq[j], 2);
        }
        if ((lane & 3) == 0) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            float* dst = p.stat_sums + ((long long)n * p.cout + o0 + wq * 32 + 8 * j + (lane >> 2)) * 2;
            atomicAdd(dst, ssum[j]);
            atomicAdd(dst + 1, ssq[j]);
          }
        }
      }
    }
  } else if (warp >= 2) {
    // ---------------- finishing warps (2 and 3): TMA store, reflect halo, residual load -------
    // They run one staging tile behind the epilogue warps and off its critical path.  The
    // residual of half-tile hc + 2 is bulk-copied (TMA, same SWIZZLE_128B box as the store) into
    // staging tile x as soon as half-tile hc's store has read it: it has a whole half-tile period
    // to arrive, costs no registers and no LSU traffic, and the epilogue warps add it with
    // ldmatrix / packed-bf16 add / stmatrix in place.
    const int ht = threadIdx.x - 64;  // 0 .. 63
    // (tile, half) of the residual to load next: two half-tiles ahead of the one being finished
    int rt = blockIdx.x, rhalf = 0;
    auto load_residual = [&](int x) {  // one thread
      if (rt >= pp.total_tiles) return;
      int n, h0, w0, o0;
      decode(rt, n, h0, w0, o0);
      const uint32_t bar = smem_u32(&res_full[x]);
      mbar_expect_tx(bar, TILE_BYTES);
#pragma unroll
      for (int sb = 0; sb < BN / 64; ++sb)
        tma_load_4d(smem_u32(stage_base + x * TILE_BYTES + sb * SUB_BYTES), &tmR, bar, o0 + sb * 64, w0,
                    h0 + rhalf * TH, n);
      if (rhalf == 0 && (h0 + TH) < p.y.h) rhalf = 1;
      else { rhalf = 0; rt += gridDim.x; }
    };
    if (has_res && ht == 0) {
      tma_prefetch_desc(&tmR);
      load_residual(0);
      load_residual(1);
    }
    __syncwarp();
    // both staging tiles start out free
    asm volatile("bar.arrive 2, 192;" ::: "memory");
    asm volatile("bar.arrive 3, 192;" ::: "memory");
    int hc = 0;
    for (int t = blockIdx.x; t < pp.total_tiles; t += gridDim.x) {
      int n, h0, w0, o0;
      decode(t, n, h0, w0, o0);
      const int nhalf = (h0 + TH) < p.y.h ? 2 : 1;
#pragma unroll 1
      for (int half = 0; half < nhalf; ++half, ++hc) {
        const int x = hc & 1;
        const int hh0 = h0 + half * TH;
        uint8_t* const stage_out = stage_base + x * TILE_BYTES;
        asm volatile("bar.sync %0, 192;" ::"r"(4 + x) : "memory");  // half-tile staged
        if (ht == 0) {
#pragma unroll
          for (int sb = 0; sb < BN / 64; ++sb)
            tma_store_4d(&tmY, smem_u32(stage_out + sb * SUB_BYTES), o0 + sb * 64, w0, hh0, n);
          tma_store_commit();
        }
        if (p.y_halo > 0) {
          // reflect border: staged pixel rows ht and ht + 64 (second stores of interior pixels)
#pragma unroll 1
          for (int r = 0; r < 2; ++r) {
            const int m = ht + 64 * r;
            const int oh = hh0 + m / TW, ow = w0 + m % TW;
            if (oh < p.y.h && ow < p.y.w) {
              int hs[3], ws[3];
              const int nh = mirror_set(oh, p.y.h, p.y_halo, hs);
              const int nw = mirror_set(ow, p.y.w, p.y_halo, ws);
              if (nh * nw > 1) {
                const uint8_t* myrow = stage_out + m * 128;
                const int sw = m & 7;
                for (int a = 0; a < nh; ++a)
                  for (int b = 0; b < nw; ++b)
                    if (a + b > 0) {
                      uint4* dst = reinterpret_cast<uint4*>(
                          vptr_mut<__nv_bfloat16>(p.y, n, hs[a], ws[b], o0));
#pragma unroll 4
                      for (int pc = 0; pc < BN / 8; ++pc)
                        dst[pc] = *reinterpret_cast<const uint4*>(
                            myrow + (pc >> 3) * SUB_BYTES + (((pc & 7) ^ sw) << 4));
                    }
              }
            }
          }
        }
        // both warps are done reading the staged rows; once the TMA store has read the tile too
        // it takes the residual of half-tile hc + 2 and is handed back to the epilogue warps
        asm volatile("bar.sync 1, 64;" ::: "memory");
        if (ht == 0) {
          tma_store_wait_read();
          if (has_res) load_residual(x);
        }
        __syncwarp();
        asm volatile("bar.arrive %0, 192;" ::"r"(2 + x) : "memory");
      }
    }
    if (ht == 0) tma_store_wait_all();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// wgrad kernel
// ---------------------------------------------------------------------------
struct TcWgP {
  float* dw;
  float* ws;  // [tap][M][N] fp32 partial-sum workspace (NULL: scalar reds into dw)
  const __nv_bfloat16* wfwd;  // forward pack [n | 1][cout][taps][cin] (fused P term)
  float* P;                   // [n][cout]
  const float* rs;
  const float* cs;
  float alpha;
  int cin, cout, kh, kw;
  int a_is_x;        // 1: M side = x (Cin), N side = dy (Cout); 0: M = dy, N = x
  int x_coord_off;   // x_halo - pad
  int tiles_w, tiles_total, splits;
  int n_tiles_n;     // number of N tiles (grid.y = m_tiles * n_tiles_n)
  long long wfwd_n_stride;  // elements between per-sample forward packs (0: one shared pack)
};

constexpr int WG_CHUNK_BYTES = 64 * 128;  // 64 pixels x 64 channels bf16

// TPC = filter taps per CTA.  The dy tile (the M operand when Cout % 128 == 0) is the same for
// every tap, only the x tile shifts: TPC x tiles are laid side by side in the ring as ONE
// MN-major B operand of N = TPC * BN columns (the 64-channel chunk stride is uniform), so one
// 128 x (TPC*BN) x 16 MMA serves TPC taps and reads the dy tile once.  A 128 x 128 MMA needs
// 128 B/clk of operands (= the SM's shared-memory bandwidth, ~1.0-1.1 PFLOP/s measured with the
// reduction skipped); at N = 256 it is 96 B/clk, the same ratio that lifted the forward kernel.
template <int BN, int STAGES, int OCC, int TPC>
__global__ void __launch_bounds__(256, OCC)
conv_tc_wgrad_kernel(const __grid_constant__ CUtensorMap tmA,
                     const __grid_constant__ CUtensorMap tmB, TcWgP p) {
  static_assert(BN * TPC <= 256, "accumulator exceeds one MMA's N / the TMEM share of a CTA");
  constexpr int NCOLS = BN * TPC;
  constexpr int A_BYTES = 2 * WG_CHUNK_BYTES;
  constexpr int B_TAP_BYTES = (BN / 64) * WG_CHUNK_BYTES;
  constexpr int STAGE_BYTES = A_BYTES + TPC * B_TAP_BYTES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint64_t* bars = (uint64_t*)(smem + STAGES * STAGE_BYTES);
  uint64_t* full = bars;
  uint64_t* empty = bars + STAGES;
  uint64_t* accum_full = bars + 2 * STAGES;
  uint32_t* tmem_ptr = (uint32_t*)(bars + 2 * STAGES + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  const int taps = p.kh * p.kw;
  const int tap0 = blockIdx.x * TPC;
  const int ntap = min(TPC, taps - tap0);  // the last CTA of an odd tap count takes fewer
  const int mt = blockIdx.y / p.n_tiles_n, nt = blockIdx.y % p.n_tiles_n;
  const int m0 = mt * 128, n0 = nt * BN;
  const int n = blockIdx.z / p.splits, split = blockIdx.z % p.splits;
  const int t_beg = (int)((long long)p.tiles_total * split / p.splits);
  const int t_end = (int)((long long)p.tiles_total * (split + 1) / p.splits);
  const int num_k = t_end - t_beg;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < STAGES; ++st) {
      mbar_init(smem_u32(&full[st]), 1);
      mbar_init(smem_u32(&empty[st]), 1);
    }
    mbar_init(smem_u32(accum_full), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<NCOLS>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (num_k > 0) {
    if (warp == 0) {
      // TPC > 1 only with a_is_x == 0 (host): the A operand (dy) does not depend on the tap
      const int r0 = tap0 / p.kw, s0 = tap0 - r0 * p.kw;
      const int sh_a = p.a_is_x ? (r0 + p.x_coord_off) : 0, sw_a = p.a_is_x ? (s0 + p.x_coord_off) : 0;
      const uint32_t stage_tx = (uint32_t)(A_BYTES + ntap * B_TAP_BYTES);
      for (int it = 0; it < num_k; ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = ((it / STAGES) & 1) ^ 1;
        mbar_wait(smem_u32(&empty[stage]), parity);
        if (lane == 0) {
          const int t = t_beg + it;
          const int h0 = (t / p.tiles_w) * 8, w0 = (t % p.tiles_w) * 8;
          const uint32_t bar = smem_u32(&full[stage]);
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          mbar_expect_tx(bar, stage_tx);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_4d(sa + c * WG_CHUNK_BYTES, &tmA, bar, m0 + c * 64, w0 + sw_a, h0 + sh_a, n);
          for (int tp = 0; tp < ntap; ++tp) {
            const int tap = tap0 + tp;
            const int r = tap / p.kw, s = tap - r * p.kw;
            const int sh_b = p.a_is_x ? 0 : (r + p.x_coord_off), sw_b = p.a_is_x ? 0 : (s + p.x_coord_off);
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_4d(sa + A_BYTES + tp * B_TAP_BYTES + c * WG_CHUNK_BYTES, &tmB, bar, n0 + c * 64,
                          w0 + sw_b, h0 + sh_b, n);
          }
        }
        __syncwarp();
      }
    } else if (warp == 1) {
      const uint32_t idesc = make_idesc(128, BN * ntap, 1, 1);
      for (int it = 0; it < num_k; ++it) {
        const int stage = it % STAGES;
        const uint32_t parity = (it / STAGES) & 1;
        mbar_wait(smem_u32(&full[stage]), parity);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * STAGE_BYTES);
          // MN-major SW128: LBO = stride between 64-channel chunks, SBO = 8 pixel rows (1024 B)
          const uint64_t da = make_desc(sa, WG_CHUNK_BYTES, 1024);
          const uint64_t db = make_desc(sa + A_BYTES, WG_CHUNK_BYTES, 1024);
#pragma unroll
          for (int k = 0; k < 4; ++k)  // 16 pixels (K) = 2048 B per step
            umma_bf16(tmem_base, da + (uint64_t)(k * 128), db + (uint64_t)(k * 128), idesc,
                      (it > 0 || k > 0) ? 1u : 0u);
          umma_commit(smem_u32(&empty[stage]));
          if (it == num_k - 1) umma_commit(smem_u32(accum_full));
        }
        __syncwarp();
      }
    } else if (warp >= 4) {
      const int wq = warp - 4;
      mbar_wait(smem_u32(accum_full), 0);
      tc_fence_after();
      const int m = m0 + wq * 32 + lane;
      // row factor
      float rowf = p.alpha;
      if (p.a_is_x) { if (p.cs) rowf *= p.cs[(long long)n * p.cin + m]; }
      else { if (p.rs) rowf *= p.rs[(long long)n * p.cout + m]; }
      const int Mtot = p.a_is_x ? p.cin : p.cout, Ntot = p.a_is_x ? p.cout : p.cin;
      const float* colv = p.a_is_x ? p.rs : p.cs;  // per-sample factor along the N (column) axis
      float pacc = 0.f;  // sum_i G[o,i,tap] * wfwd[n][o][tap][i] over this tile's columns
#pragma unroll 1
      for (int tp = 0; tp < ntap; ++tp) {
        const int tap = tap0 + tp;
        const __nv_bfloat16* wrow =
            p.P ? p.wfwd + (long long)n * p.wfwd_n_stride + ((long long)m * taps + tap) * p.cin : nullptr;
#pragma unroll 1
        for (int j = 0; j < BN / 16; ++j) {
          float v[16];
          tmem_ld16(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(tp * BN + j * 16), v);
          const int nn0 = n0 + j * 16;
          if (wrow) {
            float w0[8], w1[8];
            load_vec<__nv_bfloat16, 8>(wrow + nn0, w0);
            load_vec<__nv_bfloat16, 8>(wrow + nn0 + 8, w1);
#pragma unroll
            for (int i = 0; i < 8; ++i) pacc = fmaf(v[i], w0[i], fmaf(v[8 + i], w1[i], pacc));
          }
          if (colv) {
            const float4* cp = reinterpret_cast<const float4*>(colv + (long long)n * Ntot + nn0);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              float4 t = cp[q];
              v[4 * q] *= t.x; v[4 * q + 1] *= t.y; v[4 * q + 2] *= t.z; v[4 * q + 3] *= t.w;
            }
          }
#pragma unroll
          for (int i = 0; i < 16; ++i) v[i] *= rowf;
          if (p.ws) {  // 128-bit vector reds, contiguous along N
            float4* dst = reinterpret_cast<float4*>(p.ws + ((long long)tap * Mtot + m) * Ntot + nn0);
#pragma unroll
            for (int q = 0; q < 4; ++q)
              atomicAdd(dst + q, make_float4(v[4 * q], v[4 * q + 1], v[4 * q + 2], v[4 * q + 3]));
          } else {
#pragma unroll
            for (int i = 0; i < 16; ++i) {
              const int nn = nn0 + i;
              const int o = p.a_is_x ? nn : m, ci = p.a_is_x ? m : nn;
              atomicAdd(p.dw + ((long long)o * p.cin + ci) * taps + tap, v[i]);
            }
          }
        }
      }
      if (p.P) {
        const float rsv = p.rs ? p.rs[(long long)n * p.cout + m] : 1.f;
        atomicAdd(p.P + (long long)n * p.cout + m, pacc * rsv);
      }
      tc_fence_before();
    }
  }
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<NCOLS>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// wgrad, filter-column form (3x3, Cout % 128 == 0, Cin % 128 == 0): the heavy weight gradients
// ---------------------------------------------------------------------------
// The one-tap-per-CTA kernel above uses every operand byte it stages exactly once: per K = 64
// pixels it WRITES 32 KB into shared memory (TMA) and READS 32 KB (four 128x128x16 MMAs), 256
// B/clk against the SM's 128 B/clk -- ncu shows the tensor pipe at 51 %.  Here a CTA owns one
// FILTER COLUMN s of a (128 cout x 128 cin) block, i.e. the three taps (r, s), r = 0..2:
//   * a K-group of 8 pixels is one image-row segment = one 1024-byte swizzle atom, the 8 rows of
//     a pixel tile are consecutive atoms (the layout the one-tap kernel already uses);
//   * the x box holds 10 rows x 8 columns; tap r reads pixel rows r .. r+7, i.e. the SAME bytes
//     at an atom-aligned offset of r KB.  With LBO = SBO = 1024 the three taps are the three
//     64-channel "chunks" of ONE MN-major B operand: a 128 x 192 x 16 MMA per 64 input channels
//     computes all three taps and the dy tile is read once for them;
//   * accumulators: 2 channel chunks x 192 columns in TMEM.
// Per 64 pixels: 36 KB staged, 80 KB read by eight 96-cycle MMAs = 151 B/clk (was 256).
// Persistent, stream-K: the (sample, block, filter column) items x pixel tiles are one sequence
// cut into equal ranges, one per CTA; a CTA flushes its accumulator (per-sample row / column
// factors, fused P term, bulk reduce-add into the fp32 workspace) whenever the item changes, so
// any batch size balances.  Measured (n = 96, 128 -> 128 @ 64x64): 116 us = 1.0 PFLOP/s; the
// main loop alone 1.25-1.3.  A double-buffered 192-column variant (64 input channels per CTA, the
// flush hidden) stages 45 % more bytes per MMA and ran at 152 us: the staging traffic, not the
// flush, is what binds.
struct TcWgRowP {
  float* ws;                  // [tap][cout][cin] fp32
  const __nv_bfloat16* wfwd;  // forward pack [n | 1][cout][taps][cin] (fused P term) or NULL
  float* P;                   // [n][cout]
  float* Q;                   // [n][cin]: sum_{o,tap} rs * G_n[o,i,tap] * wfwd[o][tap][i], or NULL
  const float* rs;            // [n][cout] or NULL
  const float* cs;            // [n][cin] or NULL
  float alpha;
  int cin, cout;
  int x_coord_off;            // x_halo - pad
  int tiles_w, tiles_total;   // 8x8 pixel tiles of the dy image
  int m_tiles, n_tiles;       // cout / 128, cin / 128
  long long units_total;      // items * tiles_total
  long long wfwd_n_stride;
};

constexpr int WGR_STAGES = 4;
constexpr int WGR_A_BYTES = 2 * 8 * 1024;        // dy: 2 chunks x 8 pixel rows x 1 KB
constexpr int WGR_B_CHUNK = 10 * 1024;           // x: 10 pixel rows x 1 KB per 64 channels
constexpr int WGR_STAGE_BYTES = WGR_A_BYTES + 2 * WGR_B_CHUNK;
constexpr int WGR_RED_PITCH = 256 + 16;          // one thread's 64 fp32 partial sums (+ bank skew)
constexpr int WGR_RED_BYTES = 256 * WGR_RED_PITCH;  // one row per epilogue thread

__global__ void __launch_bounds__(384, 1)
conv_tc_wgrad_row_kernel(const __grid_constant__ CUtensorMap tmDy,
                         const __grid_constant__ CUtensorMap tmX, TcWgRowP p) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* stage_red = smem + WGR_STAGES * WGR_STAGE_BYTES;  // staging rows + cs_s[128]
  uint64_t* bars = (uint64_t*)(stage_red + WGR_RED_BYTES + 512);
  uint64_t* full = bars;
  uint64_t* empty = bars + WGR_STAGES;
  uint64_t* accum_full = bars + 2 * WGR_STAGES;
  uint64_t* accum_empty = accum_full + 1;
  uint32_t* tmem_ptr = (uint32_t*)(accum_empty + 1);

  const int warp = threadIdx.x / 32, lane = threadIdx.x % 32;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmDy);
    tma_prefetch_desc(&tmX);
  }
  if (warp == 1 && lane == 0) {
    for (int st = 0; st < WGR_STAGES; ++st) {
      mbar_init(smem_u32(&full[st]), 1);
      mbar_init(smem_u32(&empty[st]), 1);
    }
    mbar_init(smem_u32(accum_full), 1);
    mbar_init(smem_u32(accum_empty), 256);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(smem_u32(tmem_ptr));
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  // this CTA's range of (item, tile) units; item = ((n * m_tiles + mt) * n_tiles + nt) * 3 + fs
  const long long u_beg = p.units_total * blockIdx.x / gridDim.x;
  const long long u_end = p.units_total * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ---------------- TMA producer ----------------
    int it = 0;
    long long u = u_beg;
    while (u < u_end) {
      const long long item = u / p.tiles_total;
      const int t0 = (int)(u - item * p.tiles_total);
      const int t1 = (int)min((long long)p.tiles_total, t0 + (u_end - u));
      const int fs = (int)(item % 3);  // filter column
      long long rest = item / 3;
      const int nt = (int)(rest % p.n_tiles); rest /= p.n_tiles;
      const int mt = (int)(rest % p.m_tiles);
      const int n = (int)(rest / p.m_tiles);
      for (int t = t0; t < t1; ++t, ++it) {
        const int stage = it % WGR_STAGES;
        mbar_wait(smem_u32(&empty[stage]), ((it / WGR_STAGES) & 1) ^ 1);
        if (lane == 0) {
          const int h0 = (t / p.tiles_w) * 8, w0 = (t % p.tiles_w) * 8;
          const uint32_t bar = smem_u32(&full[stage]);
          const uint32_t sa = smem_u32(smem + stage * WGR_STAGE_BYTES);
          mbar_expect_tx(bar, WGR_STAGE_BYTES);
#pragma unroll
          for (int c = 0; c < 2; ++c)
            tma_load_4d(sa + c * (WGR_A_BYTES / 2), &tmDy, bar, mt * 128 + c * 64, w0, h0, n);
#pragma unroll
          for (int c = 0; c < 2; ++c)  // 8 pixel columns from w0 + fs, 10 pixel rows from h0
            tma_load_4d(sa + WGR_A_BYTES + c * WGR_B_CHUNK, &tmX, bar, nt * 128 + c * 64,
                        w0 + fs + p.x_coord_off, h0 + p.x_coord_off, n);
        }
        __syncwarp();
      }
      u += t1 - t0;
    }
  } else if (warp == 1) {
    // ---------------- MMA issuer ----------------
    constexpr uint32_t idesc = make_idesc(128, 192, 1, 1);
    int it = 0, seg = 0;
    long long u = u_beg;
    while (u < u_end) {
      const long long item = u / p.tiles_total;
      const int t0 = (int)(u - item * p.tiles_total);
      const int t1 = (int)min((long long)p.tiles_total, t0 + (u_end - u));
      mbar_wait(smem_u32(accum_empty), (seg & 1) ^ 1);  // previous segment flushed
      tc_fence_after();
      for (int t = t0; t < t1; ++t, ++it) {
        const int stage = it % WGR_STAGES;
        mbar_wait(smem_u32(&full[stage]), (it / WGR_STAGES) & 1);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t sa = smem_u32(smem + stage * WGR_STAGE_BYTES);
          // MN-major SW128.  A: LBO = next 64 couts (8 KB), SBO = next 8 pixels (1 KB).
          // B: LBO = next TAP (the same pixels one row down = 1 KB), SBO = next 8 pixels (1 KB).
          const uint64_t da = make_desc(sa, WGR_A_BYTES / 2, 1024);
#pragma unroll
          for (int c = 0; c < 2; ++c) {
            const uint64_t db = make_desc(sa + WGR_A_BYTES + c * WGR_B_CHUNK, 1024, 1024);
#pragma unroll
            for (int k = 0; k < 4; ++k)  // 16 pixels (K) = 2 pixel rows = 2048 B per step
              umma_bf16(tmem_base + (uint32_t)(c * 192), da + (uint64_t)(k * 128),
                        db + (uint64_t)(k * 128), idesc, (t > t0 || k > 0) ? 1u : 0u);
          }
          umma_commit(smem_u32(&empty[stage]));
          if (t == t1 - 1) umma_commit(smem_u32(accum_full));
        }
        __syncwarp();
      }
      u += t1 - t0;
      ++seg;
    }
  } else if (warp >= 4) {
    // ---------------- epilogue: flush one segment (8 warps) ----------------
    // Warps 4-7 drain channel chunk 0 (TMEM columns 0..191), warps 8-11 chunk 1; thread =
    // output channel (TMEM lane).  Per filter row the 64 scaled partial sums of a thread go to
    // ITS OWN 256-byte row of a staging tile, one 128-byte half at a time, and from there to the
    // fp32 workspace with bulk reduce-adds (cp.reduce.async.bulk .add.f32: the L2 adds whole
    // sectors) instead of scattered 16-byte reds through the LSU; the two halves alternate, so a
    // half is rewritten only after its reduce has read it.  The per-sample column factors come
    // from shared memory and the forward weights of the fused P term are fetched one filter row
    // ahead: no global-load latency sits between the TMEM reads.
    const int te = threadIdx.x - 128;   // 0 .. 255
    const int c = te >> 7;              // channel chunk of this warpgroup
    const int lane_m = te & 127;        // TMEM lane
    const int wq = warp & 3;
    float* cs_s = reinterpret_cast<float*>(stage_red + WGR_RED_BYTES);  // [128]
    uint8_t* row = stage_red + te * WGR_RED_PITCH;
    int seg = 0;
    long long u = u_beg;
    while (u < u_end) {
      const long long item = u / p.tiles_total;
      const int t0 = (int)(u - item * p.tiles_total);
      const int t1 = (int)min((long long)p.tiles_total, t0 + (u_end - u));
      const int fs = (int)(item % 3);
      long long rest = item / 3;
      const int nt = (int)(rest % p.n_tiles); rest /= p.n_tiles;
      const int mt = (int)(rest % p.m_tiles);
      const int n = (int)(rest / p.m_tiles);
      const int m = mt * 128 + lane_m;  // cout
      const float rsv = p.rs ? p.rs[(long long)n * p.cout + m] : 1.f;
      const float rowf = p.alpha * rsv;
      // the previous segment's flush has read cs_s (barrier at its end)
      if (te < 128) cs_s[te] = p.cs ? p.cs[(long long)n * p.cin + nt * 128 + te] : 1.f;
      const __nv_bfloat16* wbase =
          (p.P || p.Q)
              ? p.wfwd + (long long)n * p.wfwd_n_stride + (long long)m * 9 * p.cin + nt * 128 + c * 64
              : nullptr;
      float qacc[2] = {0.f, 0.f};  // this lane's column (hf * 32 + lane) of the Q term
      uint4 wnext[8];
      if (wbase) {  // filter row 0: tap fs
#pragma unroll
        for (int i = 0; i < 8; ++i)
          wnext[i] = *reinterpret_cast<const uint4*>(wbase + (long long)fs * p.cin + i * 8);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // cs_s complete
      mbar_wait(smem_u32(accum_full), seg & 1);
      tc_fence_after();
      float pacc = 0.f;  // sum_i G[o,i,tap] * wfwd[n][o][tap][i] over this thread's columns
#pragma unroll 1
      for (int fr = 0; fr < 3; ++fr) {  // filter row of this accumulator group
        const int tap = fr * 3 + fs;
        uint4 wcur[8];
        if (wbase) {
#pragma unroll
          for (int i = 0; i < 8; ++i) wcur[i] = wnext[i];
          if (fr < 2) {
#pragma unroll
            for (int i = 0; i < 8; ++i)
              wnext[i] = *reinterpret_cast<const uint4*>(wbase + (long long)((fr + 1) * 3 + fs) * p.cin + i * 8);
          }
        }
        float* dst = p.ws + ((long long)tap * p.cout + m) * p.cin + nt * 128 + c * 64;
#pragma unroll
        for (int hf = 0; hf < 2; ++hf) {  // 32 columns = one 128-byte half row
          float v[32];
          tmem_ld32(tmem_base + ((uint32_t)(wq * 32) << 16) + (uint32_t)(c * 192 + fr * 64 + hf * 32), v);
          if (p.P) {
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat162* wp2 = reinterpret_cast<const __nv_bfloat162*>(&wcur[4 * hf + k]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 wf = __bfloat1622float2(wp2[i]);
                pacc = fmaf(v[8 * k + 2 * i], wf.x, fmaf(v[8 * k + 2 * i + 1], wf.y, pacc));
              }
            }
          }
          if (p.Q) {
            // column sums over the 32 output channels of this warp: transpose-reduce, after
            // step `off` a lane keeps the half of the columns its bit `off` selects, so lane l
            // ends with column l (31 shuffles for 32 columns)
            float pr[32];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
              const __nv_bfloat162* wp2 = reinterpret_cast<const __nv_bfloat162*>(&wcur[4 * hf + k]);
#pragma unroll
              for (int i = 0; i < 4; ++i) {
                const float2 wf = __bfloat1622float2(wp2[i]);
                pr[8 * k + 2 * i] = v[8 * k + 2 * i] * wf.x;
                pr[8 * k + 2 * i + 1] = v[8 * k + 2 * i + 1] * wf.y;
              }
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) {
              const bool up = (lane & off) != 0;
#pragma unroll
              for (int i = 0; i < off; ++i) {
                const float send = up ? pr[i] : pr[i + off];
                const float keep = up ? pr[i + off] : pr[i];
                pr[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
              }
            }
            qacc[hf] += pr[0];
          }
          // the bulk reduce issued from this half row one filter row ago has read it
          asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");
          const float4* cp = reinterpret_cast<const float4*>(cs_s + c * 64 + hf * 32);
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const float4 tq = cp[q];
            reinterpret_cast<float4*>(row + hf * 128)[q] =
                make_float4(v[4 * q] * tq.x * rowf, v[4 * q + 1] * tq.y * rowf,
                            v[4 * q + 2] * tq.z * rowf, v[4 * q + 3] * tq.w * rowf);
          }
          fence_proxy_async();  // this thread's half row: generic writes -> read by the bulk reduce
          asm volatile(
              "cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], 128;" ::"l"(
                  dst + hf * 32),
              "r"(smem_u32(row + hf * 128))
              : "memory");
          asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        }
      }
      tc_fence_before();
      mbar_arrive(smem_u32(accum_empty));  // accumulator drained: the next segment may start
      if (p.P) atomicAdd(p.P + (long long)n * p.cout + m, pacc * rsv);
      if (p.Q) {  // (offered only without rs: a row factor would have to scale the products)
        float* qd = p.Q + (long long)n * p.cin + nt * 128 + c * 64 + lane;
        atomicAdd(qd, qacc[0]);
        atomicAdd(qd + 32, qacc[1]);
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");  // everybody has read cs_s
      u += t1 - t0;
      ++seg;
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                  const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                  const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &q) ==
            cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// 4-D map over an NHWC bf16 view including `halo` pixels on every side
static int make_act_map(CUtensorMap* map, const otm_tensor& t, int halo, int box_w, int box_h) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(OTM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  char* base = (char*)t.ptr - (long long)halo * (t.sh + t.sw) * 2;
  cuuint64_t dims[4] = {(cuuint64_t)t.c, (cuuint64_t)(t.w + 2 * halo), (cuuint64_t)(t.h + 2 * halo),
                        (cuuint64_t)t.n};
  cuuint64_t strides[3] = {(cuuint64_t)t.sw * 2, (cuuint64_t)t.sh * 2, (cuuint64_t)t.sn * 2};
  cuuint32_t box[4] = {64, (cuuint32_t)box_w, (cuuint32_t)box_h, 1};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, base, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS)
    return fail(OTM_ERR_CUDA, "cuTensorMapEncodeTiled(act) failed: %d (c=%d w=%d h=%d n=%d)", (int)r,
                t.c, t.w, t.h, t.n);
  return OTM_OK;
}

static int make_weight_map(CUtensorMap* map, const void* w, long long rows, long long ktot,
                           int box_rows) {
  EncodeTiledFn enc = get_encode();
  if (!enc) return fail(OTM_ERR_CUDA, "cuTensorMapEncodeTiled entry point not available");
  cuuint64_t dims[2] = {(cuuint64_t)ktot, (cuuint64_t)rows};
  cuuint64_t strides[1] = {(cuuint64_t)ktot * 2};
  cuuint32_t box[2] = {64, (cuuint32_t)box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(OTM_ERR_CUDA, "cuTensorMapEncodeTiled(weights) failed: %d", (int)r);
  return OTM_OK;
}

static bool act_tma_ok(const otm_tensor& t) {
  return t.dtype == OTM_BF16 && t.c % 64 == 0 && t.sw % 8 == 0 && t.sh % 8 == 0 && t.sn % 8 == 0 &&
         ((uintptr_t)t.ptr % 16 == 0);
}

bool conv_fwd_tc_eligible(const otm_conv_fwd_args* a) {
  if (a->kh != a->kw || (a->kh != 3 && a->kh != 4)) return false;  // the row-reuse kernels
  if (!act_tma_ok(a->x) || a->y.dtype != OTM_BF16) return false;
  if (a->y.c % 64 != 0 || a->y.sw % 8 || a->y.sh % 8 || a->y.sn % 8) return false;
  if ((uintptr_t)a->y.ptr % 16 || (uintptr_t)a->wpack % 16) return false;
  if (a->w_batch_stride != 0 && a->w_batch_stride != (long long)a->y.c * a->kh * a->kw * a->x.c)
    return false;
  if (a->residual.ptr && (a->residual.dtype != OTM_BF16 || a->residual.sw % 8 ||
                          a->residual.sh % 8 || a->residual.sn % 8 ||
                          (uintptr_t)a->residual.ptr % 16))
    return false;
  return true;
}

template <int BN, int KS, int NA, int NB>
static int launch_fwd_rr(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                         const TcFwdPP& pp, int ctas, cudaStream_t st) {
  constexpr int smem = NA * (16 + KS - 1) * 8 * 128 + NB * BN * 128 + (BN / 64) * 128 * 128 + 1024 +
                       512 + 3 * BN * 4;
  static_assert(smem <= 227 * 1024, "row-reuse conv kernel exceeds shared memory");
  auto kern = conv_tc_fwd_rr_kernel<BN, KS, NA, NB>;
  OTM_ENSURE_SMEM(kern, smem);
  kern<<<ctas, 256, smem, st>>>(tmA, tmB, tmY, pp);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

template <int BN, int KS, int NA, int NB>
static int launch_fwd_rr2(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                          const TcFwdPP& pp, int ctas, cudaStream_t st) {
  constexpr int smem = NA * 2 * (16 + KS - 1) * 8 * 128 + NB * BN * 128 + (BN / 64) * 128 * 128 +
                       1024 + 512 + 3 * BN * 4;
  static_assert(smem <= 227 * 1024, "pair conv kernel exceeds shared memory");
  static_assert(4 * BN <= 512, "pair conv kernel exceeds TMEM");
  auto kern = conv_tc_fwd_rr2_kernel<BN, KS, NA, NB>;
  OTM_ENSURE_SMEM(kern, smem);
  kern<<<ctas, 256, smem, st>>>(tmA, tmB, tmY, pp);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

template <int KS, int NA, int NB>
static int launch_fwd_rr2t(const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmY,
                           const CUtensorMap& tmR, const TcFwdPP& pp, int ctas, cudaStream_t st) {
  constexpr int smem = NA * (32 + KS - 1) * 8 * 128 + NB * 128 * 128 + 2 * (2 * 128 * 128) + 1024 + 512;
  static_assert(smem <= 227 * 1024, "transposed pair conv kernel exceeds shared memory");
  auto kern = conv_tc_fwd_rr2t_kernel<KS, NA, NB>;
  OTM_ENSURE_SMEM(kern, smem);
  kern<<<ctas, 256, smem, st>>>(tmA, tmB, tmY, tmR, pp);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

// would conv_fwd_tc run these arguments on the transposed pair kernel? (the one whose epilogue
// can accumulate the consumer's InstanceNorm statistics); mirrors the choice made below
bool conv_fwd_tc_uses_rr2t(const otm_conv_fwd_args* a) {
  const int cout = a->y.c;
  if (cout % 128 != 0) return false;
  const int tile_rows = (a->y.h + 15) / 16, tiles_w = (a->y.w + 7) / 8;
  return (long long)((tile_rows + 1) / 2) * tiles_w * (cout / 128) * a->y.n >= 2LL * num_sms();
}

int conv_fwd_tc(const otm_conv_fwd_args* a, cudaStream_t st) {
  const int Ho = a->y.h, Wo = a->y.w, cout = a->y.c, cin = a->x.c;
  const int KS = a->kh;  // 3 or 4 (conv_fwd_tc_eligible)
  // row-reuse kernels: fixed 16 x 8 pixel tile
  constexpr int TW = 8, TH = 16;
  const int tile_rows = (Ho + TH - 1) / TH, tiles_w = (Wo + TW - 1) / TW;
  const int tiles = tile_rows * tiles_w;
  // 128-channel tiles go to the transposed pair kernel whenever there are enough tile pairs to
  // fill the SMs twice -- also for 256 / 512 output channels (two / four weight tiles per pixel
  // pair: measured 1465-1580 vs 1260-1350 TFLOP/s for the 256-wide tile at 256 -> 256 @64x64);
  // the 256-wide tile is kept for the small layers
  int BN = 128;
  if (cout % 128 != 0) BN = 64;
  else if (cout % 256 == 0 &&
           (long long)((tile_rows + 1) / 2) * tiles_w * (cout / 128) * a->y.n < 2LL * num_sms() &&
           (long long)tiles * a->y.n * (cout / 256) >= 2 * num_sms())
    BN = 256;

  CUtensorMap tmA, tmB, tmY;
  int rc = make_act_map(&tmA, a->x, a->x_halo, TW, TH + KS - 1);
  if (rc) return rc;
  const long long ktot = (long long)KS * KS * cin;
  const long long rows = (a->w_batch_stride ? (long long)a->x.n : 1) * cout;
  rc = make_weight_map(&tmB, a->wpack, rows, ktot, BN);
  if (rc) return rc;
  rc = make_act_map(&tmY, a->y, 0, TW, TH);  // interior of y: the TMA store clips tile tails
  if (rc) return rc;

  TcFwdPP pp;
  TcFwdP& p = pp.p;
  p.y = make_view(a->y);
  p.res = a->residual.ptr ? make_view(a->residual) : null_view();
  p.row_scale = a->row_scale; p.post_scale = a->post_scale; p.bias = a->bias; p.alpha = a->alpha;
  p.stat_sums = nullptr;
  p.dot_sums = nullptr;
  p.res_mode = 0;
  p.act = a->act;
  p.y_halo = a->y_halo; p.cin = cin; p.cout = cout; p.kh = KS; p.kw = KS;
  p.coord_off = a->x_halo - a->pad;
  p.TW = TW; p.TH = TH; p.tiles_w = tiles_w;
  p.w_rows_per_sample = a->w_batch_stride ? cout : 0;
  pp.cout_tiles = cout / BN;
  pp.tiles_per_img = tiles;
  pp.total_tiles = tiles * pp.cout_tiles * a->y.n;
  int ctas = num_sms();
  // pair kernels: two vertically adjacent tiles per CTA step when there are enough pairs to fill
  // the SMs twice
  const long long pairs = (long long)((tile_rows + 1) / 2) * tiles_w * pp.cout_tiles * a->y.n;
  if (BN <= 128 && pairs >= 2LL * num_sms()) {
    pp.tiles_per_img = ((tile_rows + 1) / 2) * tiles_w;
    pp.total_tiles = (int)pairs;
    if (BN == 128) {  // weights as the M operand, ONE 256-pixel box for both tiles of the pair
      CUtensorMap tmA2;
      rc = make_act_map(&tmA2, a->x, a->x_halo, TW, 2 * TH + KS - 1);
      if (rc) return rc;
      if (a->stat_sums) {
        p.stat_sums = a->stat_sums;
        OTM_CHECK_CUDA(cudaMemsetAsync(a->stat_sums, 0, sizeof(float) * 2 * (size_t)a->y.n * cout, st));
      }
      if (a->residual_mode != 0 && a->residual.ptr) {
        p.res_mode = a->residual_mode;
        p.dot_sums = a->dot_sums;
        if (a->dot_sums)
          OTM_CHECK_CUDA(cudaMemsetAsync(a->dot_sums, 0, sizeof(float) * (size_t)a->y.n * cout, st));
      }
      CUtensorMap tmR = tmY;  // residual: same boxes as the output tile, its own strides
      if (a->residual.ptr) {
        rc = make_act_map(&tmR, a->residual, 0, TW, TH);
        if (rc) return rc;
      }
      if (KS == 3) return launch_fwd_rr2t<3, 2, 4>(tmA2, tmB, tmY, tmR, pp, ctas, st);
      return launch_fwd_rr2t<4, 2, 4>(tmA2, tmB, tmY, tmR, pp, ctas, st);
    }
    if (KS == 3) return launch_fwd_rr2<64, 3, 3, 8>(tmA, tmB, tmY, pp, ctas, st);
    return launch_fwd_rr2<64, 4, 3, 8>(tmA, tmB, tmY, pp, ctas, st);
  }
  if (ctas > pp.total_tiles) ctas = pp.total_tiles;
  if (KS == 3) {
    if (BN == 64) return launch_fwd_rr<64, 3, 4, 12>(tmA, tmB, tmY, pp, ctas, st);
    if (BN == 128) return launch_fwd_rr<128, 3, 3, 8>(tmA, tmB, tmY, pp, ctas, st);
    return launch_fwd_rr<256, 3, 3, 3>(tmA, tmB, tmY, pp, ctas, st);
  }
  if (BN == 64) return launch_fwd_rr<64, 4, 4, 12>(tmA, tmB, tmY, pp, ctas, st);
  if (BN == 128) return launch_fwd_rr<128, 4, 3, 8>(tmA, tmB, tmY, pp, ctas, st);
  return launch_fwd_rr<256, 4, 3, 3>(tmA, tmB, tmY, pp, ctas, st);
}

bool conv_wgrad_tc_eligible(const otm_conv_wgrad_args* a) {
  if (!act_tma_ok(a->x) || !act_tma_ok(a->dy)) return false;
  const int cin = a->x.c, cout = a->dy.c;
  if (cout % 128 != 0 && cin % 128 != 0) return false;
  return true;
}

// dw[o][c][tap] += ws[tap][m][n]  (m,n = (o,c) or (c,o) in the GEMM's orientation)
__global__ void __launch_bounds__(256) wgrad_fold_kernel(const float* __restrict__ ws, float* dw,
                                                        int cout, int cin, int taps, int a_is_x) {
  const long long total = (long long)cout * cin * taps;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int tap = (int)(idx % taps);
    const long long oc = idx / taps;
    const int ci = (int)(oc % cin), o = (int)(oc / cin);
    const long long src = a_is_x ? ((long long)tap * cin + ci) * cout + o
                                 : ((long long)tap * cout + o) * cin + ci;
    dw[idx] += ws[src];
  }
}

template <int BN, int STAGES, int OCC, int TPC>
static int launch_wgrad(const CUtensorMap& tmA, const CUtensorMap& tmB, const TcWgP& p, dim3 grid,
                        cudaStream_t st) {
  constexpr int smem = STAGES * (2 * WG_CHUNK_BYTES + TPC * (BN / 64) * WG_CHUNK_BYTES) + 1024 + 256;
  static_assert(smem * OCC <= 226 * 1024, "wgrad ring exceeds the shared memory of an SM");
  auto kern = conv_tc_wgrad_kernel<BN, STAGES, OCC, TPC>;
  OTM_ENSURE_SMEM(kern, smem);
  kern<<<grid, 256, smem, st>>>(tmA, tmB, p);
  OTM_LAUNCH_CHECK();
  return OTM_OK;
}

static void launch_wgrad_fold(float* ws, float* dw, int cout, int cin, int taps, int a_is_x,
                              cudaStream_t st);

// filter-row kernel: 3x3, both channel counts multiples of 128, a workspace to reduce into
static bool wgrad_row_eligible(const otm_conv_wgrad_args* a) {
  return a->kh == 3 && a->kw == 3 && a->x.c % 128 == 0 && a->dy.c % 128 == 0 && a->ws != nullptr;
}

static int conv_wgrad_tc_row(const otm_conv_wgrad_args* a, cudaStream_t st) {
  const int cin = a->x.c, cout = a->dy.c;
  TcWgRowP p;
  p.ws = a->ws; p.wfwd = (const __nv_bfloat16*)a->wfwd; p.P = a->P; p.Q = a->Q; p.rs = a->rs; p.cs = a->cs;
  p.alpha = a->alpha; p.cin = cin; p.cout = cout;
  p.x_coord_off = a->x_halo - a->pad;
  p.tiles_w = (a->dy.w + 7) / 8;
  p.tiles_total = p.tiles_w * ((a->dy.h + 7) / 8);
  p.m_tiles = cout / 128; p.n_tiles = cin / 128;
  p.units_total = (long long)a->dy.n * p.m_tiles * p.n_tiles * 3 * p.tiles_total;
  p.wfwd_n_stride = a->wfwd_batch_stride;
  if (p.P) OTM_REQUIRE(p.wfwd, "conv_wgrad: fused P needs wfwd");
  if (p.Q) OTM_REQUIRE(p.wfwd && !p.rs, "conv_wgrad: fused Q needs wfwd and no rs");
  CUtensorMap tmX, tmDy;
  int rc = make_act_map(&tmX, a->x, a->x_halo, 8, 10);
  if (rc) return rc;
  rc = make_act_map(&tmDy, a->dy, 0, 8, 8);
  if (rc) return rc;
  OTM_CHECK_CUDA(cudaMemsetAsync(p.ws, 0, sizeof(float) * (size_t)9 * cin * cout, st));
  constexpr int smem = WGR_STAGES * WGR_STAGE_BYTES + WGR_RED_BYTES + 512 + 1024 + 256;
  static_assert(smem <= 227 * 1024, "filter-column wgrad ring exceeds shared memory");
  auto kern = conv_tc_wgrad_row_kernel;
  OTM_ENSURE_SMEM(kern, smem);
  long long ctas = num_sms();
  if (ctas > p.units_total) ctas = p.units_total;
  kern<<<(int)ctas, 384, smem, st>>>(tmDy, tmX, p);
  OTM_LAUNCH_CHECK();
  launch_wgrad_fold(p.ws, a->dw, cout, cin, 9, 0, st);
  return OTM_OK;
}

int conv_wgrad_tc(const otm_conv_wgrad_args* a, cudaStream_t st) {
  if (wgrad_row_eligible(a)) return conv_wgrad_tc_row(a, st);
  OTM_REQUIRE(!a->Q, "conv_wgrad: the fused Q term is not available for this launch "
                     "(ask otm_conv_wgrad_fuses_Q first)");
  const int cin = a->x.c, cout = a->dy.c;
  TcWgP p;
  p.dw = a->dw; p.rs = a->rs; p.cs = a->cs; p.alpha = a->alpha;
  p.cin = cin; p.cout = cout; p.kh = a->kh; p.kw = a->kw;
  p.a_is_x = (cout % 128 != 0) ? 1 : 0;
  p.x_coord_off = a->x_halo - a->pad;
  const int M = p.a_is_x ? cin : cout, N = p.a_is_x ? cout : cin;
  int BN = (N % 256 == 0) ? 256 : (N % 128 == 0 ? 128 : 64);
  const int Ho = a->dy.h, Wo = a->dy.w;
  p.tiles_w = (Wo + 7) / 8;
  p.tiles_total = p.tiles_w * ((Ho + 7) / 8);
  const int taps = a->kh * a->kw;
  const int m_tiles = M / 128;
  p.n_tiles_n = N / BN;
  // taps per CTA: laying 2 / 4 shifted x tiles side by side as ONE 256-column operand next to the
  // shared dy tile was measured SLOWER here (717 vs 921 TFLOP/s at n=96 128->128: the 48 KB
  // stages leave a 2-deep ring at two CTAs per SM); the filter-column kernel above is the form
  // of that idea that works.  One tap per CTA.
  constexpr int tpc_cap = 1;
  int tpc = p.a_is_x ? 1 : 256 / BN;
  if (tpc > tpc_cap) tpc = tpc_cap;
  const int tap_ctas = (taps + tpc - 1) / tpc;
  long long base = (long long)tap_ctas * m_tiles * p.n_tiles_n * a->dy.n;
  int splits = (int)((4LL * num_sms() + base - 1) / base);  // ~2 waves at 2 CTAs per SM
  int max_splits = (p.tiles_total + 31) / 32;  // keep >= 32 K-stages per CTA (amortise the reds)
  if (splits > max_splits) splits = max_splits;
  if (splits < 1) splits = 1;
  p.splits = splits;
  OTM_REQUIRE((long long)a->dy.n * splits <= 65535, "wgrad_tc: grid.z too large");
  p.ws = a->ws;
  p.wfwd = (const __nv_bfloat16*)a->wfwd;
  p.wfwd_n_stride = a->wfwd_batch_stride;
  p.P = a->P;
  if (p.P) OTM_REQUIRE(!p.a_is_x && p.wfwd, "conv_wgrad: fused P needs Cout %% 128 == 0 and wfwd");
  CUtensorMap tmX, tmDy;
  int rc = make_act_map(&tmX, a->x, a->x_halo, 8, 8);
  if (rc) return rc;
  rc = make_act_map(&tmDy, a->dy, 0, 8, 8);
  if (rc) return rc;
  const CUtensorMap& tmA = p.a_is_x ? tmX : tmDy;
  const CUtensorMap& tmB = p.a_is_x ? tmDy : tmX;
  dim3 grid(tap_ctas, m_tiles * p.n_tiles_n, a->dy.n * splits);
  if (p.ws) OTM_CHECK_CUDA(cudaMemsetAsync(p.ws, 0, sizeof(float) * (size_t)taps * cin * cout, st));
  if (BN == 64) {
    rc = launch_wgrad<64, 4, 2, 1>(tmA, tmB, p, grid, st);
  } else if (BN == 128) {
    rc = launch_wgrad<128, 3, 2, 1>(tmA, tmB, p, grid, st);
  } else {
    rc = launch_wgrad<256, 2, 2, 1>(tmA, tmB, p, grid, st);
  }
  if (rc) return rc;
  if (p.ws) launch_wgrad_fold(p.ws, p.dw, cout, cin, taps, p.a_is_x, st);
  return OTM_OK;
}

static void launch_wgrad_fold(float* ws, float* dw, int cout, int cin, int taps, int a_is_x,
                              cudaStream_t st) {
  const long long total = (long long)taps * cin * cout;
  int blocks = (int)((total + 255) / 256);
  if (blocks > num_sms() * 8) blocks = num_sms() * 8;
  wgrad_fold_kernel<<<blocks, 256, 0, st>>>(ws, dw, cout, cin, taps, a_is_x);
  g_launches.fetch_add(1);
}

int conv_fwd_simt(const otm_conv_fwd_args* a, cudaStream_t st);
int conv_wgrad_simt(const otm_conv_wgrad_args* a, cudaStream_t st);

static int validate_fwd(const otm_conv_fwd_args* a) {
  OTM_REQUIRE(a && a->x.ptr && a->y.ptr && a->wpack, "conv_fwd: null argument");
  OTM_REQUIRE(a->kh >= 1 && a->kw >= 1 && a->pad >= 0 && a->x_halo >= 0 && a->y_halo >= 0,
              "conv_fwd: bad geometry");
  OTM_REQUIRE(a->y.h == a->x.h + 2 * a->pad - a->kh + 1 && a->y.w == a->x.w + 2 * a->pad - a->kw + 1,
              "conv_fwd: output %dx%d does not match input %dx%d k=%dx%d pad=%d", a->y.h, a->y.w,
              a->x.h, a->x.w, a->kh, a->kw, a->pad);
  OTM_REQUIRE(a->y.n == a->x.n, "conv_fwd: batch mismatch");
  OTM_REQUIRE(a->x_halo <= a->pad, "conv_fwd: x_halo (%d) larger than pad (%d)", a->x_halo, a->pad);
  OTM_REQUIRE(a->y_halo == 0 || (a->y_halo < a->y.h && a->y_halo < a->y.w),
              "conv_fwd: reflect halo too large");
  if (a->residual.ptr)
    OTM_REQUIRE(a->residual.n == a->y.n && a->residual.h == a->y.h && a->residual.w == a->y.w &&
                    a->residual.c == a->y.c && a->residual.dtype == a->y.dtype,
                "conv_fwd: residual mismatch");
  return OTM_OK;
}

}  // namespace otm

using namespace otm;

extern "C" {

int otm_conv_fwd_uses_tcgen05(const otm_conv_fwd_args* a) {
  if (!a || a->path == OTM_PATH_SIMT) return 0;
  return conv_fwd_tc_eligible(a) ? 1 : 0;
}

int otm_conv_fwd_fuses_stats(const otm_conv_fwd_args* a) {
  if (!a || a->path == OTM_PATH_SIMT || !conv_fwd_tc_eligible(a)) return 0;
  return conv_fwd_tc_uses_rr2t(a) ? 1 : 0;
}

int otm_conv_fwd_fuses_gate(const otm_conv_fwd_args* a) { return otm_conv_fwd_fuses_stats(a); }

int otm_conv_fwd(const otm_conv_fwd_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  int rc = validate_fwd(a);
  if (rc) return rc;
  const bool tc = conv_fwd_tc_eligible(a);
  if (a->stat_sums)
    OTM_REQUIRE(tc && a->path != OTM_PATH_SIMT && conv_fwd_tc_uses_rr2t(a),
                "conv_fwd: stat_sums given but this launch cannot accumulate them "
                "(ask otm_conv_fwd_fuses_stats first)");
  if (a->residual_mode != 0) {
    OTM_REQUIRE(a->residual_mode == 1 || a->residual_mode == 2, "conv_fwd: unknown residual_mode %d",
                a->residual_mode);
    OTM_REQUIRE(a->residual.ptr && !a->bias && a->act == OTM_ACT_NONE && !a->stat_sums,
                "conv_fwd: residual_mode 1 needs a gate tensor, no bias, no activation, no stat_sums");
    OTM_REQUIRE(tc && a->path != OTM_PATH_SIMT && conv_fwd_tc_uses_rr2t(a),
                "conv_fwd: residual_mode 1 is not available for this launch "
                "(ask otm_conv_fwd_fuses_gate first)");
  }
  if (a->path == OTM_PATH_TCGEN05 && !tc)
    return fail(OTM_ERR_UNSUPPORTED, "conv_fwd: tcgen05 path not available for this shape/dtype");
  if (tc && a->path != OTM_PATH_SIMT) return conv_fwd_tc(a, st);
  return conv_fwd_simt(a, st);
}

int otm_conv_wgrad_uses_tcgen05(const otm_conv_wgrad_args* a) {
  if (!a || a->path == OTM_PATH_SIMT) return 0;
  return conv_wgrad_tc_eligible(a) ? 1 : 0;
}

int64_t otm_conv_wgrad_workspace_bytes(const otm_conv_wgrad_args* a) {
  if (!a || a->path == OTM_PATH_SIMT || !conv_wgrad_tc_eligible(a)) return 0;
  return (int64_t)sizeof(float) * a->kh * a->kw * a->x.c * a->dy.c;
}

int otm_conv_wgrad_fuses_P(const otm_conv_wgrad_args* a) {
  if (!a || a->path == OTM_PATH_SIMT || !conv_wgrad_tc_eligible(a)) return 0;
  return (a->dy.c % 128 == 0) ? 1 : 0;
}

int otm_conv_wgrad_fuses_Q(const otm_conv_wgrad_args* a) {
  if (!a || a->path == OTM_PATH_SIMT || !conv_wgrad_tc_eligible(a)) return 0;
  return (a->kh == 3 && a->kw == 3 && a->x.c % 128 == 0 && a->dy.c % 128 == 0 && !a->rs) ? 1 : 0;
}

int otm_conv_wgrad(const otm_conv_wgrad_args* a, otm_stream stream) {
  cudaStream_t st = (cudaStream_t)stream;
  OTM_REQUIRE(a && a->x.ptr && a->dy.ptr && a->dw, "conv_wgrad: null argument");
  OTM_REQUIRE(a->dy.h == a->x.h + 2 * a->pad - a->kh + 1 && a->dy.w == a->x.w + 2 * a->pad - a->kw + 1 &&
                  a->dy.n == a->x.n,
              "conv_wgrad: dy %dx%d does not match x %dx%d k=%d pad=%d", a->dy.h, a->dy.w, a->x.h,
              a->x.w, a->kh, a->pad);
  OTM_REQUIRE(a->x_halo <= a->pad, "conv_wgrad: x_halo larger than pad");
  const bool tc = conv_wgrad_tc_eligible(a);
  if (a->path == OTM_PATH_TCGEN05 && !tc)
    return fail(OTM_ERR_UNSUPPORTED, "conv_wgrad: tcgen05 path not available for this shape/dtype");
  if (tc && a->path != OTM_PATH_SIMT) return conv_wgrad_tc(a, st);
  OTM_REQUIRE(!a->P && !a->Q, "conv_wgrad: the fused P / Q terms are only available on the tcgen05 path");
  return conv_wgrad_simt(a, st);
}

}  // extern "C"
