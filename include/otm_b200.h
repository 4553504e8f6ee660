/*
 * otm_b200.h — C-ABI of the B200-native one-to-many-GAN training-step kernels.
 *
 * The reference (struan-robertson/one-to-many-gan) has no FFI or plugin boundary:
 * every GPU op is a PyTorch library call made from src/model/*.py and
 * src/core/training.py.  This header is the boundary a maintainer would bind instead
 * (ctypes stub in INTEGRATION.md).  Each entry point names the reference call site it
 * replaces.
 *
 * Conventions
 *  - plain pointers and sizes only; no torch types.  The CALLER allocates every output
 *    and workspace (the library never owns device memory).
 *  - every call only ENQUEUES work on `stream` (no synchronisation, CUDA-graph
 *    capturable).  One call at a time per stream.
 *  - return 0 on success, negative otm_status on failure; otm_last_error() returns a
 *    thread-local message.
 *  - activations are NHWC "views": channel stride is 1, the n/h/w strides are given in
 *    ELEMENTS, so an interior view of a halo-padded buffer is a valid tensor.  `halo`
 *    fields say how many pixels around the view are materialised and readable/writable.
 *  - dtype: OTM_F32 or OTM_BF16 storage; all arithmetic accumulates in fp32.
 */
#ifndef OTM_B200_H
#define OTM_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* otm_stream; /* cudaStream_t */

enum otm_status {
  OTM_OK = 0,
  OTM_ERR_INVALID = -1,     /* bad argument / unsupported shape */
  OTM_ERR_CUDA = -2,        /* a CUDA runtime or driver call failed */
  OTM_ERR_UNSUPPORTED = -3, /* requested path not available for this shape/dtype */
};

enum otm_dtype { OTM_F32 = 0, OTM_BF16 = 1 };
enum otm_act { OTM_ACT_NONE = 0, OTM_ACT_RELU = 1, OTM_ACT_LRELU = 2, OTM_ACT_TANH = 3 };
enum otm_path { OTM_PATH_AUTO = 0, OTM_PATH_SIMT = 1, OTM_PATH_TCGEN05 = 2 };

typedef struct {
  void* ptr;     /* element (n=0,h=0,w=0,c=0) of the logical view */
  int32_t dtype; /* otm_dtype */
  int32_t n, h, w, c;
  int64_t sn, sh, sw; /* element strides; channel stride is 1 */
} otm_tensor;

const char* otm_last_error(void);
int otm_version(void);
/* number of kernels this library has launched since load (bench.py's gpu_launches) */
int64_t otm_launch_count(void);

/* ---------------------------------------------------------------------------------
 * Convolution (stride 1) as implicit GEMM.
 * Replaces F.conv2d in EqualisedConv2d.forward (reference src/model/layers.py:82-102)
 * and the grouped F.conv2d of Conv2dWeightModulate.forward (layers.py:163-168), plus
 * the cuDNN dgrad/wgrad that autograd runs for both.
 *
 *   y[n,h,w,o] = epi( alpha * sum_{r,s,i} x[n, h+r-pad, w+s-pad, i] * wpack[nb][o][r][s][i] )
 *   epi(v) = act( v * row_scale[n,o] + bias[o] ) * post_scale[n,o] + residual[n,h,w,o]
 *
 * post_scale is how a modulated conv gets SHARED weights (y = sigma_inv *
 * conv(s * x, cW)): the producer of x multiplies its output by the consumer's style scale
 * s[n,i] after its activation, so the consumer reads x~ = s * x and its weights need no
 * per-sample copy.
 *
 * x positions outside [-x_halo, H+x_halo) x [-x_halo, W+x_halo) read as zero; positions
 * inside the halo read the materialised halo (e.g. a reflect halo written by a producer).
 * Output size: y.h = x.h + 2*pad - kh + 1 (same for w).  If y_halo > 0 the epilogue also
 * writes the REFLECT halo of that width around y (nn.ReflectionPad2d of the consumer,
 * reference blocks.py:21,25,49,54; builder.py:162,202).
 * dgrad = the same call on dy with a flipped/transposed pack and pad' = k-1-pad.
 * --------------------------------------------------------------------------------- */
typedef struct {
  otm_tensor x;
  int32_t x_halo;
  const void* wpack;      /* [wbatch][Cout][kh][kw][Cin], dtype == x.dtype */
  int64_t w_batch_stride; /* elements between per-sample packs; 0 = shared weights */
  int32_t kh, kw, pad;
  otm_tensor y;
  int32_t y_halo;
  float alpha;
  const float* row_scale; /* [n, Cout] or NULL */
  const float* bias;      /* [Cout] or NULL */
  int32_t act;            /* otm_act */
  otm_tensor residual;    /* ptr NULL = none */
  int32_t path;           /* otm_path */
  const float* post_scale; /* [n, Cout] or NULL */
  float* stat_sums;       /* [n, Cout, 2] fp32 or NULL: the epilogue also accumulates sum(y) and
                           * sum(y^2) per (n, channel) -- the InstanceNorm statistics of the
                           * consumer (reference blocks.py:23,27; nn.InstanceNorm2d) without a
                           * separate read of y.  Zeroed by this call.  Only when
                           * otm_conv_fwd_fuses_stats() says so; finish with otm_instnorm_finalize. */
  int32_t residual_mode;  /* 0: y = epi(v) + residual (above).
                           * 1 ("gate"): `residual` is not added, it GATES the result and is
                           *   reduced against the raw product:
                           *     y[n,h,w,o]     = residual != 0 ? v * row_scale * post_scale : 0
                           *     dot_sums[n,o]  = sum_hw v * residual         (v = alpha * conv)
                           *   This is the input-side pass of a modulated conv's backward
                           *   (otm_mod_in with relu_mask) done in the dgrad's epilogue: `residual`
                           *   is the (style-scaled, so possibly negative) ReLU output the gradient
                           *   is taken w.r.t.  bias must be NULL
                           *   and act NONE.  Only when otm_conv_fwd_fuses_gate() says so.
                           * 2 ("dot"): as 1 without the gate -- y = v * row_scale * post_scale,
                           *   dot_sums = sum_hw v * residual: the style-gradient reduction of an
                           *   up-sampling modulated conv (otm_mod_in with no gx) in its dgrad. */
  float* dot_sums;        /* [n, Cout] fp32 (mode 1), zeroed by this call */
} otm_conv_fwd_args;
int otm_conv_fwd(const otm_conv_fwd_args* a, otm_stream stream);
/* 1 if the tcgen05 path would be used for these arguments, 0 if SIMT */
int otm_conv_fwd_uses_tcgen05(const otm_conv_fwd_args* a);
/* 1 if this call would run on the kernel whose epilogue can accumulate stat_sums */
int otm_conv_fwd_fuses_stats(const otm_conv_fwd_args* a);
/* 1 if this call would run on the kernel whose epilogue implements residual_mode 1 */
int otm_conv_fwd_fuses_gate(const otm_conv_fwd_args* a);

/* Reflect-border correction of a 3x3 dgrad.  The gradient of conv3x3(ReflectionPad2d(1)(x))
 * (reference blocks.py:21-27,49-56) w.r.t. x is the zero-padded "same" correlation of dy with
 * the flipped pack -- one otm_conv_fwd call with pad = 1 on the H x W output, no padded
 * (H+2) x (W+2) intermediate and no fold pass -- PLUS the contributions of the 2(H+W)+4 halo
 * positions, which land on rows / columns 1 and H-2 / W-2.  This call computes those as four
 * thin GEMMs (one per border line: positions x 3 taps x K) on warp-level tensor cores and adds
 * them into y with the same per-channel epilogue as the main call:
 *   y[n, refl(a), refl(b), o] += f(n,o) * gxp[a,b,o]   for (a,b) on the halo ring,
 *   gxp[a,b,o] = sum_{r,s,i} dy[n, a+r-1, b+s-1, i] * wpack[nb][o][r][s][i],
 *   f = row_scale * post_scale, and with `gate`: f *= (gate[n,refl(a),refl(b),o] != 0),
 *   dot_sums[n,o] += gxp * gate  (accumulated, NOT zeroed: call it after the main otm_conv_fwd).
 * bf16 only; K % 64 == 0, Cout % 64 == 0, H, W >= 3. */
typedef struct {
  otm_tensor dy;          /* [n, K, H, W]: the input of the dgrad call */
  const void* wpack;      /* [wbatch][Cout][3][3][K] (the dgrad pack of otm_weight_pack) */
  int64_t w_batch_stride; /* 0 = shared */
  otm_tensor y;           /* [n, Cout, H, W], accumulated into */
  const float* row_scale; /* [n, Cout] or NULL */
  const float* post_scale;
  otm_tensor gate;        /* ptr NULL = none */
  float* dot_sums;        /* [n, Cout] or NULL */
} otm_conv_reflect_border_args;
int otm_conv_reflect_border(const otm_conv_reflect_border_args* a, otm_stream stream);

/* wgrad:  dw[o][i][r][s] (fp32, torch parameter layout [Cout,Cin,kh,kw], ACCUMULATED into)
 *   += alpha * sum_{n,h,w} rs[n,o] * cs[n,i] * dy[n,h,w,o] * x[n, h+r-pad, w+s-pad, i]
 * rs / cs (per-sample demodulation / modulation factors, reference layers.py:148-161)
 * may be NULL.  x obeys the same halo/zero rule as the forward. */
typedef struct {
  otm_tensor x;
  int32_t x_halo;
  otm_tensor dy;
  int32_t kh, kw, pad;
  float* dw;
  float alpha;
  const float* rs; /* [n, Cout] or NULL */
  const float* cs; /* [n, Cin] or NULL */
  int32_t path;
  float* ws; /* optional fp32 workspace [kh*kw*Cout*Cin] for the tcgen05 path: split-K partials
                are reduced there with 128-bit vector reds in GEMM-tile order and then folded
                into dw by one pass; NULL = reduce straight into dw with scalar reds */
  /* Optional fused demodulation-gradient term of the modulated conv:
   *   P[n,o] += rs[n,o] * sum_{r,s,i} G_n[o,i,r,s] * wfwd[n][o][r][s][i]
   * where G_n is the un-scaled per-sample weight gradient this kernel already holds in TMEM and
   * wfwd the forward pack: per-sample (alpha*w*cs, wfwd_batch_stride = Cout*kh*kw*Cin) when x is
   * the un-modulated input, or the ONE shared pack (alpha*w, stride 0) when x already carries the
   * modulation (x~ = s*x).  This equals sum_hw dy*y, so the separate pass over dy and y
   * (otm_mod_out) is not needed.  tcgen05 path with Cout % 128 == 0 only
   * (otm_conv_wgrad_fuses_P tells); P must be zeroed by the caller. */
  const void* wfwd;
  float* P;
  int64_t wfwd_batch_stride;
  /* Optional fused direct style-gradient term (the other reduction of the same accumulator):
   *   Q[n,i] += sum_{o,r,s} G_n[o,i,r,s] * wfwd[n | 0][o][r][s][i]
   * With x the un-modulated input of a modulated conv and wfwd the shared pack alpha*w this is
   * sum_hw (dL/d(s*x)) * x, i.e. what otm_mod_in reduces from the dgrad's output: the dgrad then
   * needs no input-side pass (its epilogue applies s and adds the skip gradient).  Filter-column
   * tcgen05 kernel only, rs must be NULL (otm_conv_wgrad_fuses_Q tells); Q zeroed by the caller;
   * needs the workspace `ws`. */
  float* Q;
} otm_conv_wgrad_args;
int otm_conv_wgrad(const otm_conv_wgrad_args* a, otm_stream stream);
int otm_conv_wgrad_uses_tcgen05(const otm_conv_wgrad_args* a);
int otm_conv_wgrad_fuses_P(const otm_conv_wgrad_args* a);
int otm_conv_wgrad_fuses_Q(const otm_conv_wgrad_args* a);
/* bytes of the fp32 workspace `ws` this call wants (0: none, e.g. the FFMA path): the caller
 * allocates it -- the library never owns memory ("otm_query_workspace"). */
int64_t otm_conv_wgrad_workspace_bytes(const otm_conv_wgrad_args* a);

/* Weight staging.  Replaces EqualisedWeight.forward (layers.py:23-24) and the per-sample
 * `weights * s` materialisation (layers.py:152-161).
 *   transpose == 0: out[b][o][r][s][i]           = alpha * w[o][i][r][s] * cs[b,i] * rs[b,o]
 *   transpose == 1: out[b][i][kh-1-r][kw-1-s][o] = same value          (dgrad pack)
 * cs / rs may be NULL (factor 1); nb = number of per-sample packs (1 = shared). */
typedef struct {
  const float* w; /* [Cout, Cin, kh, kw] */
  int32_t cout, cin, kh, kw;
  float alpha;
  const float* cs; /* [nb, Cin] or NULL */
  const float* rs; /* [nb, Cout] or NULL */
  int32_t nb;
  int32_t transpose;
  void* out;
  int32_t out_dtype;
} otm_weight_pack_args;
int otm_weight_pack(const otm_weight_pack_args* a, otm_stream stream);
/* njobs SHARED packs (nb == 1, cs == rs == NULL, one out_dtype, inner packed dimension a multiple
 * of 8) in one launch: every staged pack of a network right after its optimiser step
 * (reference train.py:94-116 -> EqualisedWeight.forward layers.py:23-24 of the next iteration). */
int otm_weight_pack_multi(const otm_weight_pack_args* jobs, int32_t njobs, otm_stream stream);

/* q[o,i] = sum_k (alpha*w[o,i,k])^2   (layers.py:156-158, dense form SURVEY App. B.2) */
int otm_weight_sqsum(const float* w, int32_t cout, int32_t cin, int32_t taps, float alpha,
                     float* q, otm_stream stream);
/* sigma_inv[b,o] = rsqrt(sum_i s[b,i]^2 q[o,i] + eps) */
int otm_demod(const float* s, const float* q, int32_t nb, int32_t cout, int32_t cin, float eps,
              float* sigma_inv, otm_stream stream);
/* Backward of the modulation coefficients given P[b,o] = sum_hw dy*y and
 * Q[b,i] = sum_hw dxt*x:
 *   dd[b,o] = -0.5 * sigma_inv^2 * P ;  ds[b,i] = Q[b,i] + 2 s[b,i] sum_o dd[b,o] q[o,i]
 *   dw[o,i,k] += 2 * alpha^2 * w[o,i,k] * sum_b dd[b,o] s[b,i]^2 */
typedef struct {
  const float* w;
  int32_t cout, cin, taps;
  float alpha;
  const float* s;
  const float* sigma_inv;
  const float* q;
  const float* P;
  const float* Q;
  int32_t nb;
  float* ds; /* [nb, Cin] written */
  float* dw; /* [Cout,Cin,taps] accumulated */
  int32_t q_scaled; /* 1: Q was reduced against the MODULATED input x~ = s*x (the un-modulated x is
                       not stored): Q[b,i] is divided by s[b,i] here.  A style scale that is exactly
                       0 loses that term (x~ == 0 carries no information about x): measure zero. */
} otm_mod_bwd_args;
int otm_mod_bwd(const otm_mod_bwd_args* a, otm_stream stream);

/* ---------------------------------------------------------------------------------
 * Instance norm / activation / padding passes (HBM-bound).
 * Replace nn.InstanceNorm2d (eps 1e-5, biased variance), ReLU / LeakyReLU(0.2) / Tanh,
 * nn.ReflectionPad2d and the residual adds of reference blocks.py:20-33 and
 * builder.py:161-176,268-284.
 * --------------------------------------------------------------------------------- */
/* stats[n,c,0..1] = (mean, rstd) over h*w of x.  ws: fp32 workspace [n*c*2], zeroed here. */
/* stats[i,0..1] = (mean, rstd) from sums[i,0..1] = (sum, sum of squares) over hw elements,
 * i < count = n * c (the sums a conv epilogue accumulated, otm_conv_fwd_args.stat_sums). */
int otm_instnorm_finalize(const float* sums, float* stats, int32_t count, int32_t hw, float eps,
                          otm_stream stream);
int otm_instnorm_stats(const otm_tensor* x, float eps, float* ws, float* stats,
                       otm_stream stream);
/* y = act((x - mean) * rstd) + residual, optional reflect halo of width y_halo around y.
 * stats NULL = no normalisation. */
typedef struct {
  otm_tensor x;
  const float* stats;
  int32_t act;
  otm_tensor residual; /* ptr NULL = none */
  otm_tensor y;
  int32_t y_halo;
} otm_norm_act_args;
int otm_norm_act(const otm_norm_act_args* a, otm_stream stream);

/* Backward of otm_norm_act.  g is the gradient w.r.t. y; when g_halo > 0 it is the
 * gradient w.r.t. the reflect-padded y (size +2*g_halo) and is folded back on load.
 *   ga  = fold(g) [+ g2]                (optional second gradient source, interior size)
 *   gn  = ga * act'(.)                   gres = ga (written if gres.ptr != NULL)
 *   gx  = rstd * (gn - mean_hw(gn) - yhat * mean_hw(gn*yhat))       (stats != NULL)
 * sums: fp32 workspace [n*c*2]. */
typedef struct {
  otm_tensor g;
  int32_t g_halo;
  otm_tensor g2; /* ptr NULL = none */
  otm_tensor x;  /* forward input (pre-norm) */
  const float* stats;
  int32_t act;
  otm_tensor gx;
  otm_tensor gres; /* ptr NULL = none */
  float* sums;
  int32_t g_down; /* 1: g is the gradient w.r.t. DownSample(y) (shape [n,H/2,W/2,c]); the
                     transposed blur+bilinear stencil is applied on load, so the backward of the
                     fused otm_down needs no full-resolution intermediate */
} otm_norm_act_bwd_args;
int otm_norm_act_bwd(const otm_norm_act_bwd_args* a, otm_stream stream);

/* Second derivative of otm_norm_act (InstanceNorm + a piecewise-linear activation) for the R1
 * gradient penalty (BASELINE config 5: double backward through the discriminator).  With g the
 * gradient the FIRST backward received (w.r.t. the activation output) and gg the gradient w.r.t.
 * that backward's result gx:   dg = d L / d g,   dx = d L / d x   (either ptr may be NULL).
 * sums: fp32 workspace [n*c*5]. */
int otm_norm_act_bwd_bwd(const otm_tensor* g, const otm_tensor* gg, const otm_tensor* x,
                         const float* stats, int32_t act, const otm_tensor* dg, const otm_tensor* dx,
                         float* sums, otm_stream stream);

/* ---------------------------------------------------------------------------------
 * Resampling stencils.  Replace Smooth / UpSample / DownSample
 * (reference layers.py:191-247): replicate-pad 3x3 binomial blur, bilinear x2
 * (align_corners=False) and bilinear to (H//2, W//2) with scale H/(H//2).
 * otm_down fuses the producer's normalise+activation in front of the stencil.
 * --------------------------------------------------------------------------------- */
typedef struct {
  otm_tensor x;
  const float* stats; /* NULL = none */
  int32_t act;
  otm_tensor y; /* [n, H/2, W/2, c] */
  int32_t y_halo;
} otm_down_args;
int otm_down(const otm_down_args* a, otm_stream stream);
/* ga[n,H,W,c] = transpose(stencil)(g);  g_halo folds a reflect-padded g first. */
int otm_down_bwd(const otm_tensor* g, int32_t g_halo, const otm_tensor* ga, otm_stream stream);
/* scale: optional [n, c] per-sample channel factor on the result (the style scale of the
 * modulated conv that consumes the up-sampled tensor; the stencil is per channel, so it commutes) */
int otm_up(const otm_tensor* x, const otm_tensor* y, int32_t y_halo, const float* scale,
           otm_stream stream);
int otm_up_bwd(const otm_tensor* g, int32_t g_halo, const otm_tensor* gx, const float* scale,
               otm_stream stream);

/* ---------------------------------------------------------------------------------
 * Modulated-conv side passes (HBM-bound).  SURVEY App. B.2.
 * --------------------------------------------------------------------------------- */
/* out side:  gy = fold(g)[+g2] * act'(out) ;  P[n,o] += sum_hw gy * (out - res)
 * (out is the saved conv output AFTER activation/residual; act in {NONE, RELU}) */
typedef struct {
  otm_tensor g;
  int32_t g_halo;
  otm_tensor g2;
  otm_tensor out;
  otm_tensor res; /* ptr NULL = none */
  int32_t act;
  otm_tensor gy; /* ptr NULL = do not materialise (act NONE, no fold, no g2) */
  float* P;      /* [n, c] zeroed here then accumulated */
  const float* gy_scale; /* [n, c] or NULL: the STORED gy is multiplied by it (sigma_inv: gy is then
                            the gradient w.r.t. the conv's raw output u, ready for a shared-weight
                            dgrad / wgrad); P is reduced from the un-scaled value */
} otm_mod_out_args;
int otm_mod_out(const otm_mod_out_args* a, otm_stream stream);
/* in side:  gxt = fold(g_padded) ; Q[n,i] = sum_hw gxt * x ; gx = s[n,i] * gxt [+ gadd]
 * (gx.ptr NULL: reduction only) */
typedef struct {
  otm_tensor g;
  int32_t g_halo;
  otm_tensor x;
  const float* s; /* [n, c] */
  otm_tensor gadd; /* ptr NULL = none */
  otm_tensor gx;
  float* Q; /* [n, c] zeroed here then accumulated */
  int32_t relu_mask; /* 1: gx *= (x != 0) -- x is a (possibly style-scaled) ReLU output, so this is
                        the ReLU backward of the producer fused into this pass */
  const float* gx_scale; /* [n, c] or NULL: gx *= gx_scale after the mask (the producer conv's
                            sigma_inv: gx is then the gradient w.r.t. ITS raw output u) */
} otm_mod_in_args;
int otm_mod_in(const otm_mod_in_args* a, otm_stream stream);

/* per-channel sum over n,h,w of g: bias gradients (dy -> db[c]).  accumulate == 0: out is zeroed
 * here first; != 0: out += (straight into the gradient arena) */
int otm_channel_sum(const otm_tensor* g, float* out, int32_t accumulate, otm_stream stream);
/* global average pool fwd/bwd (StyleExtractor head, builder.py:314) */
int otm_avgpool(const otm_tensor* x, float* out /*[n,c] fp32*/, otm_stream stream);
int otm_avgpool_bwd(const float* g /*[n,c]*/, const otm_tensor* gx, otm_stream stream);

/* ---------------------------------------------------------------------------------
 * Losses (reference src/core/training.py:111-117,178-190,202-204; src/model/loss.py).
 * Every call writes its scalar(s) to `out` (fp32, device) and, when grad.ptr != NULL,
 * the backward seed d(scale*loss)/dx in the same pass.
 * --------------------------------------------------------------------------------- */
/* out[0] = mean((x - target)^2); out[1] = mean(sign(2x-1));  grad = scale*2(x-target)/N */
int otm_loss_lsgan(const otm_tensor* x, float target, float scale, float* out,
                   const otm_tensor* grad, otm_stream stream);
/* out[0] = mean|a-b| ; grad = scale*sign(a-b)/N (w.r.t. a) */
int otm_loss_l1(const otm_tensor* a, const otm_tensor* b, float scale, float* out,
                const otm_tensor* grad, otm_stream stream);
/* moments: out[0] = sum x, out[1] = sum x^2 over the whole tensor (kl_loss_func, loss.py:82-92) */
int otm_moments(const otm_tensor* x, float* out, otm_stream stream);
/* grad[...] (+)= coef[0] + coef[1]*x   (device-side coefficients; KL backward) */
int otm_affine_grad(const otm_tensor* x, const float* coef, const otm_tensor* grad,
                    int32_t accumulate, otm_stream stream);
/* path_loss_func (loss.py:98-111) for one feature pair:
 *   out[0] += weight * mean(((f1-f2)/h[n])^2) ; g1 = scale*weight*2(f1-f2)/(h^2 N), g2 = -g1 */
int otm_loss_path(const otm_tensor* f1, const otm_tensor* f2, const float* h, float weight,
                  float scale, float* out, const otm_tensor* g1, const otm_tensor* g2,
                  otm_stream stream);

/* out[0] = style_cycle_loss_func(a, b) (reference loss.py:60-75: both L2-normalised with eps
 * 1e-12, 1 - mean cosine similarity (eps 1e-8) + ratio * mse) on [batch, features] fp32 rows
 * (row strides in elements); da / db (dense [batch, features], may be NULL) = scale * d loss. */
int otm_loss_style_cycle(const float* a, int64_t a_stride, const float* b, int64_t b_stride,
                         int32_t batch, int32_t features, float ratio, float scale, float* out,
                         float* da, float* db, otm_stream stream);

/* ---------------------------------------------------------------------------------
 * Small dense layers (fp32, SIMT: K is 6 or 512, below a tensor-core tile).
 * --------------------------------------------------------------------------------- */
#define OTM_MAX_LINEAR_JOBS 16
#define OTM_MAX_STYLE_DIM 32
#define OTM_MAX_MAPPING_LAYERS 8
/* EqualisedLinear (reference layers.py:27-43): y = x @ (c W)^T + bias, c = 1/sqrt(k), W the raw
 * parameter.  Up to OTM_MAX_LINEAR_JOBS independent layers per call -- e.g. every `to_style` of
 * one decoder pass (layers.py:138-140,148) or the StyleExtractor head (builder.py:316) -- run as
 * ONE launch.  Backward: dw / dbias are accumulated (+=, each written by one thread), dx is
 * accumulated atomically (jobs may share it); jobs with dy == NULL are skipped. */
typedef struct {
  const float* x;       /* [n, k] */
  int64_t x_row_stride; /* elements; 0 = one row broadcast over n */
  const float* w;       /* [o, k] */
  const float* bias;    /* [o] or NULL */
  float* y;             /* forward: [n, o] dense */
  const float* dy;      /* backward: [n, o] dense, or NULL */
  float* dw;            /* backward: [o, k] += , or NULL */
  float* dbias;         /* backward: [o] +=, or NULL */
  float* dx;            /* backward: rows of k, atomically +=, or NULL */
  int64_t dx_row_stride;
  int32_t n, k, o;
} otm_linear_job;
int otm_linear_fwd(const otm_linear_job* jobs, int32_t n_jobs, otm_stream stream);
int otm_linear_bwd(const otm_linear_job* jobs, int32_t n_jobs, otm_stream stream);

/* MappingNetwork.forward (reference builder.py:46-49: F.normalize, [Linear, LeakyReLU 0.2] x
 * (L-1), Linear, ReLU) fused with the style mixing of _get_style_vector (builder.py:115-132) and
 * the domain-variable interpolation of get_single_w / get_two_w (builder.py:66-71,104):
 *   out[j][blk, b, :] = d_j[b] * (blk < *cross ? net(z1[b]) : net(z2[b])),   j = 0, 1
 * The host draws z1, z2 and the crossover (the reference's host-RNG order) and copies them to the
 * device; everything else is this one launch.  Backward accumulates dw / db from dout[j]. */
typedef struct {
  const float* z1;   /* [batch, features] */
  const float* z2;   /* NULL = no mixing */
  const void* cross; /* device int64 scalar, NULL = all blocks take z1 */
  const float* w[OTM_MAX_MAPPING_LAYERS]; /* [features, features] raw parameters */
  const float* b[OTM_MAX_MAPPING_LAYERS];
  float* dw[OTM_MAX_MAPPING_LAYERS];      /* backward: accumulated */
  float* db[OTM_MAX_MAPPING_LAYERS];
  int32_t features, n_layers, batch, n_blocks;
  const float* d[2];  /* [batch] per-sample domain variable, NULL = d_const[j] */
  float d_const[2];
  float* out[2];        /* [n_blocks, batch, features]; out[1] NULL = one output */
  const float* dout[2]; /* backward */
} otm_mapping_args;
int otm_mapping_fwd(const otm_mapping_args* a, otm_stream stream);
int otm_mapping_bwd(const otm_mapping_args* a, otm_stream stream);

/* ---------------------------------------------------------------------------------
 * Optimiser and data.
 * --------------------------------------------------------------------------------- */
/* torch.optim.Adam defaults (reference train.py:94-116) over a flat fp32 arena.
 * `step` is read from the device (int32) so the call is graph-replayable; grad is
 * multiplied by grad_scale first (1/world for DDP). */
typedef struct {
  float* param;
  const float* grad;
  float* m;
  float* v;
  int64_t n;
  float lr, beta1, beta2, eps, grad_scale;
  const int32_t* step; /* device pointer to the 1-based step count */
} otm_adam_args;
int otm_adam(const otm_adam_args* a, otm_stream stream);

/* Synthetic U(-1,1) batch (replaces src/data/datasets.py + DataLoader, train.py:120-169)
 * Philox4x32-10 keyed by (seed, stream_id), counter = offset + element index. */
int otm_synth_uniform(float* out, int64_t n, uint64_t seed, uint64_t stream_id,
                      uint64_t offset, otm_stream stream);

/* Device-resident real-data path (replaces ShoeDataset.__getitem__, the transform chain and the
 * DataLoaders: reference src/data/datasets.py:13-50, train.py:120-169).  data: the resized image
 * folder as uint8 [n_images, c, h, w] in HBM; idx [batch] int64 and flip [batch] uint8 (NULL = no
 * flips) on the device.  out[b,c,y,x] = data[idx[b], c, y, flip[b] ? w-1-x : x] * 2/255 - 1
 * (= Normalize(0.5, 0.5)(ToTensor(.))), fp32 [batch, c, h, w]. */
int otm_gather_batch(const uint8_t* data, int64_t n_images, int32_t c, int32_t h, int32_t w,
                     const int64_t* idx, const uint8_t* flip, int32_t batch, float* out,
                     otm_stream stream);

/* generic helpers */
int otm_cast(const otm_tensor* x, const otm_tensor* y, otm_stream stream); /* dtype/stride copy */
int otm_add_inplace(const otm_tensor* dst, const otm_tensor* src, otm_stream stream);

#ifdef __cplusplus
}
#endif
#endif /* OTM_B200_H */
