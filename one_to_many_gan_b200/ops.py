"""Fused stages of the training step as torch.autograd.Functions over the C-ABI kernels.

Each Function is one *block* of the reference network with a hand-scheduled backward
(SURVEY.md App. B): the convolutions run as implicit GEMMs (tcgen05 for the dense bf16
layers, FFMA in fp32 parity mode), everything between two convolutions is one HBM pass.
Tensors are logically NCHW / physically NHWC; a tensor "with halo p" is the interior view
of a buffer whose p-pixel reflect border has been materialised by its producer, so the
consumer convolution reads nn.ReflectionPad2d(p) for free.

Block-boundary gradients are always the true (interior-shaped) gradients; gradients of
halo-padded intermediates stay padded inside a block and are folded on load by the next
pass."""

from __future__ import annotations

import math

import torch

from . import kernels as K

ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = K.ACT_NONE, K.ACT_RELU, K.ACT_LRELU, K.ACT_TANH

# ---------------------------------------------------------------------------
# weight staging cache: packs are rebuilt when the parameter changes
# ---------------------------------------------------------------------------
# Entries are tagged with the optimiser that owns the weight (FlatAdam marks its parameters), so
# an optimiser step drops ITS packs only: inside the captured iteration Adam(D) runs in the
# middle of the generator step and must not throw away the generator's packs.
_EPOCH = 0
_PACKS: dict = {}


def invalidate_packs(owner=None) -> None:
    """Called after weights changed in place (optimiser step, checkpoint load, graph replay).
    owner: only the packs of parameters tagged `_otm_owner == owner`."""
    global _EPOCH
    if owner is None:
        _EPOCH += 1
        _PACKS.clear()
        return
    for k in [k for k, v in _PACKS.items() if v[2] == owner]:
        del _PACKS[k]


def eq_scale(weight: torch.Tensor) -> float:
    """EqualisedWeight constant c = 1/sqrt(fan_in) (reference layers.py:19)."""
    return 1.0 / math.sqrt(weight[0].numel())


def _cached(key, weight: torch.Tensor, make):
    # The entry keeps a reference to the weight's storage, so its data_ptr cannot be recycled
    # by another tensor while the entry is alive (a bare data_ptr key would alias).
    hit = _PACKS.get(key)
    if hit is None:
        if len(_PACKS) > 512:
            _PACKS.clear()
        hit = (make(), weight.detach(), getattr(weight, "_otm_owner", None))
        _PACKS[key] = hit
    return hit[0]


def _pack(weight: torch.Tensor, dtype, transpose: bool):
    key = (weight.data_ptr(), weight._version, _EPOCH, dtype, transpose)
    if _RECORDER is not None and key not in _PACKS and K.weight_pack_multi_ok(weight, transpose):
        _RECORDER.note(weight, dtype, transpose)
    return _cached(key, weight, lambda: K.weight_pack(weight.detach(), eq_scale(weight), dtype,
                                                       transpose=transpose))


class PackRecorder:
    """Batches the shared weight packs of a repeating iteration.  The first time the iteration
    runs, every pack that is built lazily is noted under the phase it was first needed in (a phase
    starts where packs went stale: the start of the iteration, an optimiser step in the middle of
    it); from then on `phase()` rebuilds all of a phase's packs with ONE launch
    (otm_weight_pack_multi) and the lazy path only sees hits.  ~50 launches of 5 us per iteration
    become 2.  Owned by the caller (engine.TrainIteration): nothing is shared between models."""

    def __init__(self):
        self.phases: dict = {}
        self.current = None

    def note(self, weight, dtype, transpose):
        if self.current is None:
            return
        rec = self.phases.setdefault(self.current, {})
        rec.setdefault((weight.data_ptr(), dtype, transpose), (weight, dtype, transpose))

    def phase(self, name):
        """Enter phase `name` (None: stop recording) and stage everything noted for it so far."""
        global _RECORDER
        _RECORDER = self if name is not None else None
        self.current = name
        rec = self.phases.get(name)
        if not rec:
            return
        by_dtype: dict = {}
        for weight, dtype, transpose in rec.values():
            key = (weight.data_ptr(), weight._version, _EPOCH, dtype, transpose)
            if key not in _PACKS:
                by_dtype.setdefault(dtype, []).append((key, weight, transpose))
        for dtype, jobs in by_dtype.items():
            outs = K.weight_pack_multi([(w.detach(), eq_scale(w), t) for _, w, t in jobs], dtype)
            for (key, w, _), out in zip(jobs, outs):
                if len(_PACKS) > 512:
                    _PACKS.clear()
                _PACKS[key] = (out, w.detach(), getattr(w, "_otm_owner", None))


_RECORDER: PackRecorder | None = None


def _sqsum(weight: torch.Tensor):
    key = (weight.data_ptr(), weight._version, _EPOCH, "q")
    return _cached(key, weight, lambda: K.weight_sqsum(weight.detach(), eq_scale(weight)))


# Training fast path (training.backward_unit): the wgrad kernels ACCUMULATE, so they can add
# straight into `weight.grad` -- a view of the optimiser's flat gradient arena that zero_grad()
# cleared with one memset -- instead of a fresh zero-filled buffer that autograd then adds to
# .grad (that was ~100 fill + ~150 add launches per iteration).  The Function returns None for
# that weight.  Off by default so the stages stay ordinary autograd Functions.
DIRECT_WEIGHT_GRADS = False


def _wgrad_buffer(weight: torch.Tensor):
    """(buffer the wgrad kernels accumulate into, value to hand back to autograd)."""
    g = weight.grad
    if DIRECT_WEIGHT_GRADS and g is not None and g.dtype == torch.float32 and g.is_contiguous():
        return g, None
    buf = torch.zeros_like(weight)
    return buf, buf


def _param_grad_buffer(p: torch.Tensor):
    """Like _wgrad_buffer for any fp32 parameter (biases, to_style / mapping weights)."""
    g = p.grad
    if DIRECT_WEIGHT_GRADS and g is not None and g.dtype == torch.float32 and g.is_contiguous():
        return g, None
    buf = torch.zeros_like(p)
    return buf, buf


def nhwc(t: torch.Tensor, dtype=None) -> torch.Tensor:
    """API-boundary normalisation (plumbing): channel-stride-1 storage in `dtype`."""
    if dtype is not None and t.dtype != dtype:
        t = t.to(dtype)
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    ok = (w == 1 or sw > 0) and (h == 1 or sh > 0) and (n == 1 or sn > 0)
    if c != 1:
        ok = ok and sc == 1
    if not ok:
        t = t.contiguous(memory_format=torch.channels_last)
        if c != 1 and t.stride(1) != 1:  # degenerate spatial dims: force NHWC strides
            t = t.permute(0, 2, 3, 1).contiguous().permute(0, 3, 1, 2)
    return t


def halo_of(t: torch.Tensor) -> int:
    """Width of the reflect halo materialised around `t` by its producer (0 if unknown)."""
    return getattr(t, "_otm_halo", 0)


def with_halo(t: torch.Tensor, p: int) -> torch.Tensor:
    t._otm_halo = p
    return t


def _g(t: torch.Tensor, like: torch.Tensor) -> torch.Tensor:
    return nhwc(t, like.dtype)


# ---------------------------------------------------------------------------
# plain convolution (+bias, + sign-preserving activation)
# ---------------------------------------------------------------------------
class ConvFn(torch.autograd.Function):
    """EqualisedConv2d (reference layers.py:82-102) with the following activation fused when
    it is ReLU/LeakyReLU.  x may carry a reflect halo (x_halo) = the ReflectionPad2d in front
    of the conv; `pad` is the total logical padding (halo + zero padding)."""

    @staticmethod
    def forward(ctx, x, weight, bias, k, pad, x_halo, y_halo, act, out_dtype, bias_dead=False,
                want_stats=False):
        ctx.bias = bias
        ctx.bias_dead = bias_dead
        cout = weight.shape[0]
        wp = _pack(weight, x.dtype, False)
        y = K.conv_fwd(x, wp, cout, k, k, pad, x_halo=x_halo, y_halo=y_halo,
                       bias=bias.detach() if bias is not None else None, act=act,
                       out_dtype=out_dtype, want_stats=want_stats)
        stats = None
        if want_stats:
            y, stats = y
        ctx.cfg = (k, pad, x_halo, act)
        ctx.has_bias = bias is not None
        ctx.save_for_backward(x, weight, y if act != ACT_NONE else None)
        if want_stats:
            ctx.mark_non_differentiable(stats)
            return y, stats
        return y

    @staticmethod
    def backward(ctx, g, _gstats=None):
        x, weight, y = ctx.saved_tensors
        bias = ctx.bias
        k, pad, x_halo, act = ctx.cfg
        cin = weight.shape[1]
        g = nhwc(g)
        if act != ACT_NONE:  # relu / lrelu: derivative from the sign of the output
            g = K.norm_act_bwd(g, y, None, act)
        gw = gb = gx = None
        if ctx.needs_input_grad[1]:
            buf, gw = _wgrad_buffer(weight)
            K.conv_wgrad(x, g, buf, k, k, pad, x_halo=x_halo, alpha=eq_scale(weight))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            buf, gb = _param_grad_buffer(bias)
            # bias_dead: the conv feeds an InstanceNorm, which cancels the bias exactly -- its
            # gradient is identically 0 (sum_hw of the norm's input gradient vanishes);
            # the zero-initialised buffer IS the exact gradient, no reduction pass needed
            if not ctx.bias_dead:
                K.channel_sum(g, out=buf)
        if ctx.needs_input_grad[0]:
            wpt = _pack(weight, g.dtype, True)
            gxp = K.conv_fwd(g, wpt, cin, k, k, k - 1 - pad + x_halo, out_dtype=x.dtype)
            if x_halo:
                h, w = x.shape[2:]
                gint = gxp[:, :, x_halo : x_halo + h, x_halo : x_halo + w]
                gx = K.norm_act_bwd(gint, None, None, ACT_NONE, g_halo=x_halo)
            else:
                gx = gxp
        return gx, gw, gb, None, None, None, None, None, None, None, None


def conv(x, weight, bias, k, pad, *, x_halo=0, y_halo=0, act=ACT_NONE, out_dtype=None,
         bias_dead=False, want_stats=False):
    """bias_dead=True: the output goes straight into an InstanceNorm (the bias gradient is exactly
    zero and is not computed).  want_stats=True: returns (y, stats) with the InstanceNorm
    statistics of y -- from the conv epilogue where the kernel supports it -- to hand to
    `norm_act(..., stats=)` / `down(..., stats=)`."""
    return ConvFn.apply(x, weight, bias, k, pad, x_halo, y_halo, act, out_dtype or x.dtype, bias_dead,
                        want_stats)


# ---------------------------------------------------------------------------
# instance norm + activation (+ residual) (+ reflect halo)
# ---------------------------------------------------------------------------
class NormActFn(torch.autograd.Function):
    """nn.InstanceNorm2d -> activation -> (+ residual) -> ReflectionPad2d(y_halo) in one pass
    (reference blocks.py:20-33, builder.py:161-165)."""

    @staticmethod
    def forward(ctx, x, residual, norm, act, y_halo, stats=None):
        if not norm:
            stats = None
        elif stats is None:
            stats = K.instnorm_stats(x)
        y = K.norm_act(x, stats, act, residual=residual, y_halo=y_halo)
        ctx.act = act
        ctx.has_res = residual is not None
        ctx.save_for_backward(x, stats)
        return y

    @staticmethod
    def backward(ctx, g):
        x, stats = ctx.saved_tensors
        g = _g(g, x)
        want_res = ctx.has_res and ctx.needs_input_grad[1]
        r = K.norm_act_bwd(g, x, stats, ctx.act, want_gres=want_res)
        gx, gres = (r if want_res else (r, None))
        return gx, gres, None, None, None, None


def norm_act(x, *, norm=True, act=ACT_NONE, residual=None, y_halo=0, stats=None):
    return NormActFn.apply(x, residual, norm, act, y_halo, stats)


class DownFn(torch.autograd.Function):
    """(InstanceNorm -> activation ->) DownSample (reference layers.py:232-247) in one pass."""

    @staticmethod
    def forward(ctx, x, norm, act, y_halo, stats=None):
        if not norm:
            stats = None
        elif stats is None:
            stats = K.instnorm_stats(x)
        y = K.down(x, stats, act, y_halo)
        ctx.act = act
        ctx.save_for_backward(x, stats)
        return y

    @staticmethod
    def backward(ctx, g):
        x, stats = ctx.saved_tensors
        g = _g(g, x)
        if stats is None and ctx.act == ACT_NONE:
            return K.down_bwd(g, x.shape[2:]), None, None, None, None
        # (otm_norm_act_bwd can apply the stencil transpose on load (g_down=1) and skip this
        # temporary, but gathering 4-16 taps in both of its passes measured 0.6 ms/iteration
        # slower than materialising `ga` once, so the two-kernel form is used.)
        ga = K.down_bwd(g, x.shape[2:])
        return K.norm_act_bwd(ga, x, stats, ctx.act), None, None, None, None


def down(x, *, norm=False, act=ACT_NONE, y_halo=0, stats=None):
    return DownFn.apply(x, norm, act, y_halo, stats)


class UpFn(torch.autograd.Function):
    """UpSample: bilinear x2 + blur (reference layers.py:217-229)."""

    @staticmethod
    def forward(ctx, x):
        return K.up(x)

    @staticmethod
    def backward(ctx, g):
        return K.up_bwd(nhwc(g))


def up(x):
    return UpFn.apply(x)


class StackHaloFn(torch.autograd.Function):
    """Batch ranges of a halo-carrying tensor stacked along the batch, halo included:
    out = cat([src[a : a + k] for (a, k) in blocks]) as straight copies of the padded NHWC
    buffers.  The step batches its three decodes and its two path-length extractions this way
    (training.generator_losses); `torch.cat` on the interior views wrote an NCHW tensor that
    then needed a transposing copy and a halo pass (and the three matching passes in backward)."""

    @staticmethod
    def forward(ctx, src, blocks, halo):
        n = sum(k for _, k in blocks)
        _, c, h, w = src.shape
        out = K.alloc(n, c, h, w, src.dtype, src.device, halo)
        dst, sp = K.padded_view(out, halo), K.padded_view(src, halo)
        off = 0
        for a, k in blocks:
            dst[off : off + k].copy_(sp[a : a + k])
            off += k
        ctx.blocks, ctx.n_src = blocks, src.shape[0]
        return out

    @staticmethod
    def backward(ctx, g):
        g = nhwc(g)
        _, c, h, w = g.shape
        gs = K.alloc(ctx.n_src, c, h, w, g.dtype, g.device)
        covered = [False] * ctx.n_src
        off = 0
        for a, k in ctx.blocks:
            if not any(covered[a : a + k]):
                gs[a : a + k].copy_(g[off : off + k])
            elif all(covered[a : a + k]):
                gs[a : a + k].add_(g[off : off + k])
            else:
                raise ValueError("stack_halo: partially overlapping blocks")
            covered[a : a + k] = [True] * k
            off += k
        if not all(covered):  # batch entries no block reads get a zero gradient
            i = 0
            while i < ctx.n_src:
                if covered[i]:
                    i += 1
                    continue
                j = i
                while j < ctx.n_src and not covered[j]:
                    j += 1
                gs[i:j].zero_()
                i = j
        return gs, None, None


def stack_halo(src, blocks):
    """src: tensor whose producer materialised a reflect halo (halo_of(src) > 0); blocks:
    [(start, length)] batch ranges.  Returns the stacked tensor with the same halo."""
    halo = halo_of(src)
    if halo <= 0:
        raise ValueError("stack_halo needs a tensor with a materialised halo")
    return with_halo(StackHaloFn.apply(src, tuple(blocks), halo), halo)


class ResBlockFn(torch.autograd.Function):
    """ResnetBlock (reference blocks.py:9-33): x + IN(conv(refl(ReLU(IN(conv(refl(x))))))).
    x carries a reflect halo of 1."""

    @staticmethod
    def forward(ctx, x, w1, w2, y_halo):
        f = w1.shape[0]
        raw1, st1 = K.conv_fwd(x, _pack(w1, x.dtype, False), f, 3, 3, 1, x_halo=1, want_stats=True)
        t = K.norm_act(raw1, st1, ACT_RELU, y_halo=1)
        raw2, st2 = K.conv_fwd(t, _pack(w2, x.dtype, False), f, 3, 3, 1, x_halo=1, want_stats=True)
        out = K.norm_act(raw2, st2, ACT_NONE, residual=x, y_halo=y_halo)
        ctx.save_for_backward(x, w1, w2, raw1, st1, t, raw2, st2)
        return out

    @staticmethod
    def backward(ctx, g):
        x, w1, w2, raw1, st1, t, raw2, st2 = ctx.saved_tensors
        g = _g(g, x)
        n, f, h, w = x.shape
        g_raw2 = K.norm_act_bwd(g, raw2, st2, ACT_NONE)
        gw1 = gw2 = None
        if ctx.needs_input_grad[2]:
            buf, gw2 = _wgrad_buffer(w2)
            K.conv_wgrad(t, g_raw2, buf, 3, 3, 1, x_halo=1, alpha=eq_scale(w2))
        # dgrad through ReflectionPad2d(1): "same" dgrad + halo-ring correction where the library
        # has it (bf16), else the padded (h+2) x (w+2) dgrad folded by the next pass
        direct = K.dgrad_reflect_ok(g_raw2, f)
        if direct:
            gt = K.conv_dgrad_reflect(g_raw2, _pack(w2, g.dtype, True), f)
            g_raw1 = K.norm_act_bwd(gt, raw1, st1, ACT_RELU)
        else:
            gt_p = K.conv_fwd(g_raw2, _pack(w2, g.dtype, True), f, 3, 3, 2)
            g_raw1 = K.norm_act_bwd(gt_p[:, :, 1 : 1 + h, 1 : 1 + w], raw1, st1, ACT_RELU, g_halo=1)
        if ctx.needs_input_grad[1]:
            buf, gw1 = _wgrad_buffer(w1)
            K.conv_wgrad(x, g_raw1, buf, 3, 3, 1, x_halo=1, alpha=eq_scale(w1))
        gx = None
        if ctx.needs_input_grad[0]:
            if direct:  # the skip gradient is the dgrad's residual: no fold + add pass
                gx = K.conv_dgrad_reflect(g_raw1, _pack(w1, g.dtype, True), f, residual=g)
            else:
                gx_p = K.conv_fwd(g_raw1, _pack(w1, g.dtype, True), f, 3, 3, 2)
                gx = K.norm_act_bwd(gx_p[:, :, 1 : 1 + h, 1 : 1 + w], None, None, ACT_NONE, g_halo=1,
                                    g2=g)
        return gx, gw1, gw2, None


def res_block(x, w1, w2, *, y_halo):
    return ResBlockFn.apply(x, w1, w2, y_halo)


# ---------------------------------------------------------------------------
# modulated convolutions (reference layers.py:111-182, dense form SURVEY App. B.2)
# ---------------------------------------------------------------------------
def _modconv_fwd(x, s, weight, pad, x_halo, act, residual, y_halo):
    """y = act(sigma_inv[b,o] * conv(x, cW * s[b,i])) (+ residual).
    Returns (y, sigma_inv, per-sample forward pack)."""
    n = x.shape[0]
    cout = weight.shape[0]
    c = eq_scale(weight)
    sig = K.demod(s, _sqsum(weight))
    wp = K.weight_pack(weight.detach(), c, x.dtype, cs=s, nb=n)
    y = K.conv_fwd(x, wp, cout, 3, 3, pad, x_halo=x_halo, y_halo=y_halo, row_scale=sig, act=act,
                   residual=residual, per_sample=True)
    return y, sig, wp


def _modconv_bwd(gy, P, x, s, sig, weight, pad, x_halo, gadd, need_x, *, wfwd=None,
                 relu_mask=False):
    """gy: gradient w.r.t. the pre-activation conv output.  P[b,o] = sum_hw gy*y is either
    given (computed by an otm_mod_out pass) or, when P is None, produced by the wgrad kernel's
    epilogue from the per-sample forward pack `wfwd`.  relu_mask: x is a ReLU output and the
    returned gx is masked by (x > 0), i.e. it is already the gradient of the ReLU's input.
    Returns (gx, ds, dw)."""
    n, cin, h, w = x.shape
    c = eq_scale(weight)
    dw, dw_ret = _wgrad_buffer(weight)
    if P is None:
        P = torch.zeros((n, weight.shape[0]), dtype=torch.float32, device=x.device)
        K.conv_wgrad(x, gy, dw, 3, 3, pad, x_halo=x_halo, alpha=c, rs=sig, cs=s, wfwd=wfwd, P=P)
    else:
        K.conv_wgrad(x, gy, dw, 3, 3, pad, x_halo=x_halo, alpha=c, rs=sig, cs=s)
    # dgrad with the un-modulated, demodulated weights: gxt = d L / d (s*x)
    wpt = K.weight_pack(weight.detach(), c, gy.dtype, rs=sig, nb=n, transpose=True)
    gxt_p = K.conv_fwd(gy, wpt, cin, 3, 3, 2 - pad + x_halo, per_sample=True)
    gint = gxt_p[:, :, x_halo : x_halo + h, x_halo : x_halo + w] if x_halo else gxt_p
    gx, Q = K.mod_in(gint, x, s, g_halo=x_halo, gadd=gadd, relu_mask=relu_mask)
    ds = K.mod_bwd(weight.detach(), c, s, sig, _sqsum(weight), P, Q, dw)
    return (gx if need_x else None), ds, dw_ret


class ModConvFn(torch.autograd.Function):
    """Conv2dWeightModulate (+ ReLU) with zero padding 1 (the up-sampling stages,
    reference builder.py:190-197)."""

    @staticmethod
    def forward(ctx, x, s, weight, act, y_halo, pad):
        s = s.contiguous().float()
        y, sig, _ = _modconv_fwd(x, s, weight, pad, 0, act, None, y_halo)
        ctx.act, ctx.pad = act, pad
        ctx.save_for_backward(x, s, weight, sig, y)
        return y

    @staticmethod
    def backward(ctx, g):
        x, s, weight, sig, y = ctx.saved_tensors
        g = _g(g, y)
        gy, P = K.mod_out(g, y, act=ctx.act, materialise=ctx.act != ACT_NONE)
        if gy is None:
            gy = g
        gx, ds, dw = _modconv_bwd(gy, P, x, s, sig, weight, ctx.pad, 0, None,
                                  ctx.needs_input_grad[0])
        return gx, ds, dw, None, None, None


def mod_conv(x, s, weight, *, act=ACT_RELU, y_halo=0, pad=1):
    return ModConvFn.apply(x, s, weight, act, y_halo, pad)


class UpModConvFn(torch.autograd.Function):
    """UpSample -> Conv2dWeightModulate(zero pad 1) -> activation (reference builder.py:190-197)
    with SHARED conv weights: the up-sampling stencil writes x~ = s[b,i] * up(z) (it is per
    channel, so the style scale commutes), the conv reads the one shared pack and applies
    sigma_inv as its row scale.  Backward: the activation/P pass stores sigma_inv * gy, so dgrad
    and wgrad are shared-weight too; Q is reduced against x~ and un-scaled in otm_mod_bwd."""

    @staticmethod
    def forward(ctx, z, s, weight, act, y_halo):
        s = s.contiguous().float()
        cout = weight.shape[0]
        sig = K.demod(s, _sqsum(weight))
        xt = K.up(z, scale=s)
        y = K.conv_fwd(xt, _pack(weight, xt.dtype, False), cout, 3, 3, 1, y_halo=y_halo, row_scale=sig,
                       act=act)
        ctx.act = act
        ctx.save_for_backward(xt, s, weight, sig, y)
        return y

    @staticmethod
    def backward(ctx, g):
        xt, s, weight, sig, y = ctx.saved_tensors
        c = eq_scale(weight)
        g = _g(g, y)
        gu, P = K.mod_out(g, y, act=ctx.act, gy_scale=sig)  # gu = sigma_inv * act'(y) * g
        dw, dw_ret = _wgrad_buffer(weight)
        K.conv_wgrad(xt, gu, dw, 3, 3, 1, alpha=c)
        # Qt = sum_hw gxt * x~ (= s * Q): from the dgrad's epilogue where the launch can
        fused = K.conv_fwd_dot(gu, _pack(weight, gu.dtype, True), weight.shape[1], 1, xt)
        if fused is not None:
            gxt, Qt = fused
        else:
            gxt = K.conv_fwd(gu, _pack(weight, gu.dtype, True), weight.shape[1], 3, 3, 1)
            _, Qt = K.mod_in(gxt, xt, s, want_gx=False)
        ds = K.mod_bwd(weight.detach(), c, s, sig, _sqsum(weight), P, Qt, dw, q_scaled=True)
        gz = K.up_bwd(gxt, scale=s) if ctx.needs_input_grad[0] else None
        return gz, ds, dw_ret, None, None


def up_mod_conv(z, s, weight, *, act=ACT_RELU, y_halo=0):
    return UpModConvFn.apply(z, s, weight, act, y_halo)


class ModResBlockFn(torch.autograd.Function):
    """ModulatedResnetBlock (reference blocks.py:36-68):
    x + modconv2(refl(ReLU(modconv1(refl(x), w))), w); x carries a reflect halo of 1.

    conv1 runs on per-sample packs cW * s1 (its input is the un-modulated residual stream) and
    writes h~ = s2 * ReLU(sigma1 * u1): conv2's modulation is conv1's post-activation scale, so
    conv2 reads ONE shared weight pack.  Backward needs two HBM passes besides the four GEMMs:
      * conv2's input-side pass (otm_mod_in) folds dgrad2's halo, applies s2, the ReLU mask of
        h~ and sigma1, i.e. it hands conv1 the gradient w.r.t. its RAW output u1, so conv1's
        dgrad and wgrad use the shared pack; its reduction Q~2 = sum fold * h~ is at the same
        time conv1's demodulation term P1 (sum dy1 * y1 = sum s2 * fold * ReLU(y1)) and, divided
        by s2 in otm_mod_bwd, conv2's direct style term Q2;
      * conv1's input-side pass adds the residual gradient.
    conv2's demodulation term P2 comes out of its wgrad epilogue (tcgen05 path)."""

    @staticmethod
    def forward(ctx, x, s1, s2, w1, w2, y_halo):
        s1 = s1.contiguous().float()
        s2 = s2.contiguous().float()
        n, f = x.shape[0], w1.shape[0]
        sig1 = K.demod(s1, _sqsum(w1))
        sig2 = K.demod(s2, _sqsum(w2))
        wp1 = K.weight_pack(w1.detach(), eq_scale(w1), x.dtype, cs=s1, nb=n)
        ht = K.conv_fwd(x, wp1, f, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig1, act=ACT_RELU,
                        post_scale=s2, per_sample=True)
        out = K.conv_fwd(ht, _pack(w2, x.dtype, False), f, 3, 3, 1, x_halo=1, y_halo=y_halo,
                         row_scale=sig2, residual=x)
        ctx.fused = K.wgrad_fuses_P(ht, out, 3, 3, 1, 1)
        ctx.save_for_backward(x, s1, s2, w1, w2, ht, sig1, sig2, None if ctx.fused else out)
        return out

    @staticmethod
    def backward(ctx, g):
        x, s1, s2, w1, w2, ht, sig1, sig2, out = ctx.saved_tensors
        n, f, h, w = x.shape
        c1, c2 = eq_scale(w1), eq_scale(w2)
        g = _g(g, x)
        # ---- conv2: dy2 = g, input h~ (already modulated) ------------------------------------
        dw2, dw2_ret = _wgrad_buffer(w2)
        if ctx.fused:
            P2 = torch.zeros((n, f), dtype=torch.float32, device=x.device)
            K.conv_wgrad(ht, g, dw2, 3, 3, 1, x_halo=1, alpha=c2, rs=sig2,
                         wfwd=_pack(w2, ht.dtype, False), P=P2, wfwd_per_sample=False)
        else:
            _, P2 = K.mod_out(g, out, res=x, materialise=False)
            K.conv_wgrad(ht, g, dw2, 3, 3, 1, x_halo=1, alpha=c2, rs=sig2)
        wpt2 = K.weight_pack(w2.detach(), c2, g.dtype, rs=sig2, nb=n, transpose=True)
        direct = K.dgrad_reflect_ok(g, f)
        if direct and K.dgrad_reflect_fuses_gate(g, wpt2, f, per_sample=True):
            # the input-side pass runs in the dgrad's epilogue: gate by ReLU(h~), s2 * sigma1 as
            # the row / post scale, Q~2 reduced against the raw accumulator
            gu1, Qt2 = K.conv_dgrad_reflect(g, wpt2, f, per_sample=True, gate=ht, row_scale=s2,
                                            post_scale=sig1, want_dot=True)
        elif direct:
            ght = K.conv_dgrad_reflect(g, wpt2, f, per_sample=True)
            gu1, Qt2 = K.mod_in(ght, ht, s2, relu_mask=True, gx_scale=sig1)
        else:
            ght_p = K.conv_fwd(g, wpt2, f, 3, 3, 2, per_sample=True)
            gu1, Qt2 = K.mod_in(ght_p[:, :, 1 : 1 + h, 1 : 1 + w], ht, s2, g_halo=1, relu_mask=True,
                                gx_scale=sig1)
        ds2 = K.mod_bwd(w2.detach(), c2, s2, sig2, _sqsum(w2), P2, Qt2, dw2, q_scaled=True)
        # ---- conv1: dy = gu1 (w.r.t. its raw output), P1 == Qt2 --------------------------------
        dw1, dw1_ret = _wgrad_buffer(w1)
        if direct and K.wgrad_fuses_Q(x, gu1, 3, 3, 1, 1):
            # Q1 = sum_hw (dL/d(s1 x)) * x is a column reduction of the per-sample accumulator the
            # wgrad kernel already holds; the dgrad's epilogue applies s1 and adds the skip
            # gradient: conv1 needs no input-side pass at all
            Q1 = torch.zeros((n, f), dtype=torch.float32, device=x.device)
            K.conv_wgrad(x, gu1, dw1, 3, 3, 1, x_halo=1, alpha=c1, cs=s1,
                         wfwd=_pack(w1, x.dtype, False), Q=Q1, wfwd_per_sample=False)
            gx = (K.conv_dgrad_reflect(gu1, _pack(w1, g.dtype, True), f, residual=g, row_scale=s1)
                  if ctx.needs_input_grad[0] else None)
            ds1 = K.mod_bwd(w1.detach(), c1, s1, sig1, _sqsum(w1), Qt2, Q1, dw1)
            return gx, ds1, ds2, dw1_ret, dw2_ret, None
        K.conv_wgrad(x, gu1, dw1, 3, 3, 1, x_halo=1, alpha=c1, cs=s1)
        if direct:
            gxt = K.conv_dgrad_reflect(gu1, _pack(w1, g.dtype, True), f)
            gx, Q1 = K.mod_in(gxt, x, s1, gadd=g, want_gx=ctx.needs_input_grad[0])
        else:
            gxt_p = K.conv_fwd(gu1, _pack(w1, g.dtype, True), f, 3, 3, 2)
            gx, Q1 = K.mod_in(gxt_p[:, :, 1 : 1 + h, 1 : 1 + w], x, s1, g_halo=1, gadd=g,
                              want_gx=ctx.needs_input_grad[0])
        ds1 = K.mod_bwd(w1.detach(), c1, s1, sig1, _sqsum(w1), Qt2, Q1, dw1)
        return gx, ds1, ds2, dw1_ret, dw2_ret, None


def mod_res_block(x, s1, s2, w1, w2, *, y_halo):
    return ModResBlockFn.apply(x, s1, s2, w1, w2, y_halo)


# ---------------------------------------------------------------------------
# global average pool (StyleExtractor head)
# ---------------------------------------------------------------------------
class AvgPoolFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x):
        ctx.shape, ctx.dtype = tuple(x.shape), x.dtype
        return K.avgpool(x)

    @staticmethod
    def backward(ctx, g):
        return K.avgpool_bwd(g.contiguous().float(), ctx.shape, ctx.dtype)


def avgpool(x):
    return AvgPoolFn.apply(x)


# ---------------------------------------------------------------------------
# losses: forward writes the scalar and the backward seed in the same pass
# ---------------------------------------------------------------------------
# The training step sums its loss terms with unit coefficients (each term's lambda is folded
# into the kernel that writes the backward seed), so the upstream gradient is exactly 1 and
# the seeds can be returned as they are.  Set to False to use the loss functions generically.
UNIT_LOSS_GRADS = False


def _scaled(seed, g):
    """seed was computed for an upstream gradient of 1; rescale if autograd hands us another."""
    if UNIT_LOSS_GRADS or seed is None:
        return seed
    return seed * g.to(seed.dtype)


class LsganFn(torch.autograd.Function):
    """mean((scores - target)^2) and mean(sign(2 s - 1)) (reference training.py:86,111-117)."""

    @staticmethod
    def forward(ctx, scores, target, weight):
        out, seed = K.loss_lsgan(scores, target, weight, want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(seed)
        conf = out[1:2].clone()
        raw = out[0:1].clone()
        ctx.mark_non_differentiable(conf, raw)
        return out[0:1] * weight, conf, raw

    @staticmethod
    def backward(ctx, g, _c, _r):
        (seed,) = ctx.saved_tensors
        return _scaled(seed, g), None, None


def lsgan(scores, target, weight=1.0):
    """Returns (weight * loss, confidence, loss) as 1-element tensors."""
    return LsganFn.apply(nhwc(scores), float(target), float(weight))


class L1Fn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, a, b, weight):
        out, seed = K.loss_l1(a, b, weight, want_grad=ctx.needs_input_grad[0])
        ctx.save_for_backward(seed)
        ctx.mark_non_differentiable(out)
        return out * weight, out

    @staticmethod
    def backward(ctx, g, _):
        (seed,) = ctx.saved_tensors
        return _scaled(seed, g), None, None


def l1(a, b, weight=1.0):
    """Returns (weight * loss, loss): the raw value is what the reference logs."""
    return L1Fn.apply(nhwc(a), nhwc(b, a.dtype), float(weight))


class KlFn(torch.autograd.Function):
    """kl_loss_func (reference loss.py:82-92): mean^2 + (biased var - 1)^2 over all elements."""

    @staticmethod
    def forward(ctx, x, weight):
        mom = K.moments(x)
        n = float(x.numel())
        m = mom[0] / n
        v = mom[1] / n - m * m
        # d kl / dx_i = c0 + c1 x_i
        c1 = 4.0 * (v - 1.0) / n * weight
        c0 = (2.0 * m / n) * weight - c1 * m
        ctx.save_for_backward(x, torch.stack([c0, c1]))
        raw = (m * m + (v - 1.0) ** 2).reshape(1)
        ctx.mark_non_differentiable(raw)
        return raw * weight, raw

    @staticmethod
    def backward(ctx, g, _):
        x, coef = ctx.saved_tensors
        return K.affine_grad(x, (coef * g.reshape(())).contiguous()), None


def kl(x, weight=1.0):
    """Returns (weight * loss, loss)."""
    return KlFn.apply(x, float(weight))


class PathFn(torch.autograd.Function):
    """path_loss_func (reference loss.py:98-111).  Each feature tensor holds the two
    extractions of the same latent stacked along the batch: [f1 ; f2] of shape [2B, ...]."""

    @staticmethod
    def forward(ctx, h, weight, *feats):
        L = len(feats)
        b = feats[0].shape[0] // 2
        out = torch.zeros(1, dtype=torch.float32, device=h.device)
        seeds = []
        for idx, f in enumerate(feats):
            g = None
            if ctx.needs_input_grad[2 + idx]:
                n, c, hh, ww = f.shape
                g = K.alloc(n, c, hh, ww, f.dtype, f.device)
            K.loss_path(f[:b], f[b:], h, 1.0 / L, weight, out, want_grad=g is not None,
                        g1=None if g is None else g[:b], g2=None if g is None else g[b:])
            seeds.append(g)
        ctx.save_for_backward(*seeds)
        ctx.mark_non_differentiable(out)
        return out * weight, out

    @staticmethod
    def backward(ctx, g, _):
        return (None, None, *[_scaled(s, g) for s in ctx.saved_tensors])


def path(feats, h, weight=1.0):
    """feats: list of [2B,C,H,W] tensors ([f1 ; f2] stacked); h: [B] finite-difference steps.
    Returns (weight * loss, loss)."""
    return PathFn.apply(h.contiguous().float(), float(weight), *feats)


# ---------------------------------------------------------------------------
# small dense layers: EqualisedLinear, MappingNetwork, style-cycle loss (csrc/style.cu)
# ---------------------------------------------------------------------------
def _rows32(t: torch.Tensor) -> torch.Tensor:
    """fp32 [n, k] rows with unit last stride; a stride-0 batch expand is kept as it is."""
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() != 2:
        raise ValueError(f"expected [n, k], got {tuple(t.shape)}")
    if t.shape[1] > 1 and t.stride(1) != 1:
        t = t.contiguous()
    return t


class LinearsFn(torch.autograd.Function):
    """J EqualisedLinear layers (reference layers.py:27-43) in ONE launch each way:
    y_j = x_j @ (c_j W_j)^T + b_j.  Inputs may alias (the two `to_style` layers of a
    ModulatedResnetBlock read the same w[i], blocks.py:62-68): their gradients are summed."""

    @staticmethod
    def forward(ctx, n_jobs, *tensors):
        xs, ws, bs = tensors[:n_jobs], tensors[n_jobs : 2 * n_jobs], tensors[2 * n_jobs :]
        xs = [_rows32(x) for x in xs]
        ys = [torch.empty((x.shape[0], w.shape[0]), dtype=torch.float32, device=x.device)
              for x, w in zip(xs, ws)]
        K.linear_fwd([dict(x=x, w=w.detach(), bias=b.detach(), y=y)
                      for x, w, b, y in zip(xs, ws, bs, ys)])
        ctx.n_jobs = n_jobs
        ctx.params = (ws, bs)
        ctx.save_for_backward(*xs)
        return tuple(ys)

    @staticmethod
    def backward(ctx, *gys):
        J = ctx.n_jobs
        xs = ctx.saved_tensors
        ws, bs = ctx.params
        specs, gx, gw, gb = [], [None] * J, [None] * J, [None] * J
        # one zero-filled buffer for every input gradient (dx is accumulated atomically)
        need = [j for j in range(J) if gys[j] is not None and ctx.needs_input_grad[1 + j]]
        big = (torch.zeros(sum(xs[j].numel() if xs[j].stride(0) else xs[j].shape[0] * xs[j].shape[1]
                               for j in need), dtype=torch.float32, device=xs[0].device)
               if need else None)
        off = 0
        for j in range(J):
            if gys[j] is None:
                continue
            sp = dict(x=xs[j], w=ws[j].detach(), dy=gys[j].contiguous().float())
            if ctx.needs_input_grad[1 + J + j]:
                sp["dw"], gw[j] = _param_grad_buffer(ws[j])
            if ctx.needs_input_grad[1 + 2 * J + j]:
                sp["dbias"], gb[j] = _param_grad_buffer(bs[j])
            if ctx.needs_input_grad[1 + j]:
                n, k = xs[j].shape
                gx[j] = sp["dx"] = big[off : off + n * k].view(n, k)
                off += n * k
            specs.append(sp)
        if specs:
            K.linear_bwd(specs)
        return (None, *gx, *gw, *gb)


def linears(xs, layers):
    """layers: EqualisedLinear modules (attributes .weight.weight [o,k], .bias [o])."""
    return LinearsFn.apply(len(xs), *xs, *[m.weight.weight for m in layers], *[m.bias for m in layers])


def linear(x, layer):
    return LinearsFn.apply(1, x, layer.weight.weight, layer.bias)[0]


class MappingFn(torch.autograd.Function):
    """MappingNetwork.forward (reference builder.py:46-49) with, optionally, the style mixing and
    domain-variable interpolation of get_single_w / get_two_w fused (builder.py:51-132)."""

    @staticmethod
    def forward(ctx, z1, z2, cross, n_blocks, d0, d1, d_const, n_out, *params):
        L = len(params) // 2
        ws, bs = params[:L], params[L:]
        z1 = z1.contiguous().float()
        z2 = None if z2 is None else z2.contiguous().float()
        d = tuple(None if t is None else t.contiguous().float() for t in (d0, d1))
        outs = K.mapping_fwd(z1, z2, cross, [w.detach() for w in ws], [b.detach() for b in bs],
                             n_blocks, d, d_const, n_out)
        ctx.cfg = (cross, n_blocks, d, d_const, z2 is not None)
        ctx.params = (ws, bs)
        ctx.save_for_backward(z1, *([z2] if z2 is not None else []))
        return tuple(outs)

    @staticmethod
    def backward(ctx, *gouts):
        cross, n_blocks, d, d_const, has_z2 = ctx.cfg
        ws, bs = ctx.params
        z1 = ctx.saved_tensors[0]
        z2 = ctx.saved_tensors[1] if has_z2 else None
        bufs_w, bufs_b, ret_w, ret_b = [], [], [], []
        for w, b in zip(ws, bs):
            bw, rw = _param_grad_buffer(w)
            bb, rb = _param_grad_buffer(b)
            bufs_w.append(bw), ret_w.append(rw), bufs_b.append(bb), ret_b.append(rb)
        douts = [None if g is None else g.contiguous().float() for g in gouts]
        douts += [None] * (2 - len(douts))
        if any(g is not None for g in douts):
            K.mapping_bwd(z1, z2, cross, [w.detach() for w in ws], [b.detach() for b in bs], n_blocks,
                          d, d_const, douts, bufs_w, bufs_b)
        return (None,) * 8 + (*ret_w, *ret_b)


def mapping(net_layers, z1, z2=None, cross=None, *, n_blocks=1, d=(None, None), d_const=(1.0, 1.0),
            n_out=1):
    """net_layers: the EqualisedLinear modules of MappingNetwork.net.  Returns n_out tensors
    [n_blocks, batch, features]."""
    if any(t is not None and t.requires_grad for t in (z1, z2)):
        raise ValueError("the mapping kernel does not differentiate w.r.t. z (the reference never does)")
    return MappingFn.apply(z1, z2, cross, n_blocks, d[0], d[1], tuple(d_const), n_out,
                           *[m.weight.weight for m in net_layers], *[m.bias for m in net_layers])


class StyleCycleFn(torch.autograd.Function):
    """style_cycle_loss_func (reference loss.py:60-75): scalar and both backward seeds in one
    single-CTA launch."""

    @staticmethod
    def forward(ctx, a, b, weight, ratio):
        a, b = _rows32(a), _rows32(b)
        out, da, db = K.loss_style_cycle(a, b, ratio, weight, want_grad=tuple(ctx.needs_input_grad[:2]))
        ctx.save_for_backward(da, db)
        ctx.mark_non_differentiable(out)
        return out * weight, out

    @staticmethod
    def backward(ctx, g, _):
        da, db = ctx.saved_tensors
        return _scaled(da, g), _scaled(db, g), None, None


def style_cycle(original_w, reconstructed_w, weight=1.0, cos_l2_ratio=0.2):
    """Returns (weight * loss, loss) as 1-element tensors."""
    return StyleCycleFn.apply(original_w, reconstructed_w, float(weight), float(cos_l2_ratio))
