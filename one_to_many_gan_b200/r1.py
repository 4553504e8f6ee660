"""R1 gradient penalty on real images (BASELINE config 5: "R1 gradient penalty every step, double
backward through D").  An extension: the reference trains without it; the oracle
is `torch.autograd.grad(D(x).sum(), x, create_graph=True)` on the reference Discriminator
(oracle/reference_port.py `r1_penalty`).

    penalty = gamma / 2 * mean_b || d sum(D(x_b)) / d x_b ||^2

and its gradient w.r.t. D's weights needs the derivative OF a backward pass.  The convolution is
bilinear in (x, W), so {forward, dgrad, wgrad} is closed under differentiation: the three
autograd Functions below call the library's conv kernels (tcgen05 in bf16 mode) and express
their own backward through each other, to any order.  The per-sample normalisation, LeakyReLU
and the blur/bilinear stencils between the convs are twice-differentiable Functions over the
library's own passes as well: the stencils are linear (their backward is the transposed stencil,
whose backward is the stencil), LeakyReLU is piecewise linear, and the second derivative of
InstanceNorm -- its Jacobian depends on its input -- is the kernel pair otm_norm_act_bwd_bwd."""

from __future__ import annotations

import torch

from . import kernels as K
from . import ops


def _cl(t: torch.Tensor, dtype) -> torch.Tensor:
    return ops.nhwc(t.detach(), dtype)


class _Conv(torch.autograd.Function):
    """y = conv(x, w), stride 1, zero padding `pad`; w already carries the equalised-LR scale."""

    @staticmethod
    def forward(ctx, x, w, pad, dtype, out_dtype, bias=None):
        """bias: added in the conv epilogue; first-order gradient only (see _eq_conv)."""
        cout, cin, k, _ = w.shape
        xin = _cl(x, torch.float32 if cin == 1 else dtype)
        wp = K.weight_pack(w.detach().float().contiguous(), 1.0, xin.dtype)
        y = K.conv_fwd(xin, wp, cout, k, k, pad, out_dtype=out_dtype,
                       bias=None if bias is None else bias.detach())
        ctx.save_for_backward(x, w)
        ctx.cfg = (pad, dtype, out_dtype)
        return y

    @staticmethod
    def backward(ctx, g):
        x, w = ctx.saved_tensors
        pad, dtype, _ = ctx.cfg
        gx = gw = None
        if ctx.needs_input_grad[0]:
            gx = _Dgrad.apply(g, w, pad, dtype, x.dtype)
        if ctx.needs_input_grad[1]:
            gw = _Wgrad.apply(x, g, w.shape[2], pad, dtype)
        gb = None
        if ctx.needs_input_grad[5]:
            gb = K.channel_sum(_cl(g, g.dtype))
        return gx, gw, None, None, None, gb


class _Dgrad(torch.autograd.Function):
    """gx = conv_transpose(g, w): gradient of _Conv w.r.t. its input."""

    @staticmethod
    def forward(ctx, g, w, pad, dtype, out_dtype):
        cout, cin, k, _ = w.shape
        gin = _cl(g, torch.float32 if cout == 1 else dtype)
        wpt = K.weight_pack(w.detach().float().contiguous(), 1.0, gin.dtype, transpose=True)
        gx = K.conv_fwd(gin, wpt, cin, k, k, k - 1 - pad, out_dtype=out_dtype)
        ctx.save_for_backward(g, w)
        ctx.cfg = (pad, dtype, g.dtype)
        return gx

    @staticmethod
    def backward(ctx, gg):
        g, w = ctx.saved_tensors
        pad, dtype, g_dtype = ctx.cfg
        d_g = d_w = None
        if ctx.needs_input_grad[0]:
            d_g = _Conv.apply(gg, w, pad, dtype, g_dtype)
        if ctx.needs_input_grad[1]:  # <gg, dgrad(g, w)> = <conv(gg, w), g>
            d_w = _Wgrad.apply(gg, g, w.shape[2], pad, dtype)
        return d_g, d_w, None, None, None


class _Wgrad(torch.autograd.Function):
    """dw[o,i,r,s] = sum x[., i, .+r-pad, .+s-pad] g[., o, ., .]: gradient of _Conv w.r.t. w."""

    @staticmethod
    def forward(ctx, x, g, k, pad, dtype):
        cin, cout = x.shape[1], g.shape[1]
        xin = _cl(x, torch.float32 if cin == 1 else dtype)
        gin = _cl(g, torch.float32 if cout == 1 else dtype)
        dw = torch.zeros((cout, cin, k, k), dtype=torch.float32, device=x.device)
        K.conv_wgrad(xin, gin, dw, k, k, pad, alpha=1.0)
        ctx.save_for_backward(x, g)
        ctx.cfg = (pad, dtype)
        return dw

    @staticmethod
    def backward(ctx, gw):
        x, g = ctx.saved_tensors
        pad, dtype = ctx.cfg
        d_x = d_g = None
        if ctx.needs_input_grad[0]:
            d_x = _Dgrad.apply(g, gw, pad, dtype, x.dtype)
        if ctx.needs_input_grad[1]:
            d_g = _Conv.apply(x, gw, pad, dtype, g.dtype)
        return d_x, d_g, None, None, None


class _NormAct(torch.autograd.Function):
    """y = act(InstanceNorm(x)) (norm=False: y = act(x)), act piecewise linear -- twice
    differentiable on the library's kernels: the first backward is otm_norm_act_bwd, the second
    otm_norm_act_bwd_bwd (the InstanceNorm Jacobian depends on x; the activation's does not)."""

    @staticmethod
    def forward(ctx, x, norm, act, dtype):
        xin = _cl(x, dtype)
        stats = K.instnorm_stats(xin) if norm else None
        y = K.norm_act(xin, stats, act)
        ctx.act, ctx.dtype = act, dtype
        ctx.save_for_backward(x, stats)  # x itself: the second derivative flows back into it
        return y

    @staticmethod
    def backward(ctx, g):
        x, stats = ctx.saved_tensors
        return _NormActBwd.apply(g, x, stats, ctx.act, ctx.dtype), None, None, None


class _NormActBwd(torch.autograd.Function):
    """gx = B(g * act'(yh)); `stats` (mean, rstd of x) is auxiliary: the x-derivative computed in
    backward() already contains the dependence through the statistics."""

    @staticmethod
    def forward(ctx, g, x, stats, act, dtype):
        xin = _cl(x, dtype)
        gin = _cl(g, xin.dtype)
        gx = K.norm_act_bwd(gin, xin, stats, act)
        ctx.act = act
        ctx.save_for_backward(gin, xin, stats)
        return gx

    @staticmethod
    def backward(ctx, gg):
        gin, xin, stats = ctx.saved_tensors
        ggin = _cl(gg, xin.dtype)
        if stats is None:  # activation only: linear in g, no dependence on x almost everywhere
            return K.norm_act_bwd(ggin, xin, None, ctx.act), None, None, None, None
        d_g, d_x = K.norm_act_bwd_bwd(gin, ggin, xin, stats, ctx.act,
                                      want_dg=ctx.needs_input_grad[0], want_dx=ctx.needs_input_grad[1])
        return d_g, d_x, None, None, None


class _Down(torch.autograd.Function):
    """DownSample (reference layers.py:232-247): linear, so its backward is the transposed stencil
    and the backward of THAT is the stencil again."""

    @staticmethod
    def forward(ctx, x):
        ctx.hw = tuple(x.shape[2:])
        return K.down(ops.nhwc(x.detach()))

    @staticmethod
    def backward(ctx, g):
        return _DownT.apply(g, ctx.hw)


class _DownT(torch.autograd.Function):
    @staticmethod
    def forward(ctx, g, hw):
        return K.down_bwd(ops.nhwc(g.detach()), hw)

    @staticmethod
    def backward(ctx, gg):
        return _Down.apply(gg), None


def _eq_conv(x, layer, dtype, out_dtype=None, with_bias=False):
    """EqualisedConv2d (reference layers.py:82-102) on the twice-differentiable path.  Only layer
    0 takes its bias (with_bias): the biases of the three layers that feed an InstanceNorm are
    cancelled by it, and `grad_x sum(D(x))` does not depend on the last layer's.  Layer 0's bias
    DOES matter: it shifts the activations every later InstanceNorm Jacobian is evaluated at."""
    w = layer.weight.weight
    return _Conv.apply(x, w * ops.eq_scale(w), layer.padding, dtype, out_dtype or dtype,
                       layer.bias if with_bias else None)


def discriminator_scores(discriminator, x):
    """The Discriminator forward (reference builder.py:268-287) on the twice-differentiable path."""
    m, dt = discriminator.model, discriminator.act_dtype
    a = _eq_conv(x, m[0], dt, with_bias=True)
    a = _Down.apply(_NormAct.apply(a, False, ops.ACT_LRELU, dt))
    for idx in (3, 7):
        a = _Down.apply(_NormAct.apply(_eq_conv(a, m[idx], dt), True, ops.ACT_LRELU, dt))
    a = _NormAct.apply(_eq_conv(a, m[11], dt), True, ops.ACT_LRELU, dt)
    return _eq_conv(a, m[14], dt, torch.float32)


def r1_penalty(discriminator, real, gamma: float):
    """gamma/2 * mean_b ||grad_x sum D(x_b)||^2 as a 1-element tensor that is differentiable
    w.r.t. the discriminator's parameters."""
    x = real.detach().float().requires_grad_(True)
    with torch.enable_grad():
        scores = discriminator_scores(discriminator, x)
        (gx,) = torch.autograd.grad(scores.sum(), x, create_graph=True)
        pen = gx.float().square().sum(dim=(1, 2, 3)).mean() * (0.5 * gamma)
    return pen.reshape(1)
