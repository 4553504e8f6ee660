"""R1 gradient penalty on real images (BASELINE config 5: "R1 gradient penalty every step, double
backward through D").  An extension: the reference trains without it; the oracle
is `torch.autograd.grad(D(x).sum(), x, create_graph=True)` on the reference Discriminator
(oracle/reference_port.py `r1_penalty`).

    penalty = gamma / 2 * mean_b || d sum(D(x_b)) / d x_b ||^2

and its gradient w.r.t. D's weights needs the derivative OF a backward pass.  The convolution is
bilinear in (x, W), so {forward, dgrad, wgrad} is closed under differentiation: the three
autograd Functions below call the library's conv kernels (tcgen05 in bf16 mode) and express
their own backward through each other, to any order.  The per-sample normalisation, LeakyReLU
and the blur/bilinear stencils between the convs are evaluated with differentiable tensor ops on
this path only -- their second derivatives are needed (the InstanceNorm Jacobian depends on its
input); they are HBM-bound and a few percent of an iteration."""

from __future__ import annotations

import torch
import torch.nn.functional as F

from . import kernels as K
from . import ops


def _cl(t: torch.Tensor, dtype) -> torch.Tensor:
    return ops.nhwc(t.detach(), dtype)


class _Conv(torch.autograd.Function):
    """y = conv(x, w), stride 1, zero padding `pad`; w already carries the equalised-LR scale."""

    @staticmethod
    def forward(ctx, x, w, pad, dtype, out_dtype):
        cout, cin, k, _ = w.shape
        xin = _cl(x, torch.float32 if cin == 1 else dtype)
        wp = K.weight_pack(w.detach().float().contiguous(), 1.0, xin.dtype)
        y = K.conv_fwd(xin, wp, cout, k, k, pad, out_dtype=out_dtype)
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
        return gx, gw, None, None, None


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


def _eq_conv(x, layer, dtype, out_dtype=None):
    """EqualisedConv2d (reference layers.py:82-102) on the twice-differentiable path."""
    w = layer.weight.weight
    y = _Conv.apply(x, w * ops.eq_scale(w), layer.padding, dtype, out_dtype or dtype)
    if getattr(layer, "bias", None) is not None:
        y = y + layer.bias.view(1, -1, 1, 1).to(y.dtype)
    return y


def _instance_norm(x, eps=1e-5):
    xf = x.float()
    m = xf.mean(dim=(2, 3), keepdim=True)
    v = xf.var(dim=(2, 3), unbiased=False, keepdim=True)
    return ((xf - m) * torch.rsqrt(v + eps)).to(x.dtype)


def _down(x):
    """DownSample (reference layers.py:232-247): replicate-pad blur, bilinear to (H//2, W//2)."""
    xf = F.pad(x.float(), (1, 1, 1, 1), mode="replicate")
    h, w = x.shape[2:]
    rows = xf[:, :, 0:h] + 2.0 * xf[:, :, 1 : h + 1] + xf[:, :, 2 : h + 2]
    blur = (rows[..., 0:w] + 2.0 * rows[..., 1 : w + 1] + rows[..., 2 : w + 2]) * (1.0 / 16.0)
    y = F.interpolate(blur, size=(h // 2, w // 2), mode="bilinear", align_corners=False)
    return y.to(x.dtype)


def discriminator_scores(discriminator, x):
    """The Discriminator forward (reference builder.py:268-287) on the twice-differentiable path."""
    m, dt = discriminator.model, discriminator.act_dtype
    a = _down(F.leaky_relu(_eq_conv(x, m[0], dt), 0.2))
    for idx in (3, 7):
        a = _down(F.leaky_relu(_instance_norm(_eq_conv(a, m[idx], dt)), 0.2))
    a = F.leaky_relu(_instance_norm(_eq_conv(a, m[11], dt)), 0.2)
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
