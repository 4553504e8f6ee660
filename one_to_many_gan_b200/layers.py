"""Primitive layers with the constructors, attribute names and state_dict keys of the
reference's src/model/layers.py, executing on the otm_b200 kernels.

The networks in builder.py never call these layer-by-layer on their hot path (they run
fused blocks from ops.py that read the parameters held here); the layer `forward`s exist so
that a user who composes layers by hand, as the reference allows, gets the same semantics."""

from __future__ import annotations

import math

import torch
from torch import nn

from . import ops


class EqualisedWeight(nn.Module):
    """Learning-rate equalised weight (reference layers.py:12-24): raw N(0,1) parameter,
    scale c = 1/sqrt(fan_in) applied at use (folded into the kernels' weight staging)."""

    def __init__(self, shape: list[int]):
        super().__init__()
        self.c = 1 / math.sqrt(math.prod(shape[1:]))
        self.weight = nn.Parameter(torch.randn(shape))

    def forward(self):
        return self.weight * self.c


class EqualisedLinear(nn.Module):
    """Reference layers.py:27-43.  K is 6 or 512 here: far below a tensor-core tile."""

    def __init__(self, in_features: int, out_features: int, bias: float = 0.0):
        super().__init__()
        self.in_features = in_features
        self.out_features = out_features
        self.weight = EqualisedWeight([out_features, in_features])
        self.bias = nn.Parameter(torch.zeros(out_features) + bias)

    def forward(self, x: torch.Tensor):
        lead = x.shape[:-1]
        y = ops.linear(x.reshape(-1, self.in_features), self)
        return y.reshape(*lead, self.out_features)

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}"


class EqualisedConv2d(nn.Module):
    """Reference layers.py:46-102 (stride 1, dilation 1 — the only values the reference uses)."""

    def __init__(self, in_features: int, out_features: int, kernel_size, stride: int = 1,
                 padding: int = 0, dilation: int = 1, *, use_bias: bool = True):
        super().__init__()
        if stride != 1 or dilation != 1:
            raise ValueError("the B200 path implements stride=1, dilation=1 (all the reference uses)")
        if not isinstance(kernel_size, int):
            if kernel_size[0] != kernel_size[1]:
                raise ValueError("square kernels only")
            kernel_size = kernel_size[0]
        self.in_features = in_features
        self.out_features = out_features
        self.kernel_size = kernel_size
        self.stride = stride
        self.padding = padding
        self.dilation = dilation
        self.weight = EqualisedWeight([out_features, in_features, kernel_size, kernel_size])
        self.use_bias = use_bias
        if use_bias:
            self.bias = nn.Parameter(torch.zeros(out_features))

    def forward(self, x: torch.Tensor):
        x = ops.nhwc(x)
        return ops.conv(x, self.weight.weight, self.bias if self.use_bias else None,
                        self.kernel_size, self.padding)

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features},"
                f" kernel_size={self.kernel_size}, stride={self.stride}, dilation={self.dilation}")


class Conv2dWeightModulate(nn.Module):
    """Reference layers.py:111-182 in its dense form: the per-sample style scales the input
    channels, the demodulation scales the output channels, the convolution weight is shared."""

    def __init__(self, in_features: int, out_features: int, kernel_size: int, w_dim: int,
                 padding: int, *, use_bias: bool = False, demodulate: bool = True,
                 eps: float = 1e-8):
        super().__init__()
        if kernel_size != 3 or not demodulate or use_bias or eps != 1e-8:
            raise ValueError("the B200 path implements the configuration the reference uses: "
                             "3x3, demodulated, no bias, eps 1e-8")
        self.in_features = in_features
        self.out_features = out_features
        self.demodulate = demodulate
        self.padding = padding
        self.weight = EqualisedWeight([out_features, in_features, kernel_size, kernel_size])
        self.eps = eps
        self.use_bias = use_bias
        self.to_style = EqualisedLinear(w_dim, in_features, bias=1)

    def forward(self, x: torch.Tensor, w: torch.Tensor):
        """x is what the reference passes: already reflect-padded when padding == 0."""
        x = ops.nhwc(x)
        return ops.mod_conv(x, self.to_style(w), self.weight.weight, act=ops.ACT_NONE,
                            pad=self.padding)

    def extra_repr(self):
        return (f"in_features={self.in_features}, out_features={self.out_features},"
                f"demodulate={self.demodulate}, padding={self.padding}, eps={self.eps}")


class Smooth(nn.Module):
    """Reference layers.py:191-214.  Only ever used inside UpSample/DownSample, where the
    blur is fused with the resampling stencil; the buffer keeps the state_dict identical."""

    def __init__(self):
        super().__init__()
        kernel = torch.tensor([[[[1.0, 2.0, 1.0], [2.0, 4.0, 2.0], [1.0, 2.0, 1.0]]]])
        self.register_buffer("kernel", kernel / kernel.sum())


class UpSample(nn.Module):
    """Reference layers.py:217-229."""

    def __init__(self):
        super().__init__()
        self.smooth = Smooth()

    def forward(self, x: torch.Tensor):
        return ops.up(ops.nhwc(x))


class DownSample(nn.Module):
    """Reference layers.py:232-247."""

    def __init__(self, *, smooth=True):
        super().__init__()
        if not smooth:
            raise ValueError("the B200 path implements smooth=True (all the reference uses)")
        self.smooth_map = smooth
        self.smooth = Smooth()

    def forward(self, x: torch.Tensor):
        return ops.down(ops.nhwc(x))
