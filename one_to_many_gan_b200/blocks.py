"""Residual blocks with the constructors and state_dict keys of the reference's
src/model/blocks.py; each block executes as ONE fused autograd function (ops.py)."""

from __future__ import annotations

import torch
from torch import nn

from . import ops
from .layers import Conv2dWeightModulate, EqualisedConv2d


def _ensure_halo(x: torch.Tensor, p: int) -> torch.Tensor:
    """Materialise the reflect halo the block needs unless the producer already wrote it."""
    if ops.halo_of(x) >= p:
        return x
    return ops.with_halo(ops.norm_act(ops.nhwc(x), norm=False, y_halo=p), p)


class ResnetBlock(nn.Module):
    """x + IN(conv3(refl(ReLU(IN(conv3(refl(x))))))) (reference blocks.py:9-33)."""

    def __init__(self, dim: int, *, use_bias: bool = False):
        super().__init__()
        if use_bias:
            raise ValueError("the B200 path implements use_bias=False (all the reference uses)")
        # same member layout as the reference so the state_dict keys are identical
        self.conv_block = nn.Sequential(
            nn.ReflectionPad2d(1),
            EqualisedConv2d(dim, dim, kernel_size=3, padding=0, use_bias=False),
            nn.InstanceNorm2d(dim),
            nn.ReLU(inplace=True),
            nn.ReflectionPad2d(1),
            EqualisedConv2d(dim, dim, kernel_size=3, padding=0, use_bias=False),
            nn.InstanceNorm2d(dim),
        )
        self.out_halo = 0

    def forward(self, x: torch.Tensor):
        x = _ensure_halo(x, 1)
        y = ops.res_block(x, self.conv_block[1].weight.weight, self.conv_block[5].weight.weight,
                          y_halo=self.out_halo)
        return ops.with_halo(y, self.out_halo)


class ModulatedResnetBlock(nn.Module):
    """x + modconv(refl(ReLU(modconv(refl(x), w))), w) (reference blocks.py:36-68)."""

    def __init__(self, dim: int, w_dim: int, *, use_bias: bool = False):
        super().__init__()
        self.conv_block = nn.ModuleList([
            nn.ReflectionPad2d(1),
            Conv2dWeightModulate(dim, dim, w_dim=w_dim, kernel_size=3, padding=0, use_bias=use_bias),
            nn.ReLU(inplace=True),
            nn.ReflectionPad2d(1),
            Conv2dWeightModulate(dim, dim, w_dim=w_dim, kernel_size=3, padding=0, use_bias=use_bias),
        ])
        self.out_halo = 0

    def forward(self, x: torch.Tensor, w: torch.Tensor):
        x = _ensure_halo(x, 1)
        c1, c2 = self.conv_block[1], self.conv_block[4]
        s1, s2 = ops.linears([w, w], [c1.to_style, c2.to_style])
        y = ops.mod_res_block(x, s1, s2, c1.weight.weight, c2.weight.weight, y_halo=self.out_halo)
        return ops.with_halo(y, self.out_halo)
