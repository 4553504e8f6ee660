"""The four networks with the reference's constructors, forward signatures, attribute names
and state_dict keys (src/model/builder.py), executing as fused B200 stages.

`act_dtype` (torch.float32 = parity mode, torch.bfloat16 = tensor-core mode) is the storage
type of the activations between kernels; images in and out stay fp32 NCHW."""

from __future__ import annotations

import math
from typing import cast

import torch
from torch import nn

from . import ops
from .blocks import ModulatedResnetBlock, ResnetBlock, _ensure_halo
from .layers import Conv2dWeightModulate, DownSample, EqualisedConv2d, EqualisedLinear, UpSample


class MappingNetwork(nn.Module):
    """z -> style s (reference builder.py:16-132).  [B,6] tensors: host-launched torch ops;
    the RNG draw order of `_get_style_vector` is the reference's."""

    def __init__(self, features: int, n_layers: int, style_mixing_prob: float):
        super().__init__()
        if features > 32 or n_layers > 8:
            raise ValueError("the B200 mapping kernel implements w_dim <= 32 and <= 8 layers")
        self.d_latent = features
        self.style_mixing_prob = style_mixing_prob
        layers: list[nn.Module] = []
        for _ in range(n_layers):
            layers.append(EqualisedLinear(features, features))
            layers.append(nn.LeakyReLU(negative_slope=0.2, inplace=True))
        layers[-1] = nn.ReLU(inplace=True)  # theta = 0 must give the zero style
        self.net = nn.Sequential(*layers)
        self.register_buffer("shoeprint_style_vector",
                             torch.zeros((1, 1, features), dtype=torch.float), persistent=False)

    def linears(self):
        return [m for m in self.net if isinstance(m, EqualisedLinear)]

    def forward(self, z: torch.Tensor):
        """One fused kernel: F.normalize -> [Linear, LeakyReLU] ... Linear, ReLU (builder.py:46-49)."""
        return ops.mapping(self.linears(), z)[0][0]

    def get_two_w(self, batch_size, n_gen_blocks, device, domain_variables, *, mix_styles=True):
        d1, d2 = domain_variables
        s = self._get_style_vector(batch_size, n_gen_blocks, device, mix_styles=mix_styles)
        zero = cast(torch.Tensor, self.shoeprint_style_vector)
        return torch.lerp(zero, s, d1.view(1, -1, 1)), torch.lerp(zero, s, d2.view(1, -1, 1))

    def get_single_w(self, batch_size, n_gen_blocks, device, domain_variable, *, mix_styles=True):
        zero = cast(torch.Tensor, self.shoeprint_style_vector)
        if not isinstance(domain_variable, torch.Tensor) and domain_variable == 0:
            return zero.expand((n_gen_blocks, batch_size, self.d_latent))
        s = self._get_style_vector(batch_size, n_gen_blocks, device, mix_styles=mix_styles)
        if isinstance(domain_variable, torch.Tensor):
            d = domain_variable.view(1, -1, 1)
        else:
            d = torch.tensor(domain_variable, dtype=torch.float, device=device).view(1, 1, 1)
        return torch.lerp(zero, s, d)

    def _get_style_vector(self, batch_size, n_gen_blocks, device, *, mix_styles=True):
        if mix_styles and torch.rand(()).lt(self.style_mixing_prob):
            cross = int(torch.randint(0, n_gen_blocks, ()))
            z1 = torch.randn(batch_size, self.d_latent).to(device)
            z2 = torch.randn(batch_size, self.d_latent).to(device)
            s1 = self.forward(z1)[None].expand(cross, -1, -1)
            s2 = self.forward(z2)[None].expand(n_gen_blocks - cross, -1, -1)
            return torch.cat((s1, s2), dim=0)
        z = torch.randn(batch_size, self.d_latent).to(device)
        return self.forward(z)[None].expand(n_gen_blocks, -1, -1)


class Generator(nn.Module):
    """Reference builder.py:138-253."""

    def __init__(self, input_nc: int, w_dim: int, image_size, min_latent_resolution: int,
                 n_resnet_blocks: int, start_filters: int = 64, *, act_dtype=torch.float32):
        super().__init__()
        self.act_dtype = act_dtype
        filters = start_filters
        n_down = max(0, math.ceil(math.log2(min(image_size) / min_latent_resolution)))
        n_enc_res = n_resnet_blocks // 2
        n_dec_res = math.ceil(n_resnet_blocks / 2)
        self.n_down, self.n_enc_res, self.n_dec_res = n_down, n_enc_res, n_dec_res

        encoder: list[nn.Module] = [
            nn.ReflectionPad2d(3),
            EqualisedConv2d(input_nc, filters, kernel_size=7),
            nn.InstanceNorm2d(filters),
            nn.ReLU(inplace=True),
        ]
        for _ in range(n_down):
            encoder += [
                EqualisedConv2d(filters, filters * 2, kernel_size=3, padding=1),
                nn.InstanceNorm2d(filters * 2),
                nn.ReLU(inplace=True),
                DownSample(),
            ]
            filters *= 2
        encoder += [ResnetBlock(filters) for _ in range(n_enc_res)]
        self.encoder = nn.Sequential(*encoder)
        self.latent_channels = filters

        decoder: list[nn.Module] = [ModulatedResnetBlock(filters, w_dim=w_dim) for _ in range(n_dec_res)]
        for _ in range(n_down):
            decoder += [
                UpSample(),
                Conv2dWeightModulate(filters, filters // 2, kernel_size=3, padding=1, w_dim=w_dim),
                nn.ReLU(inplace=True),
            ]
            filters //= 2
        decoder += [nn.ReflectionPad2d(3), EqualisedConv2d(filters, input_nc, kernel_size=7), nn.Tanh()]
        self.decoder = nn.ModuleList(decoder)
        self.n_style_blocks = sum(
            [isinstance(m, ModulatedResnetBlock | Conv2dWeightModulate) for m in self.decoder]
        )
        self.latent_halo = 1 if n_dec_res > 0 else 0

    def latent_shape(self, image_shape):
        """Shape of encode(x) for an image batch of shape [B,C,H,W] (every DownSample floors)."""
        b, _, h, w = image_shape
        for _ in range(self.n_down):
            h, w = h // 2, w // 2
        return (b, self.latent_channels, h, w)

    # -- encoder ---------------------------------------------------------------------
    def encode(self, x: torch.Tensor):
        """Image [B,C,H,W] fp32 -> latent [B,F,h,w] (act_dtype, NHWC storage, reflect halo 1)."""
        enc = self.encoder
        x = ops.nhwc(x, torch.float32)
        xp = ops.norm_act(x, norm=False, y_halo=3)  # ReflectionPad2d(3)
        c0 = enc[1]
        raw = ops.conv(xp, c0.weight.weight, c0.bias, 7, 3, x_halo=3, out_dtype=self.act_dtype,
                       bias_dead=True)  # every encoder conv feeds an InstanceNorm
        n_tail = self.n_down + self.n_enc_res  # stages that follow the first conv

        def halo_after(stage_idx: int) -> int:
            # stage_idx counts stages after the first conv; what does the NEXT consumer need?
            if stage_idx < self.n_down:
                return 0  # next is a zero-padded 3x3 conv
            if stage_idx < n_tail:
                return 1  # next is a ResnetBlock
            return self.latent_halo  # the latent feeds the decoder

        a = ops.norm_act(raw, norm=True, act=ops.ACT_RELU, y_halo=halo_after(0))
        i = 4
        for d in range(self.n_down):
            c = enc[i]
            raw, st = ops.conv(a, c.weight.weight, c.bias, 3, 1, bias_dead=True, want_stats=True)
            a = ops.down(raw, norm=True, act=ops.ACT_RELU, y_halo=halo_after(d + 1), stats=st)
            i += 4
        for r in range(self.n_enc_res):
            blk = enc[i]
            a = ops.res_block(a, blk.conv_block[1].weight.weight, blk.conv_block[5].weight.weight,
                              y_halo=halo_after(self.n_down + r + 1))
            i += 1
        return ops.with_halo(a, halo_after(n_tail))

    # -- decoder ---------------------------------------------------------------------
    def _styles(self, w: torch.Tensor):
        """s = to_style(w[i]) of EVERY modulated conv of the decoder in one launch
        (reference layers.py:138-140,148; blocks.py:62-68: both convs of a block read w[i])."""
        dec = self.decoder
        layers, xs = [], []
        for r in range(self.n_dec_res):
            for c in (1, 4):
                layers.append(dec[r].conv_block[c].to_style)
                xs.append(w[r])
        for d in range(self.n_down):
            layers.append(dec[self.n_dec_res + 3 * d + 1].to_style)
            xs.append(w[self.n_dec_res + d])
        return ops.linears(xs, layers)

    def _decode(self, z: torch.Tensor, w: torch.Tensor, collect: bool):
        dec = self.decoder
        styles = self._styles(w) if self.n_style_blocks else ()
        z = ops.nhwc(z, self.act_dtype) if (z.dtype != self.act_dtype or ops.halo_of(z) == 0) else z
        feats = []
        n_styled = self.n_style_blocks
        i = 0
        j = 0
        for r in range(self.n_dec_res):
            blk = dec[j]
            z = _ensure_halo(z, 1)
            last_res = r == self.n_dec_res - 1
            nxt = 1 if not last_res else (0 if self.n_down > 0 else 3)
            c1, c2 = blk.conv_block[1], blk.conv_block[4]
            z = ops.mod_res_block(z, styles[2 * r], styles[2 * r + 1], c1.weight.weight,
                                  c2.weight.weight, y_halo=nxt)
            ops.with_halo(z, nxt)
            feats.append(z)
            i += 1
            j += 1
        for d in range(self.n_down):
            conv = dec[j + 1]
            last = d == self.n_down - 1
            i += 1
            if collect and i == n_styled:
                # the final styled layer is returned before its ReLU (reference builder.py:243-244)
                feats.append(ops.up_mod_conv(z, styles[2 * self.n_dec_res + d], conv.weight.weight,
                                             act=ops.ACT_NONE))
                return feats
            z = ops.up_mod_conv(z, styles[2 * self.n_dec_res + d], conv.weight.weight,
                                act=ops.ACT_RELU, y_halo=3 if last else 0)
            ops.with_halo(z, 3 if last else 0)
            # nn.ReLU(inplace=True) overwrites the tensor `extract` appended (builder.py:190-197,
            # 241): non-final up-sampling features reach the path loss post-ReLU.
            feats.append(z)
            j += 3
        if collect:
            return feats
        z = _ensure_halo(z, 3)
        cf = dec[j + 1]
        raw = ops.conv(z, cf.weight.weight, cf.bias, 7, 3, x_halo=3, out_dtype=torch.float32)
        return ops.norm_act(raw, norm=False, act=ops.ACT_TANH)

    def decode(self, z: torch.Tensor, w: torch.Tensor):
        """Latent + styles [n_style_blocks,B,w_dim] -> image [B,C,H,W] fp32."""
        return self._decode(z, w, collect=False)

    def extract(self, z: torch.Tensor, w: torch.Tensor):
        """Feature maps of every styled layer (reference builder.py:232-249)."""
        if self.n_style_blocks == 0:
            raise ValueError("No return layers specified.")
        return self._decode(z, w, collect=True)

    def forward(self, x: torch.Tensor, w: torch.Tensor):
        return self.decode(self.encode(x), w)


def _patch_trunk(model: nn.Sequential, x: torch.Tensor, act_dtype):
    """Shared PatchGAN trunk of Discriminator / StyleExtractor (reference builder.py:268-283)."""
    x = ops.nhwc(x, torch.float32)
    c = model[0]
    a = ops.conv(x, c.weight.weight, c.bias, 4, 1, act=ops.ACT_LRELU, out_dtype=act_dtype)
    a = ops.down(a)
    for idx in (3, 7):
        c = model[idx]
        raw, st = ops.conv(a, c.weight.weight, c.bias, 4, 1, bias_dead=True, want_stats=True)
        a = ops.down(raw, norm=True, act=ops.ACT_LRELU, stats=st)
    c = model[11]
    raw, st = ops.conv(a, c.weight.weight, c.bias, 4, 1, bias_dead=True, want_stats=True)
    return ops.norm_act(raw, norm=True, act=ops.ACT_LRELU, stats=st)


class Discriminator(nn.Module):
    """PatchGAN discriminator (reference builder.py:259-287); raw scores, fp32."""

    def __init__(self, input_nc: int, *, act_dtype=torch.float32):
        super().__init__()
        self.act_dtype = act_dtype
        self.model = nn.Sequential(
            EqualisedConv2d(input_nc, 64, kernel_size=4, padding=1),
            nn.LeakyReLU(0.2, inplace=True),
            DownSample(),
            EqualisedConv2d(64, 128, kernel_size=4, padding=1),
            nn.InstanceNorm2d(128),
            nn.LeakyReLU(0.2, inplace=True),
            DownSample(),
            EqualisedConv2d(128, 256, kernel_size=4, padding=1),
            nn.InstanceNorm2d(256),
            nn.LeakyReLU(0.2, inplace=True),
            DownSample(),
            EqualisedConv2d(256, 512, kernel_size=4, padding=1),
            nn.InstanceNorm2d(512),
            nn.LeakyReLU(negative_slope=0.2, inplace=True),
            EqualisedConv2d(512, 1, kernel_size=4, padding=1),
        )

    def forward(self, x: torch.Tensor):
        a = _patch_trunk(self.model, x, self.act_dtype)
        c = self.model[14]
        return ops.conv(a, c.weight.weight, c.bias, 4, 1, out_dtype=torch.float32)


class StyleExtractor(nn.Module):
    """Image -> style vector (reference builder.py:293-320)."""

    def __init__(self, input_nc: int = 1, w_dim: int = 8, *, act_dtype=torch.float32):
        super().__init__()
        self.act_dtype = act_dtype
        self.model = nn.Sequential(
            EqualisedConv2d(input_nc, 64, kernel_size=4, padding=1),
            nn.LeakyReLU(0.2, inplace=True),
            DownSample(),
            EqualisedConv2d(64, 128, kernel_size=4, padding=1),
            nn.InstanceNorm2d(128),
            nn.LeakyReLU(0.2, inplace=True),
            DownSample(),
            EqualisedConv2d(128, 256, kernel_size=4, padding=1),
            nn.InstanceNorm2d(256),
            nn.LeakyReLU(0.2, inplace=True),
            DownSample(),
            EqualisedConv2d(256, 512, kernel_size=4, padding=1),
            nn.InstanceNorm2d(512),
            nn.LeakyReLU(0.2, inplace=True),
            nn.AdaptiveAvgPool2d(1),
            nn.Flatten(),
            EqualisedLinear(512, w_dim),
        )

    def forward(self, x):
        a = _patch_trunk(self.model, x, self.act_dtype)
        return self.model[16](ops.avgpool(a))
