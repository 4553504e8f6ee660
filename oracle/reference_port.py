"""CPU oracle for the G+D training step of struan-robertson/one-to-many-gan.

TEST INFRASTRUCTURE ONLY.  Nothing in the product package
(`one_to_many_gan_b200/`, `train.py`) may import this module; only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
legs do, and there only as the checker / the CPU baseline.

This is an independent, functional restatement (plain torch on the CPU, fp32 or
fp64, autograd for the gradients) of the reference algorithm.  Every function
cites the reference file:line it follows.  Parameters live in flat dicts keyed
by the reference's own ``state_dict`` names so weights interchange with the
reference modules and with the product's facades.

Pinning: `tests/golden/make_golden.py` imports the real reference from
`/root/reference` in the build container and records losses, outputs and
gradient fingerprints for fixed seeds; `tests/test_oracle_golden.py` checks this
module against those fixtures (parity PINNED by reference-generated vectors; the
reference ships no tests or golden vectors of its own).  The one unpinned piece
is pytorch-ada (not vendored by the reference, not installed): it is replaced by
the identity, which is exact while ADA p == 0 (reference loss.py:22,28,33).
"""

from __future__ import annotations

import math
import random
from dataclasses import dataclass, field

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# ---------------------------------------------------------------------------
# Architecture description (reference src/model/builder.py)
# ---------------------------------------------------------------------------


@dataclass
class Arch:
    """Hyper-parameters that fix every tensor shape (config.toml:36-45)."""

    image_size: tuple[int, int] = (512, 256)
    image_channels: int = 1
    w_dim: int = 6
    min_latent_resolution: int = 64
    n_resnet_blocks: int = 7
    mapping_network_layers: int = 2
    start_filters: int = 64
    style_mixing_prob: float = 0.9

    @property
    def n_down(self) -> int:
        # builder.py:153-156
        return max(
            0,
            math.ceil(math.log2(min(self.image_size) / self.min_latent_resolution)),
        )

    @property
    def n_enc_res(self) -> int:
        return self.n_resnet_blocks // 2  # builder.py:157

    @property
    def n_dec_res(self) -> int:
        return math.ceil(self.n_resnet_blocks / 2)  # builder.py:158

    @property
    def n_style_blocks(self) -> int:
        return self.n_dec_res + self.n_down  # builder.py:209-214


def _randn_weight(shape):
    # layers.py:21  raw N(0,1) parameter; equalised scale applied at use.
    return torch.randn(shape)


def _smooth_kernel():
    # layers.py:197-203
    k = torch.tensor([[[[1.0, 2.0, 1.0], [2.0, 4.0, 2.0], [1.0, 2.0, 1.0]]]])
    return k / k.sum()


def init_discriminator(arch: Arch) -> dict[str, Tensor]:
    """builder.py:268-284 — construction order fixes the RNG draw order."""
    p: dict[str, Tensor] = {}
    chans = [arch.image_channels, 64, 128, 256, 512]
    idx = [0, 3, 7, 11]
    for j, i in enumerate(idx):
        p[f"model.{i}.weight.weight"] = _randn_weight([chans[j + 1], chans[j], 4, 4])
        p[f"model.{i}.bias"] = torch.zeros(chans[j + 1])
    for i in (2, 6, 10):
        p[f"model.{i}.smooth.kernel"] = _smooth_kernel()
    p["model.14.weight.weight"] = _randn_weight([1, 512, 4, 4])
    p["model.14.bias"] = torch.zeros(1)
    return p


def init_style_extractor(arch: Arch) -> dict[str, Tensor]:
    """builder.py:299-317."""
    p: dict[str, Tensor] = {}
    chans = [arch.image_channels, 64, 128, 256, 512]
    for j, i in enumerate([0, 3, 7, 11]):
        p[f"model.{i}.weight.weight"] = _randn_weight([chans[j + 1], chans[j], 4, 4])
        p[f"model.{i}.bias"] = torch.zeros(chans[j + 1])
    for i in (2, 6, 10):
        p[f"model.{i}.smooth.kernel"] = _smooth_kernel()
    p["model.16.weight.weight"] = _randn_weight([arch.w_dim, 512])
    p["model.16.bias"] = torch.zeros(arch.w_dim)
    return p


def init_mapping(arch: Arch) -> dict[str, Tensor]:
    """builder.py:25-38 — Linear at even indices of `net`."""
    p: dict[str, Tensor] = {}
    for layer in range(arch.mapping_network_layers):
        p[f"net.{2 * layer}.weight.weight"] = _randn_weight([arch.w_dim, arch.w_dim])
        p[f"net.{2 * layer}.bias"] = torch.zeros(arch.w_dim)
    return p


def generator_plan(arch: Arch):
    """Indices of the encoder Sequential / decoder ModuleList (builder.py:161-207)."""
    enc, dec = [], []
    f = arch.start_filters
    enc.append(("conv7", 1, arch.image_channels, f))  # index of the conv
    i = 4
    for _ in range(arch.n_down):
        enc.append(("down", i, f, 2 * f))  # conv at i, DownSample at i+3
        f *= 2
        i += 4
    for _ in range(arch.n_enc_res):
        enc.append(("res", i, f))
        i += 1
    j = 0
    for _ in range(arch.n_dec_res):
        dec.append(("modres", j, f))
        j += 1
    for _ in range(arch.n_down):
        dec.append(("up", j, f, f // 2))  # UpSample at j, modconv at j+1
        f //= 2
        j += 3
    dec.append(("conv7", j + 1, f, arch.image_channels))
    return enc, dec


def init_generator(arch: Arch) -> dict[str, Tensor]:
    """builder.py:161-207 in construction (= RNG draw) order."""
    p: dict[str, Tensor] = {}
    enc, dec = generator_plan(arch)
    for item in enc:
        if item[0] == "conv7":
            _, i, cin, cout = item
            p[f"encoder.{i}.weight.weight"] = _randn_weight([cout, cin, 7, 7])
            p[f"encoder.{i}.bias"] = torch.zeros(cout)
        elif item[0] == "down":
            _, i, cin, cout = item
            p[f"encoder.{i}.weight.weight"] = _randn_weight([cout, cin, 3, 3])
            p[f"encoder.{i}.bias"] = torch.zeros(cout)
            p[f"encoder.{i + 3}.smooth.kernel"] = _smooth_kernel()
        else:
            _, i, f = item
            for c in (1, 5):  # blocks.py:20-28, no bias
                p[f"encoder.{i}.conv_block.{c}.weight.weight"] = _randn_weight([f, f, 3, 3])
    for item in dec:
        if item[0] == "modres":
            _, j, f = item
            for c in (1, 4):  # blocks.py:48-59
                pre = f"decoder.{j}.conv_block.{c}"
                p[f"{pre}.weight.weight"] = _randn_weight([f, f, 3, 3])
                p[f"{pre}.to_style.weight.weight"] = _randn_weight([f, arch.w_dim])
                p[f"{pre}.to_style.bias"] = torch.ones(f)  # layers.py:138-140
        elif item[0] == "up":
            _, j, cin, cout = item
            p[f"decoder.{j}.smooth.kernel"] = _smooth_kernel()
            pre = f"decoder.{j + 1}"
            p[f"{pre}.weight.weight"] = _randn_weight([cout, cin, 3, 3])
            p[f"{pre}.to_style.weight.weight"] = _randn_weight([cin, arch.w_dim])
            p[f"{pre}.to_style.bias"] = torch.ones(cin)
        else:
            _, j, cin, cout = item
            p[f"decoder.{j}.weight.weight"] = _randn_weight([cout, cin, 7, 7])
            p[f"decoder.{j}.bias"] = torch.zeros(cout)
    return p


def is_buffer(name: str) -> bool:
    return name.endswith("smooth.kernel")


def init_all(arch: Arch, seed: int = 42):
    """train.py:35,72-90 — seed, then D, G, M, S in that order."""
    torch.manual_seed(seed)
    random.seed(seed)
    d = init_discriminator(arch)
    g = init_generator(arch)
    m = init_mapping(arch)
    s = init_style_extractor(arch)
    return {"D": d, "G": g, "M": m, "S": s}


# ---------------------------------------------------------------------------
# Primitive layers (reference src/model/layers.py)
# ---------------------------------------------------------------------------


def eq_scale(weight: Tensor) -> float:
    # layers.py:19  c = 1/sqrt(prod(shape[1:]))
    return 1.0 / math.sqrt(weight[0].numel())


def eq_conv2d(x, weight, bias=None, padding=0):
    # layers.py:82-102 with EqualisedWeight.forward layers.py:23-24
    return F.conv2d(x, weight * eq_scale(weight), bias=bias, padding=padding)


def eq_linear(x, weight, bias):
    # layers.py:39-40
    return F.linear(x, weight * eq_scale(weight), bias)


def modulated_conv2d(x, w, weight, style_weight, style_bias, padding, eps=1e-8):
    """layers.py:145-182: per-sample modulate, demodulate, grouped conv."""
    b, cin, h, wd = x.shape
    cout = weight.shape[0]
    s = eq_linear(w, style_weight, style_bias)  # [B, Cin]
    wts = (weight * eq_scale(weight))[None] * s[:, None, :, None, None]
    sigma_inv = torch.rsqrt((wts**2).sum(dim=(2, 3, 4), keepdim=True) + eps)
    wts = wts * sigma_inv
    y = F.conv2d(
        x.reshape(1, b * cin, h, wd),
        wts.reshape(b * cout, cin, *weight.shape[2:]),
        padding=padding,
        groups=b,
    )
    return y.reshape(b, cout, y.shape[2], y.shape[3])


def smooth(x, kernel):
    # layers.py:207-214: replicate pad 1 then depthwise 3x3 binomial blur
    b, c, h, w = x.shape
    y = F.pad(x.reshape(b * c, 1, h, w), (1, 1, 1, 1), mode="replicate")
    return F.conv2d(y, kernel.to(x.dtype)).reshape(b, c, h, w)


def up_sample(x, kernel):
    # layers.py:223-229
    y = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False)
    return smooth(y, kernel)


def down_sample(x, kernel):
    # layers.py:241-247
    y = smooth(x, kernel)
    return F.interpolate(
        y, (x.shape[2] // 2, x.shape[3] // 2), mode="bilinear", align_corners=False
    )


def inst_norm(x):
    # nn.InstanceNorm2d defaults: eps 1e-5, no affine, no running stats
    return F.instance_norm(x, eps=1e-5)


def refl(x, p):
    return F.pad(x, (p, p, p, p), mode="reflect")


# ---------------------------------------------------------------------------
# Networks (reference src/model/builder.py, blocks.py)
# ---------------------------------------------------------------------------


def generator_encode(P, x, arch: Arch):
    """builder.py:216-218 over the encoder built at :161-181."""
    enc, _ = generator_plan(arch)
    for item in enc:
        if item[0] == "conv7":
            i = item[1]
            x = eq_conv2d(refl(x, 3), P[f"encoder.{i}.weight.weight"], P[f"encoder.{i}.bias"])
            x = F.relu(inst_norm(x))
        elif item[0] == "down":
            i = item[1]
            x = eq_conv2d(x, P[f"encoder.{i}.weight.weight"], P[f"encoder.{i}.bias"], padding=1)
            x = F.relu(inst_norm(x))
            x = down_sample(x, P[f"encoder.{i + 3}.smooth.kernel"])
        else:
            i = item[1]  # blocks.py:32-33
            h = eq_conv2d(refl(x, 1), P[f"encoder.{i}.conv_block.1.weight.weight"])
            h = F.relu(inst_norm(h))
            h = eq_conv2d(refl(h, 1), P[f"encoder.{i}.conv_block.5.weight.weight"])
            x = x + inst_norm(h)
    return x


def _modconv(P, pre, x, w, padding):
    return modulated_conv2d(
        x,
        w,
        P[f"{pre}.weight.weight"],
        P[f"{pre}.to_style.weight.weight"],
        P[f"{pre}.to_style.bias"],
        padding,
    )


def generator_decode(P, z, w, arch: Arch, collect=False):
    """builder.py:220-249 (decode; extract when collect=True)."""
    _, dec = generator_plan(arch)
    feats = []
    i = 0
    for item in dec:
        if item[0] == "modres":
            j = item[1]  # blocks.py:62-68 — same w[i] for both convs
            h = _modconv(P, f"decoder.{j}.conv_block.1", refl(z, 1), w[i], 0)
            h = F.relu(h)
            h = _modconv(P, f"decoder.{j}.conv_block.4", refl(h, 1), w[i], 0)
            z = z + h
            feats.append(z)
            i += 1
        elif item[0] == "up":
            j = item[1]
            z = up_sample(z, P[f"decoder.{j}.smooth.kernel"])
            z = _modconv(P, f"decoder.{j + 1}", z, w[i], 1)
            i += 1
            if collect and i == arch.n_style_blocks:
                feats.append(z)  # the last styled layer is returned pre-ReLU (builder.py:243-244)
                return feats
            z = F.relu(z)
            # builder.py:190-197 uses nn.ReLU(inplace=True): it overwrites the tensor
            # `extract` already appended (builder.py:241), so every non-final
            # up-sample feature reaches path_loss_func POST-ReLU.  Pinned by golden case c.
            feats.append(z)
        else:
            if collect:
                return feats  # n_down == 0: last styled layer was a res block
            j = item[1]
            z = eq_conv2d(refl(z, 3), P[f"decoder.{j}.weight.weight"], P[f"decoder.{j}.bias"])
            z = torch.tanh(z)
    return z


def generator_extract(P, z, w, arch: Arch):
    return generator_decode(P, z, w, arch, collect=True)


def generator_forward(P, x, w, arch: Arch):
    return generator_decode(P, generator_encode(P, x, arch), w, arch)


def _patch_trunk(P, x):
    """Shared D/S trunk: builder.py:268-283 / :299-313."""
    x = eq_conv2d(x, P["model.0.weight.weight"], P["model.0.bias"], padding=1)
    x = F.leaky_relu(x, 0.2)
    x = down_sample(x, P["model.2.smooth.kernel"])
    for i, ds in ((3, 6), (7, 10)):
        x = eq_conv2d(x, P[f"model.{i}.weight.weight"], P[f"model.{i}.bias"], padding=1)
        x = F.leaky_relu(inst_norm(x), 0.2)
        x = down_sample(x, P[f"model.{ds}.smooth.kernel"])
    x = eq_conv2d(x, P["model.11.weight.weight"], P["model.11.bias"], padding=1)
    return F.leaky_relu(inst_norm(x), 0.2)


def discriminator_forward(P, x):
    x = _patch_trunk(P, x)
    return eq_conv2d(x, P["model.14.weight.weight"], P["model.14.bias"], padding=1)


def r1_penalty(P, real, gamma):
    """gamma/2 * mean_b ||grad_x sum D(x_b)||^2 on the reference Discriminator (BASELINE config
    5; not in the reference's training step: plain autograd double backward is the oracle)."""
    x = real.detach().clone().requires_grad_(True)
    scores = discriminator_forward(P, x)
    (gx,) = torch.autograd.grad(scores.sum(), x, create_graph=True)
    return gx.square().sum(dim=(1, 2, 3)).mean() * (0.5 * gamma)


def style_extractor_forward(P, x):
    x = _patch_trunk(P, x)
    x = x.mean(dim=(2, 3))  # AdaptiveAvgPool2d(1) + Flatten, builder.py:314-315
    return eq_linear(x, P["model.16.weight.weight"], P["model.16.bias"])


def mapping_forward(P, z, arch: Arch):
    """builder.py:46-49; last activation is ReLU (builder.py:36)."""
    z = F.normalize(z, dim=1)
    n = arch.mapping_network_layers
    for layer in range(n):
        z = eq_linear(z, P[f"net.{2 * layer}.weight.weight"], P[f"net.{2 * layer}.bias"])
        z = F.relu(z) if layer == n - 1 else F.leaky_relu(z, 0.2)
    return z


def sample_style(P, batch, n_blocks, arch: Arch, mix_styles=True):
    """builder.py:106-132 — draws on the default host generator, in order."""
    if mix_styles and torch.rand(()).lt(arch.style_mixing_prob):
        cross = int(torch.randint(0, n_blocks, ()))
        z1 = torch.randn(batch, arch.w_dim)  # host generator, then moved (builder.py:117-118)
        z2 = torch.randn(batch, arch.w_dim)
        s1 = mapping_forward(P, z1.to(_like(P)), arch)
        s2 = mapping_forward(P, z2.to(_like(P)), arch)
        return torch.cat(
            (s1[None].expand(cross, -1, -1), s2[None].expand(n_blocks - cross, -1, -1)), 0
        )
    z = torch.randn(batch, arch.w_dim)
    return mapping_forward(P, z.to(_like(P)), arch)[None].expand(n_blocks, -1, -1)


def _dtype_of(P):
    return next(iter(P.values())).dtype


def _like(P):
    """A parameter of the dict: `.to(_like(P))` adopts its dtype AND device (the oracle also runs
    on cuda:0 as bench.py's eager-PyTorch baseline; everywhere else it is a CPU program)."""
    return next(iter(P.values()))


def get_single_w(P, batch, n_blocks, arch: Arch, domain_variable, mix_styles=True):
    """builder.py:75-104; d == 0 draws nothing (builder.py:87-90)."""
    dt = _dtype_of(P)
    if not isinstance(domain_variable, Tensor) and domain_variable == 0:
        return torch.zeros(1, 1, arch.w_dim, dtype=dt, device=_like(P).device).expand(
            n_blocks, batch, arch.w_dim)
    s = sample_style(P, batch, n_blocks, arch, mix_styles)
    if isinstance(domain_variable, Tensor):
        d = domain_variable.view(1, -1, 1).to(s)
    else:
        d = torch.tensor(float(domain_variable), dtype=dt, device=s.device).view(1, 1, 1)
    return d * s  # lerp(0, s, d)


def get_two_w(P, batch, n_blocks, arch: Arch, d1, d2, mix_styles=True):
    """builder.py:51-73 — one sampled style, two domain variables."""
    s = sample_style(P, batch, n_blocks, arch, mix_styles)
    return d1.view(1, -1, 1).to(s) * s, d2.view(1, -1, 1).to(s) * s


# ---------------------------------------------------------------------------
# Losses (reference src/model/loss.py and the inline ones in training.py)
# ---------------------------------------------------------------------------


def style_cycle_loss(original_w, reconstructed_w, cos_l2_ratio=0.2):
    # loss.py:60-75
    a = F.normalize(original_w, dim=-1)
    b = F.normalize(reconstructed_w, dim=-1)
    cos = 1 - F.cosine_similarity(a, b, dim=-1).mean()
    return cos + cos_l2_ratio * F.mse_loss(a, b)


def kl_loss(latents):
    # loss.py:82-92 (global mean / biased variance)
    m = latents.mean()
    v = latents.var(correction=0)
    return m**2 + (v - 1) ** 2


def path_loss(f1, f2, h):
    # loss.py:98-111
    total = torch.zeros((), dtype=f1[0].dtype, device=f1[0].device)
    for a, b in zip(f1, f2, strict=True):
        jac = (a - b) / h[:, None, None, None]
        total = total + (jac**2).mean()
    return total / len(f1)


class ADAp:
    """loss.py:11-52 host-side controller (including its double append)."""

    def __init__(self, ada_e, ada_adjustment_size, batch_size, target):
        self.n_batches = ada_e // batch_size
        self.ada_adjustment = ada_adjustment_size * ada_e
        self.target = target
        self.p = 0.0
        self.curr_batch = 0
        self.scores: list[float] = []

    def update_p(self, mean_score: float):
        if self.curr_batch == self.n_batches:
            self.scores.append(mean_score)
            mean_sign = sum(self.scores) / len(self.scores)
            if mean_sign < self.target:
                self.p -= self.ada_adjustment
            elif mean_sign > self.target:
                self.p += self.ada_adjustment
            self.curr_batch = 0
            self.scores = []
            self.p = max(self.p, 0.0)
        self.curr_batch += 1
        self.scores.append(mean_score)

    def __call__(self):
        return self.p


class ImageBuffer:
    """training.py:22-65 history pool (python `random` draws once full)."""

    def __init__(self, buffer_size: int):
        if buffer_size < 1:
            raise ValueError
        self.buffer_size = buffer_size
        self.images: list[Tensor] = []

    def __call__(self, images: Tensor) -> Tensor:
        out = []
        for img in images:
            img = img.detach()[None]
            if len(self.images) < self.buffer_size:
                self.images.append(img)
                out.append(img)
            elif random.uniform(0, 1) > 0.5:
                k = random.randint(0, self.buffer_size - 1)
                out.append(self.images[k].clone())
                self.images[k] = img
            else:
                out.append(img)
        return torch.cat(out, 0)


# ---------------------------------------------------------------------------
# Adam (torch.optim.Adam defaults, train.py:94-116)
# ---------------------------------------------------------------------------


class Adam:
    def __init__(self, params: dict[str, Tensor], lr, betas, eps=1e-8):
        self.names = [k for k in params if not is_buffer(k)]
        self.lr, self.betas, self.eps = lr, betas, eps
        self.m = {k: torch.zeros_like(params[k]) for k in self.names}
        self.v = {k: torch.zeros_like(params[k]) for k in self.names}
        self.t = 0

    @torch.no_grad()
    def step(self, params, grads):
        self.t += 1
        b1, b2 = self.betas
        c1 = 1 - b1**self.t
        c2 = 1 - b2**self.t
        for k in self.names:
            g = grads.get(k)
            if g is None:
                continue
            self.m[k].mul_(b1).add_(g, alpha=1 - b1)
            self.v[k].mul_(b2).addcmul_(g, g, value=1 - b2)
            denom = (self.v[k].sqrt() / math.sqrt(c2)).add_(self.eps)
            params[k].addcdiv_(self.m[k], denom, value=-self.lr / c1)


# ---------------------------------------------------------------------------
# Training step (reference src/core/training.py)
# ---------------------------------------------------------------------------


@dataclass
class Hyper:
    """config.toml [training]/[optimisation]/[ada] values used by the step."""

    batch_size: int = 4
    image_buffer_size: int = 100
    style_cycle_loss_lambda: float = 5.0
    identity_loss_lambda: float = 5.0
    reconstruction_loss_lambda: float = 5.0
    kl_loss_lambda: float = 0.01
    path_loss_lambda: float = 0.1
    path_h_range: tuple[float, float] = (0.1, 0.2)
    learning_rate: float = 2e-3
    mapping_network_learning_rate: float = 2e-5
    adam_betas: tuple[float, float] = (0.5, 0.99)
    add_latent_noise: bool = False
    r1_gamma: float = 0.0  # extension (BASELINE config 5); 0 = the reference
    ada_target: float = 0.6
    ada_e: int = 256
    ada_adjustment_size: float = 5.12e-4


def _leaf(params):
    return {
        k: (v.detach().clone().requires_grad_(True) if not is_buffer(k) else v)
        for k, v in params.items()
    }


@dataclass
class Trainer:
    """State that train.py:72-195 builds, plus the two step functions."""

    arch: Arch
    hyper: Hyper
    params: dict[str, dict[str, Tensor]]
    dtype: torch.dtype = torch.float32
    device: str = "cpu"  # "cuda:0" only for bench.py's eager-PyTorch-on-GPU baseline
    opt: dict[str, Adam] = field(default_factory=dict)
    last_grads: dict[str, dict[str, Tensor]] = field(default_factory=dict)
    last_h: Tensor | None = None

    def __post_init__(self):
        self.params = {
            n: {k: v.to(device=self.device, dtype=self.dtype).clone() for k, v in p.items()}
            for n, p in self.params.items()
        }
        h = self.hyper
        for n in ("D", "G", "S"):
            self.opt[n] = Adam(self.params[n], h.learning_rate, h.adam_betas)
        self.opt["M"] = Adam(self.params["M"], h.mapping_network_learning_rate, h.adam_betas)
        self.buffer = ImageBuffer(h.image_buffer_size)
        self.ada_p = ADAp(h.ada_e, h.ada_adjustment_size, h.batch_size, h.ada_target)

    # training.py:71-128
    def discriminator_step(self, shoeprints: Tensor, shoemarks: Tensor):
        a, h = self.arch, self.hyper
        D = _leaf(self.params["D"])
        G, M = self.params["G"], self.params["M"]
        with torch.no_grad():  # the reference builds and discards this graph (:98)
            w = get_single_w(M, h.batch_size, a.n_style_blocks, a, 1)
            fake = generator_forward(G, shoeprints.to(_like(G)), w, a)
        fake = self.buffer(fake)
        real = shoemarks.to(_like(G))
        fake_scores = discriminator_forward(D, fake)
        real_scores = discriminator_forward(D, real)
        real_loss = F.mse_loss(real_scores, torch.ones_like(real_scores))
        fake_loss = F.mse_loss(fake_scores, torch.zeros_like(fake_scores))
        loss = (real_loss + fake_loss) / 2
        if h.r1_gamma > 0:
            loss = loss + r1_penalty(D, real, h.r1_gamma)
        sign_real = torch.sign(real_scores.detach() * 2 - 1).mean()
        sign_fake = -torch.sign(fake_scores.detach() * 2 - 1).mean()
        self.ada_p.update_p(float(sign_real))
        names = [k for k in D if not is_buffer(k)]
        grads = torch.autograd.grad(loss, [D[k] for k in names])
        g = dict(zip(names, grads))
        self.last_grads["D"] = g
        self.opt["D"].step(self.params["D"], g)
        return float(loss), (float(sign_real), float(sign_fake))

    # training.py:136-257
    def generator_step(self, shoeprints: Tensor, shoemarks: Tensor, h_override: Tensor | None = None,
                       n_styles: int = 1):
        """n_styles = K > 1 is BASELINE config 4 (one input -> K sampled outputs): every
        sampled-style pass (translation decode, D and S on the translations, both path-length
        extractions) runs on the latent of each shoeprint broadcast to K styles, image (b, k)
        at index b*K + k -- the `latent.expand(K, ...)` of evaluation.py:172-177 applied to the
        training step; reconstruction (w = 0) and identity stay at B.  K = 1 is the reference."""
        a, hy = self.arch, self.hyper
        B, nb = hy.batch_size, a.n_style_blocks
        K = int(n_styles)
        BK = B * K
        G, M, S = _leaf(self.params["G"]), _leaf(self.params["M"]), _leaf(self.params["S"])
        D = self.params["D"]
        prints = shoeprints.to(_like(G))
        marks = shoemarks.to(_like(G))
        latents = generator_encode(G, torch.cat([prints, marks], 0), a)
        kl = kl_loss(latents)
        if hy.add_latent_noise:
            latents = latents + torch.randn_like(latents)
        print_lat, mark_lat = latents.chunk(2, dim=0)
        # reconstruction (:171-180)
        w0 = get_single_w(M, B, nb, a, 0)
        recon = generator_decode(G, print_lat, w0, a)
        rec_loss = F.l1_loss(recon, prints)
        # identity (:183-190)
        mark_w = style_extractor_forward(S, marks)
        ident = generator_decode(G, mark_lat, mark_w.expand(nb, *mark_w.shape), a)
        idt_loss = F.l1_loss(ident, marks)
        # GAN (:193-204)
        print_lat_k = print_lat if K == 1 else print_lat.repeat_interleave(K, dim=0)
        tw = get_single_w(M, BK, nb, a, 1)
        transl = generator_decode(G, print_lat_k, tw, a)
        scores = discriminator_forward(D, transl)
        gan_loss = F.mse_loss(scores, torch.ones_like(scores))
        # style cycle (:207-210)
        style_loss = style_cycle_loss(tw[-1], style_extractor_forward(S, transl))
        # path length (:214-234)
        theta = torch.rand(BK).to(prints)  # host draw, then moved (training.py:214)
        if h_override is None:
            # drawn in fp32 on the host generator whatever self.dtype is, so an fp64 replay
            # consumes the generator exactly like the reference's fp32 CPU run (training.py:216-223)
            hh = torch.ones(BK).uniform_(*hy.path_h_range).to(prints)
        else:
            hh = h_override.to(prints)
        self.last_h = hh
        d1 = (theta + hh / 2).clamp(0, 1)
        d2 = (theta - hh / 2).clamp(0, 1)
        w1, w2 = get_two_w(M, BK, nb, a, d1, d2)
        f1 = generator_extract(G, print_lat_k, w1, a)
        f2 = generator_extract(G, print_lat_k, w2, a)
        p_loss = path_loss(f1, f2, hh)
        total = (
            gan_loss
            + hy.identity_loss_lambda * idt_loss
            + hy.reconstruction_loss_lambda * rec_loss
            + hy.kl_loss_lambda * kl
            + hy.path_loss_lambda * p_loss
            + hy.style_cycle_loss_lambda * style_loss
        )
        # ONE backward pass for the three networks, like the reference's total.backward() (:245)
        keys = [(name, k) for name, P in (("G", G), ("M", M), ("S", S)) for k in P if not is_buffer(k)]
        nets = {"G": G, "M": M, "S": S}
        grads = torch.autograd.grad(total, [nets[n][k] for n, k in keys], allow_unused=True)
        for name in nets:
            self.last_grads[name] = {}
        for (name, k), gr in zip(keys, grads):
            self.last_grads[name][k] = gr if gr is not None else torch.zeros_like(nets[name][k])
        for name in ("G", "M", "S"):
            self.opt[name].step(self.params[name], self.last_grads[name])
        return float(total), (
            float(gan_loss),
            float(rec_loss),
            float(idt_loss),
            float(kl),
            float(p_loss),
            float(style_loss),
        )


def synthetic_batch(batch, arch: Arch, seed: int):
    """U(-1,1) images of the configured shape (the range Normalize(0.5,0.5)
    yields, train.py:120-126) from a private generator (does not perturb the
    default host generator, like the DataLoader's own generator train.py:56)."""
    g = torch.Generator().manual_seed(seed)
    return torch.rand(batch, arch.image_channels, *arch.image_size, generator=g) * 2 - 1
