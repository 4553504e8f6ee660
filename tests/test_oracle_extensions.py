"""CPU checks of the two extension paths of the oracle (BASELINE configs 4 and 5), which have no
reference counterpart in the training step: properties that tie them to the pinned K = 1 /
gamma = 0 step and to an independent evaluation of the same quantity."""

import random

import torch

from oracle import reference_port as rp


def _setup(batch=2):
    arch = rp.Arch(image_size=(32, 32), min_latent_resolution=16, n_resnet_blocks=3)
    return arch, rp.init_all(arch, 42)


def _imgs(shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(*shape, generator=g) * 2 - 1


def test_styles_per_input_keeps_the_style_independent_terms():
    """K > 1 only widens the sampled-style passes: reconstruction, identity and KL terms are
    those of the pinned K = 1 step; GAN / path / style-cycle terms change."""
    arch, P = _setup()
    shape = (2, 1, 32, 32)
    out = {}
    for K in (1, 3):
        torch.manual_seed(5)
        random.seed(5)
        tr = rp.Trainer(arch, rp.Hyper(batch_size=2), P, dtype=torch.float64)
        h = torch.tensor([0.13, 0.17, 0.11, 0.19, 0.15, 0.12][: 2 * K])
        out[K] = tr.generator_step(_imgs(shape, 3), _imgs(shape, 4), h_override=h, n_styles=K)
        assert tr.last_h.shape == (2 * K,)
    (_, (gan1, rec1, idt1, kl1, path1, sty1)) = out[1]
    (_, (gan3, rec3, idt3, kl3, path3, sty3)) = out[3]
    assert rec1 == rec3 and idt1 == idt3 and kl1 == kl3
    assert gan1 != gan3 and path1 != path3 and sty1 != sty3


def test_styles_per_input_with_repeated_style_equals_k1():
    """With the same style for every k (no mixing, style noise repeated) the K-step is the K = 1
    step: means over (b, k) of K identical copies."""
    arch, P = _setup()
    P64 = {n: {k: v.double() for k, v in p.items()} for n, p in P.items()}
    B, K, nb = 2, 3, arch.n_style_blocks
    x = _imgs((B, 1, 32, 32), 9).double()
    w = torch.rand(nb, B, arch.w_dim, generator=torch.Generator().manual_seed(1)).double()
    with torch.no_grad():
        z = rp.generator_encode(P64["G"], x, arch)
        y1 = rp.generator_decode(P64["G"], z, w, arch)
        yk = rp.generator_decode(P64["G"], z.repeat_interleave(K, 0), w.repeat_interleave(K, 1), arch)
    torch.testing.assert_close(yk[::K], y1, rtol=1e-12, atol=1e-12)
    torch.testing.assert_close(yk[1::K], y1, rtol=1e-12, atol=1e-12)


def test_r1_penalty_properties():
    """R1: non-negative, linear in gamma, equal to gamma/2 * mean ||J^T 1||^2 with the input
    gradient evaluated by central finite differences along random directions."""
    arch, P = _setup()
    D = {k: v.double() for k, v in P["D"].items()}
    x = _imgs((2, 1, 32, 32), 11).double()
    p1 = rp.r1_penalty(D, x, 1.0)
    p10 = rp.r1_penalty(D, x, 10.0)
    assert p1.item() > 0
    torch.testing.assert_close(p10, 10.0 * p1, rtol=1e-12, atol=0)
    # directional derivative check of grad_x sum(D(x)): <g, d> = (f(x + e d) - f(x - e d)) / 2e
    xr = x.clone().requires_grad_(True)
    (g,) = torch.autograd.grad(rp.discriminator_forward(D, xr).sum(), xr)
    torch.testing.assert_close(p1, 0.5 * g.square().sum(dim=(1, 2, 3)).mean(), rtol=1e-12, atol=0)
    d = torch.randn(x.shape, generator=torch.Generator().manual_seed(2)).double()
    eps = 1e-5
    with torch.no_grad():
        fd = (rp.discriminator_forward(D, x + eps * d).sum()
              - rp.discriminator_forward(D, x - eps * d).sum()) / (2 * eps)
    torch.testing.assert_close((g * d).sum(), fd, rtol=1e-5, atol=1e-8)


def test_r1_zero_gamma_is_the_reference_step():
    arch, P = _setup()
    shape = (2, 1, 32, 32)
    res = []
    for gamma in (0.0, 0.0):
        torch.manual_seed(5)
        random.seed(5)
        tr = rp.Trainer(arch, rp.Hyper(batch_size=2, r1_gamma=gamma), P)
        res.append(tr.discriminator_step(_imgs(shape, 1), _imgs(shape, 2)))
    assert res[0] == res[1]
