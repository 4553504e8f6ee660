"""The small dense kernels (csrc/style.cu) against the oracle's restatement / plain torch autograd
in fp64: EqualisedLinear in multi-job form (reference layers.py:27-43), the MappingNetwork with
style mixing and domain-variable interpolation (builder.py:46-132), the style-cycle loss
(loss.py:60-75).  fp32 kernels: gate 1e-5 norm-relative (north_star: 1e-4)."""

import pytest
import torch
import torch.nn.functional as F

from oracle import reference_port as rp

pytestmark = pytest.mark.gpu


def relerr(a, b):
    a, b = a.detach().double().cpu(), b.detach().double().cpu()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(*shape, seed=0):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).cuda()


class _Lin:  # duck-typed EqualisedLinear
    def __init__(self, o, k, seed):
        from one_to_many_gan_b200.layers import EqualisedLinear

        torch.manual_seed(seed)
        self.m = EqualisedLinear(k, o, bias=1.0).cuda()
        with torch.no_grad():
            self.m.bias.add_(rnd(o, seed=seed + 1))


def test_multi_job_linear_forward_backward():
    from one_to_many_gan_b200 import ops

    B, wd = 12, 6
    w = rnd(3, B, wd, seed=1).requires_grad_(True)          # [n_blocks, B, w_dim]
    w0 = torch.zeros(1, 1, wd, device="cuda").expand(3, B, wd)  # stride-0 zero style (builder.py:88-90)
    pooled = rnd(B, 512, seed=2).requires_grad_(True)       # StyleExtractor head input
    lins = [_Lin(64, wd, 10).m, _Lin(128, wd, 20).m, _Lin(256, wd, 30).m, _Lin(wd, 512, 40).m,
            _Lin(64, wd, 50).m]
    xs = [w[0], w[0], w[2], pooled, w0[1]]                  # two jobs share w[0]; one broadcast row
    ys = ops.linears(xs, lins)
    gys = [rnd(*y.shape, seed=60 + i) for i, y in enumerate(ys)]
    torch.autograd.backward(ys, gys)
    got = [w.grad.clone(), pooled.grad.clone()] + [m.weight.weight.grad.clone() for m in lins] + \
          [m.bias.grad.clone() for m in lins]
    # reference: fp64 autograd of the oracle's eq_linear
    w64 = w.detach().double().requires_grad_(True)
    p64 = pooled.detach().double().requires_grad_(True)
    W64 = [m.weight.weight.detach().double().requires_grad_(True) for m in lins]
    b64 = [m.bias.detach().double().requires_grad_(True) for m in lins]
    xs64 = [w64[0], w64[0], w64[2], p64, w0[1].double()]
    ys64 = [rp.eq_linear(x, W, b) for x, W, b in zip(xs64, W64, b64)]
    for y, y64 in zip(ys, ys64):
        assert relerr(y, y64) < 1e-5
    torch.autograd.backward(ys64, [g.double() for g in gys])
    want = [w64.grad, p64.grad] + [t.grad for t in W64] + [t.grad for t in b64]
    for i, (a, b) in enumerate(zip(got, want)):
        assert relerr(a, b) < 1e-5, i


@pytest.mark.parametrize("mix", [False, True])
@pytest.mark.parametrize("two", [False, True])
def test_mapping_network_fused(mix, two):
    from one_to_many_gan_b200 import builder, ops

    arch = rp.Arch()
    torch.manual_seed(3)
    M = builder.MappingNetwork(arch.w_dim, arch.mapping_network_layers, 0.9).cuda()
    P = {k: v.detach().double().cpu().requires_grad_(True) for k, v in M.state_dict().items()}
    B, nb, cross = 10, 6, 4
    z1 = rnd(B, arch.w_dim, seed=5)
    z2 = rnd(B, arch.w_dim, seed=6) if mix else None
    z1[3] = 0.0  # F.normalize eps path
    d = [torch.rand(B, generator=torch.Generator().manual_seed(7 + j)).cuda() for j in range(2)]
    cross_dev = torch.tensor([cross if mix else nb], dtype=torch.int64, device="cuda")
    # plain forward (module API)
    assert relerr(M(z1), rp.mapping_forward(P, z1.double().cpu(), arch)) < 1e-5
    outs = ops.mapping(M.linears(), z1, z2, cross_dev, n_blocks=nb,
                       d=(d[0], d[1] if two else None) if two else (None, None), n_out=2 if two else 1)
    s1 = rp.mapping_forward(P, z1.double().cpu(), arch)
    s2 = rp.mapping_forward(P, z2.double().cpu(), arch) if mix else s1
    s = torch.cat([s1[None].expand(cross, -1, -1), s2[None].expand(nb - cross, -1, -1)]) if mix \
        else s1[None].expand(nb, -1, -1)
    refs = [dj.double().cpu().view(1, -1, 1) * s for dj in d] if two else [s]
    gs = [rnd(nb, B, arch.w_dim, seed=9 + j) for j in range(len(outs))]
    for o, r in zip(outs, refs):
        assert relerr(o, r) < 1e-5
    torch.autograd.backward(outs, gs)
    torch.autograd.backward(refs, [g.double().cpu() for g in gs])
    for k, p in M.named_parameters():
        assert relerr(p.grad, P[k].grad) < 1e-5, k


@pytest.mark.parametrize("B,Fd", [(4, 6), (96, 6), (300, 8)])
def test_style_cycle_loss(B, Fd):
    from one_to_many_gan_b200 import ops

    a = F.relu(rnd(3, B, Fd, seed=11))[-1]  # a row view like translation_w[-1]
    a[1] = 0.0                               # an all-zero style (ReLU output): normalize eps path
    b = rnd(B, Fd, seed=12)
    a_, b_ = a.clone().requires_grad_(True), b.clone().requires_grad_(True)
    wl, raw = ops.style_cycle(a_, b_, 5.0)
    a64, b64 = a.double().cpu().requires_grad_(True), b.double().cpu().requires_grad_(True)
    ref = rp.style_cycle_loss(a64, b64)
    assert abs(raw.item() - ref.item()) < 1e-5 * max(1.0, abs(ref.item()))
    assert abs(wl.item() - 5 * ref.item()) < 5e-5 * max(1.0, abs(ref.item()))
    wl.backward()
    (5 * ref).backward()
    rows = [i for i in range(B) if i != 1]  # the zero row's gradient is 1e12-scaled noise on both sides
    assert relerr(a_.grad[rows], a64.grad[rows]) < 1e-5
    assert relerr(b_.grad, b64.grad) < 1e-5
    assert torch.isfinite(a_.grad).all()


def test_channel_sum_accumulates():
    from one_to_many_gan_b200 import kernels as K

    x = rnd(3, 64, 9, 7, seed=13)
    xt = K.alloc(3, 64, 9, 7, torch.float32, x.device)
    xt.copy_(x)
    out = torch.full((64,), 2.0, device="cuda")
    K.channel_sum(xt, out=out)
    assert relerr(out - 2.0, x.sum((0, 2, 3))) < 1e-5
    assert relerr(K.channel_sum(xt), x.sum((0, 2, 3))) < 1e-5
