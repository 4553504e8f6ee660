"""The TMA row-streaming kernels (row_stream_kernel ops, down_stream_kernel) only engage on tensors
of >= 8 MB, above the shapes of test_kernels_gpu.py: here they run on 17-33 MB tensors against
fp32 torch autograd on the same (storage-rounded) inputs, including the reflect-halo folds of
width 1 and 3 and an odd-size DownSample.  The (8, 256, 64, 64) / (4, 512, 64, 64) shapes are the
256- and 512-channel latents of BASELINE configs 3 and 5: their rows (32-64 KB) do not fit a
multi-stage ring, so they are streamed as 2-8 column segments per row."""

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {torch.bfloat16: 2e-2}


@pytest.fixture(scope="module")
def K():
    from one_to_many_gan_b200 import kernels

    return kernels


def relerr(a, b):
    a, b = a.detach().double(), b.detach().double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def rnd(*shape, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(*shape, generator=g).cuda()


def to_nhwc(K, x, halo=0):
    n, c, h, w = x.shape
    t = K.alloc(n, c, h, w, torch.bfloat16, x.device, halo, zero=True)
    t.copy_(x.to(torch.bfloat16))
    return t


def padded_grad(K, gp):
    """gp: [n,c,h+2p,w+2p] gradient w.r.t. a reflect-padded tensor -> its interior view (NHWC)."""
    n, c, hp, wp = gp.shape
    buf = gp.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()
    return buf


SHAPES = [(16, 128, 64, 64), (8, 256, 64, 64), (4, 512, 64, 64), (4, 64, 256, 256)]  # >= 16.8 MB each


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("halo", [0, 1, 3])
@pytest.mark.parametrize("act", ["none", "relu"])
def test_stream_norm_act_bwd(K, halo, act, shape):
    n, c, h, w = shape
    x = rnd(n, c, h, w, seed=1).bfloat16().float()
    gp = rnd(n, c, h + 2 * halo, w + 2 * halo, seed=2).bfloat16().float()
    g2 = rnd(n, c, h, w, seed=3).bfloat16().float()
    xr = x.clone().requires_grad_(True)
    z = F.instance_norm(xr, eps=1e-5)
    z = F.relu(z) if act == "relu" else z
    zp = F.pad(z, (halo,) * 4, mode="reflect") if halo else z
    ((zp * gp).sum() + (z * g2).sum()).backward()
    xt = to_nhwc(K, x)
    stats = K.instnorm_stats(xt)
    ref_mean = x.mean(dim=(2, 3))
    assert (stats[..., 0] - ref_mean).abs().max().item() < 2e-3
    buf = gp.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()  # NHWC padded buffer
    g_int = buf[:, halo : halo + h, halo : halo + w, :].permute(0, 3, 1, 2)
    gx = K.norm_act_bwd(g_int, xt, stats, K.ACT_RELU if act == "relu" else K.ACT_NONE, g_halo=halo,
                        g2=to_nhwc(K, g2))
    assert relerr(gx.float(), xr.grad) < TOL[torch.bfloat16], (halo, act)


@pytest.mark.parametrize("shape", SHAPES)
def test_stream_fold_add_and_gres(K, shape):
    (n, c, h, w), halo = shape, 1
    gp = rnd(n, c, h + 2, w + 2, seed=4).bfloat16().float()
    g2 = rnd(n, c, h, w, seed=5).bfloat16().float()
    z = torch.zeros(n, c, h, w, device="cuda", requires_grad=True)
    ((F.pad(z, (1,) * 4, mode="reflect") * gp).sum() + (z * g2).sum()).backward()
    buf = gp.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()
    g_int = buf[:, 1 : 1 + h, 1 : 1 + w, :].permute(0, 3, 1, 2)
    gx = K.norm_act_bwd(g_int, None, None, K.ACT_NONE, g_halo=halo, g2=to_nhwc(K, g2))
    assert relerr(gx.float(), z.grad) < 1e-2


@pytest.mark.parametrize("shape", SHAPES)
def test_stream_mod_in(K, shape):
    n, c, h, w = shape
    x = F.relu(rnd(n, c, h, w, seed=6)).bfloat16().float()
    gp = rnd(n, c, h + 2, w + 2, seed=7).bfloat16().float()
    gadd = rnd(n, c, h, w, seed=8).bfloat16().float()
    s = torch.rand(n, c, device="cuda") + 0.5
    z = torch.zeros(n, c, h, w, device="cuda", requires_grad=True)
    (F.pad(z, (1,) * 4, mode="reflect") * gp).sum().backward()
    fold = z.grad
    want_q = (fold * x).sum(dim=(2, 3))
    want_gx = (fold * s[:, :, None, None] + gadd) * (x > 0)
    buf = gp.to(torch.bfloat16).permute(0, 2, 3, 1).contiguous()
    g_int = buf[:, 1 : 1 + h, 1 : 1 + w, :].permute(0, 3, 1, 2)
    gx, Q = K.mod_in(g_int, to_nhwc(K, x), s, g_halo=1, gadd=to_nhwc(K, gadd), relu_mask=True)
    assert relerr(gx.float(), want_gx) < TOL[torch.bfloat16]
    assert relerr(Q, want_q) < 5e-3


@pytest.mark.parametrize("shape", SHAPES)
def test_stream_norm_act_fwd_and_channel_sum(K, shape):
    n, c, h, w = shape
    x = rnd(n, c, h, w, seed=9).bfloat16().float()
    res = rnd(n, c, h, w, seed=10).bfloat16().float()
    xt = to_nhwc(K, x)
    stats = K.instnorm_stats(xt)
    y = K.norm_act(xt, stats, K.ACT_RELU, residual=to_nhwc(K, res), y_halo=1)
    want = F.relu(F.instance_norm(x, eps=1e-5)) + res
    assert relerr(y.float(), want) < TOL[torch.bfloat16]
    assert torch.equal(K.padded_view(y, 1).float(), F.pad(y.float(), (1,) * 4, mode="reflect"))
    cs = K.channel_sum(xt)
    assert relerr(cs, x.sum(dim=(0, 2, 3))) < 2e-3


@pytest.mark.parametrize("shape", [(16, 64, 128, 128), (16, 64, 127, 127)])
def test_stream_down(K, shape):
    from oracle import reference_port as rp

    kern = rp._smooth_kernel().cuda()
    x = rnd(*shape, seed=11).bfloat16().float()
    want = rp.down_sample(F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2), kern)
    xt = to_nhwc(K, x)
    out = K.down(xt, K.instnorm_stats(xt), K.ACT_LRELU, 1)
    assert relerr(out.float(), want) < TOL[torch.bfloat16], shape
    assert torch.equal(K.padded_view(out, 1).float(), F.pad(out.float(), (1,) * 4, mode="reflect"))
    plain = K.down(xt)
    assert relerr(plain.float(), rp.down_sample(x, kern)) < TOL[torch.bfloat16], shape


@pytest.mark.parametrize("shape", [(16, 64, 128, 128), (16, 64, 127, 127), (16, 128, 62, 62)])
def test_stream_down_bwd(K, shape):
    from oracle import reference_port as rp

    kern = rp._smooth_kernel().cuda()
    x = rnd(*shape, seed=12).requires_grad_(True)
    y = rp.down_sample(x, kern)
    g = rnd(*y.shape, seed=13).bfloat16().float()
    (want,) = torch.autograd.grad(y, x, g)
    ga = K.down_bwd(to_nhwc(K, g), shape[2:])
    assert relerr(ga.float(), want) < TOL[torch.bfloat16], shape


@pytest.mark.parametrize("shape", [(16, 128, 64, 64), (8, 64, 96, 80)])
def test_stream_up(K, shape):
    from oracle import reference_port as rp

    kern = rp._smooth_kernel().cuda()
    x = rnd(*shape, seed=14).bfloat16().float()
    want = rp.up_sample(x, kern)
    out = K.up(to_nhwc(K, x))
    assert relerr(out.float(), want) < TOL[torch.bfloat16], shape


@pytest.mark.parametrize("shape", [(16, 128, 64, 64), (4, 64, 128, 128), (8, 256, 32, 32), (16, 128, 63, 31),
                                   (4, 128, 48, 80)])
def test_stream_up_bwd(K, shape):
    """UpSample backward at bench scale (2 x 2 block kernel: interior, border and odd-size blocks),
    with and without the per-(n, c) style scale, against autograd of the oracle's UpSample."""
    from oracle import reference_port as rp

    kern = rp._smooth_kernel().cuda()
    n, c, h, w = shape
    x = rnd(n, c, h, w, seed=15).requires_grad_(True)
    g = rnd(n, c, 2 * h, 2 * w, seed=16).bfloat16().float()
    (want,) = torch.autograd.grad(rp.up_sample(x, kern), x, g)
    gt = to_nhwc(K, g)
    assert relerr(K.up_bwd(gt).float(), want) < TOL[torch.bfloat16], shape
    s = torch.rand(n, c, device="cuda") + 0.5
    assert relerr(K.up_bwd(gt, scale=s).float(), want * s[:, :, None, None]) < TOL[torch.bfloat16], shape
