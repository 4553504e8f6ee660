"""Per-kernel parity through the C-ABI (ctypes) against plain torch fp32/fp64 references of
the same op (the reference's own L0 ops: F.conv2d, F.instance_norm, F.pad, F.interpolate …).
Tolerances: 1e-4 norm-relative for fp32 storage, 2e-2 for bf16 storage (BASELINE north_star)."""

import math

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

TOL = {torch.float32: 1e-4, torch.bfloat16: 2e-2}


@pytest.fixture(scope="module")
def K():
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    from one_to_many_gan_b200 import kernels

    return kernels


def relerr(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


def nhwc(x_nchw, dtype, halo=0, mode="reflect"):
    """Channels-last copy of an NCHW reference tensor; with halo>0 the halo is materialised."""
    from one_to_many_gan_b200 import kernels as K

    n, c, h, w = x_nchw.shape
    t = K.alloc(n, c, h, w, dtype, x_nchw.device, halo)
    if halo:
        full = F.pad(x_nchw, (halo,) * 4, mode=mode) if mode != "zero" else F.pad(x_nchw, (halo,) * 4)
        K.padded_view(t, halo).copy_(full.to(dtype))
    else:
        t.copy_(x_nchw.to(dtype))
    return t


def rnd(*shape, seed=0, dev="cuda"):
    g = torch.Generator(device="cpu").manual_seed(seed)
    return torch.randn(*shape, generator=g).to(dev)


CONV_CASES = [
    # cin, cout, k, pad, halo(reflect), H, W, n
    (1, 64, 7, 3, 3, 20, 12, 2),
    (64, 128, 3, 1, 0, 16, 24, 2),
    (64, 64, 3, 1, 1, 16, 16, 3),
    (128, 128, 3, 1, 1, 16, 16, 2),
    (256, 256, 3, 1, 1, 8, 16, 2),
    (64, 128, 4, 1, 0, 31, 17, 2),
    (128, 256, 4, 1, 0, 15, 15, 2),
    (256, 512, 4, 1, 0, 9, 9, 2),
    (512, 1, 4, 1, 0, 7, 6, 2),
    (64, 1, 7, 3, 3, 16, 12, 2),
    (1, 64, 4, 1, 0, 33, 20, 2),
    (24, 40, 3, 1, 0, 9, 11, 1),
    (64, 1, 7, 3, 3, 40, 24, 2),   # smem-tiled Cout=1 kernel (G output conv)
    (64, 1, 4, 2, 0, 33, 20, 2),   # ... as the dgrad of a 1->64 4x4 conv
    (128, 1, 4, 1, 0, 20, 20, 1),  # ... two channel chunks
]


def _ref_conv(x, w, pad, halo, alpha):
    xin = F.pad(x, (halo,) * 4, mode="reflect") if halo else x
    return F.conv2d(xin, w * alpha, padding=pad - halo)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_fwd_epilogue(K, case, dtype):
    cin, cout, k, pad, halo, H, W, n = case
    x = rnd(n, cin, H, W, seed=1)
    w = rnd(cout, cin, k, k, seed=2)
    alpha = 1 / math.sqrt(cin * k * k)
    xq = x.to(dtype).float()
    wq = (w * alpha).to(dtype).float()
    ref = _ref_conv(xq, wq, pad, halo, 1.0)
    rs = torch.rand(n, cout, device="cuda") + 0.5
    bias = rnd(cout, seed=3)
    res = rnd(*ref.shape, seed=4)
    full = F.leaky_relu(ref * rs[:, :, None, None] + bias[None, :, None, None], 0.2) + res.to(dtype).float()
    xt = nhwc(x, dtype, halo)
    wp = K.weight_pack(w, alpha, dtype)
    y = K.conv_fwd(xt, wp, cout, k, k, pad, x_halo=halo)
    assert relerr(y.float(), ref) < TOL[dtype], f"plain conv {case}"
    yh = min(2, ref.shape[2] - 1, ref.shape[3] - 1)
    y2 = K.conv_fwd(xt, wp, cout, k, k, pad, x_halo=halo, row_scale=rs, bias=bias, act=K.ACT_LRELU,
                    residual=nhwc(res, dtype), y_halo=yh)
    assert relerr(y2.float(), full) < TOL[dtype], f"epilogue {case}"
    if yh:
        want = F.pad(y2.float(), (yh,) * 4, mode="reflect")
        got = K.padded_view(y2, yh).float()
        assert torch.equal(got, want), f"reflect halo {case}"


@pytest.mark.parametrize("case", [c for c in CONV_CASES if c[0] % 64 == 0 and c[1] % 64 == 0])
def test_conv_tc_matches_simt(K, case):
    """tcgen05 path vs the fp32-accumulating SIMT path on identical bf16 operands."""
    cin, cout, k, pad, halo, H, W, n = case
    x = nhwc(rnd(n, cin, H, W, seed=5), torch.bfloat16, halo)
    w = rnd(cout, cin, k, k, seed=6)
    wp = K.weight_pack(w, 1 / math.sqrt(cin * k * k), torch.bfloat16)
    a = K.conv_fwd(x, wp, cout, k, k, pad, x_halo=halo, path=K.PATH_TC)
    b = K.conv_fwd(x, wp, cout, k, k, pad, x_halo=halo, path=K.PATH_SIMT)
    assert relerr(a.float(), b.float()) < 4e-3, case


PAIR_CASES = [
    # cin, cout, k, pad, x_halo, H, W, n : large enough (>= 2 x 148 tile pairs) for the persistent
    # PAIR kernels the bench-size layers run on (two 16x8 pixel tiles per CTA step)
    (128, 128, 3, 1, 1, 64, 64, 20),   # res-block / modulated res-block conv
    (128, 128, 3, 2, 0, 64, 64, 20),   # its dgrad: 66x66 output, odd number of tile rows
    (64, 128, 4, 1, 0, 63, 63, 24),    # discriminator 4x4
    (128, 64, 3, 1, 0, 64, 64, 20),    # BN = 64 pair kernel
]


@pytest.mark.parametrize("case", PAIR_CASES)
def test_conv_tc_pair_kernels(K, case):
    """tcgen05 pair kernels (incl. the transposed 256-pixel-operand kernel) vs the FFMA path on
    identical bf16 operands: plain, and with per-sample weights + demodulation scale + bias +
    ReLU + residual + reflect halo."""
    cin, cout, k, pad, halo, H, W, n = case
    x = nhwc(rnd(n, cin, H, W, seed=21), torch.bfloat16, halo)
    w = rnd(cout, cin, k, k, seed=22)
    alpha = 1 / math.sqrt(cin * k * k)
    wp = K.weight_pack(w, alpha, torch.bfloat16)
    a = K.conv_fwd(x, wp, cout, k, k, pad, x_halo=halo, path=K.PATH_TC)
    b = K.conv_fwd(x, wp, cout, k, k, pad, x_halo=halo, path=K.PATH_SIMT)
    assert relerr(a.float(), b.float()) < 4e-3, ("plain", case)
    s = torch.rand(n, cin, device="cuda") + 0.5
    rs = torch.rand(n, cout, device="cuda") + 0.5
    bias = rnd(cout, seed=23)
    res = nhwc(rnd(*a.shape, seed=24), torch.bfloat16)
    wps = K.weight_pack(w, alpha, torch.bfloat16, cs=s, nb=n)
    kw = dict(x_halo=halo, row_scale=rs, bias=bias, act=K.ACT_RELU, residual=res, y_halo=1,
              per_sample=True)
    a = K.conv_fwd(x, wps, cout, k, k, pad, path=K.PATH_TC, **kw)
    b = K.conv_fwd(x, wps, cout, k, k, pad, path=K.PATH_SIMT, **kw)
    assert relerr(a.float(), b.float()) < 4e-3, ("epilogue", case)
    assert torch.equal(K.padded_view(a, 1).float(), F.pad(a.float(), (1,) * 4, mode="reflect")), case


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("per_sample", [False, True])
def test_conv_per_sample_weights(K, dtype, per_sample):
    n, cin, cout, H, W = 3, 64, 128, 16, 16
    x = rnd(n, cin, H, W, seed=7)
    w = rnd(cout, cin, 3, 3, seed=8)
    s = torch.rand(n, cin, device="cuda") + 0.5
    alpha = 1 / math.sqrt(cin * 9)
    wp = K.weight_pack(w, alpha, dtype, cs=s if per_sample else None, nb=n if per_sample else 1)
    y = K.conv_fwd(nhwc(x, dtype), wp, cout, 3, 3, 1, per_sample=per_sample)
    xq = x.to(dtype).float()
    if per_sample:
        ref = torch.cat([
            F.conv2d(xq[i : i + 1], (w * alpha * s[i][None, :, None, None]).to(dtype).float(), padding=1)
            for i in range(n)
        ])
    else:
        ref = F.conv2d(xq, (w * alpha).to(dtype).float(), padding=1)
    assert relerr(y.float(), ref) < TOL[dtype]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("case", CONV_CASES)
def test_conv_dgrad_wgrad(K, case, dtype):
    cin, cout, k, pad, halo, H, W, n = case
    alpha = 1 / math.sqrt(cin * k * k)
    x = rnd(n, cin, H, W, seed=9).to(dtype).float()
    w = rnd(cout, cin, k, k, seed=10)
    xp = (F.pad(x, (halo,) * 4, mode="reflect") if halo else x).requires_grad_(True)
    wr = ((w * alpha).to(dtype).float()).requires_grad_(True)
    y = F.conv2d(xp, wr, padding=pad - halo)
    dy = rnd(*y.shape, seed=11).to(dtype).float()
    gx_ref, gw_ref = torch.autograd.grad(y, (xp, wr), dy)
    # dgrad = conv of dy with the transposed pack, full padding k-1-(pad-halo)
    wpt = K.weight_pack(w, alpha, dtype, transpose=True)
    gx = K.conv_fwd(nhwc(dy, dtype), wpt, cin, k, k, k - 1 - (pad - halo))
    assert gx.shape == gx_ref.shape
    assert relerr(gx.float(), gx_ref) < TOL[dtype], f"dgrad {case}"
    # wgrad (accumulates into dw; returns grad w.r.t. the RAW weight = alpha * dL/d(alpha w))
    dw = torch.full_like(w, 0.5)
    K.conv_wgrad(nhwc(x, dtype, halo), nhwc(dy, dtype), dw, k, k, pad, x_halo=halo, alpha=alpha)
    assert relerr(dw - 0.5, gw_ref * alpha) < TOL[dtype] * 2, f"wgrad {case}"


@pytest.mark.parametrize("k,pad,halo,H,W", [(7, 3, 3, 40, 24), (4, 1, 0, 33, 20), (7, 3, 3, 16, 16)])
def test_image_side_convs_mixed_dtype(K, k, pad, halo, H, W):
    """Cin == 1 kernels: fp32 image in, bf16 activations out; wgrad with fp32 x and bf16 dy."""
    n, cout = 3, 64
    x = rnd(n, 1, H, W, seed=60)
    w = rnd(cout, 1, k, k, seed=61)
    bias = rnd(cout, seed=62)
    alpha = 1 / math.sqrt(k * k)
    xp = F.pad(x, (halo,) * 4, mode="reflect") if halo else x
    ref = F.leaky_relu(F.conv2d(xp, w * alpha, bias, padding=pad - halo), 0.2)
    wp = K.weight_pack(w, alpha, torch.float32)
    for dt in (torch.bfloat16, torch.float32):
        y = K.conv_fwd(nhwc(x, torch.float32, halo), wp, cout, k, k, pad, x_halo=halo, bias=bias,
                       act=K.ACT_LRELU, out_dtype=dt, y_halo=1)
        assert y.dtype == dt and relerr(y.float(), ref) < TOL[dt]
        assert torch.equal(K.padded_view(y, 1).float(), F.pad(y.float(), (1,) * 4, mode="reflect"))
        dy = rnd(*ref.shape, seed=63).to(dt)
        gw_ref = torch.nn.grad.conv2d_weight(xp, w.shape, dy.float(), padding=pad - halo) * alpha
        dw = torch.zeros_like(w)
        K.conv_wgrad(nhwc(x, torch.float32, halo), nhwc(dy.float(), dt), dw, k, k, pad, x_halo=halo,
                     alpha=alpha)
        # bf16 mode runs on mma.sync with the image patch rounded to bf16 (like every other
        # layer's input in that mode); fp32 mode is the FFMA kernel
        assert relerr(dw, gw_ref) < (2e-4 if dt == torch.float32 else 5e-3), (k, dt)


@pytest.mark.parametrize("case", [(128, 128, 16, 16, 3), (64, 128, 16, 24, 2), (128, 64, 16, 16, 2),
                                  (256, 256, 8, 8, 2), (128, 256, 15, 15, 2)])
def test_wgrad_tc_modulated(K, case):
    cin, cout, H, W, n = case
    x = rnd(n, cin, H, W, seed=12).bfloat16().float()
    dy = rnd(n, cout, H, W, seed=13).bfloat16().float()
    rs = torch.rand(n, cout, device="cuda") + 0.5
    cs = torch.rand(n, cin, device="cuda") + 0.5
    ref = torch.zeros(cout, cin, 3, 3, device="cuda")
    for i in range(n):
        xi = (x[i : i + 1] * cs[i][None, :, None, None])
        dyi = dy[i : i + 1] * rs[i][None, :, None, None]
        ref += torch.nn.grad.conv2d_weight(xi, (cout, cin, 3, 3), dyi, padding=1)
    dws = {}
    for path in (K.PATH_TC, K.PATH_SIMT):
        dw = torch.zeros(cout, cin, 3, 3, device="cuda")
        K.conv_wgrad(nhwc(x, torch.bfloat16), nhwc(dy, torch.bfloat16), dw, 3, 3, 1, alpha=0.25,
                     rs=rs, cs=cs, path=path)
        dws[path] = dw
        assert relerr(dw, ref * 0.25) < 5e-3, (case, path)


def test_wgrad_fused_demodulation_term(K):
    """P[n,o] = sum_hw dy*y produced by the tcgen05 wgrad epilogue from the per-sample forward
    pack (SURVEY App. B.2 identity) vs the direct reduction; and the ReLU mask of otm_mod_in."""
    n, c, H, W = 3, 128, 16, 16
    alpha = 1 / math.sqrt(c * 9)
    x = rnd(n, c, H, W, seed=70).bfloat16().float()
    w = rnd(c, c, 3, 3, seed=71)
    s = rnd(n, c, seed=72) * 0.3 + 1
    dy = rnd(n, c, H, W, seed=73).bfloat16().float()
    xt = nhwc(x, torch.bfloat16, 1)
    q = K.weight_sqsum(w, alpha)
    sig = K.demod(s, q)
    wp = K.weight_pack(w, alpha, torch.bfloat16, cs=s, nb=n)
    assert K.wgrad_fuses_P(xt, nhwc(dy, torch.bfloat16), 3, 3, 1, 1)
    y = torch.cat([
        F.conv2d(F.pad(x[i : i + 1], (1,) * 4, mode="reflect"), wp[i].permute(0, 3, 1, 2).float())
        for i in range(n)
    ]) * sig[:, :, None, None]
    P_ref = (dy * y).sum((2, 3))
    dw = torch.zeros_like(w)
    P = torch.zeros(n, c, device="cuda")
    K.conv_wgrad(xt, nhwc(dy, torch.bfloat16), dw, 3, 3, 1, x_halo=1, alpha=alpha, rs=sig, cs=s,
                 wfwd=wp, P=P)
    assert relerr(P, P_ref) < 5e-3
    dw2 = torch.zeros_like(w)
    K.conv_wgrad(xt, nhwc(dy, torch.bfloat16), dw2, 3, 3, 1, x_halo=1, alpha=alpha, rs=sig, cs=s)
    assert relerr(dw, dw2) < 1e-5
    # relu mask fused into the input-side pass
    g = rnd(n, c, H, W, seed=74).bfloat16().float()
    xr = F.relu(rnd(n, c, H, W, seed=75)).bfloat16().float()
    gx, Q = K.mod_in(nhwc(g, torch.bfloat16), nhwc(xr, torch.bfloat16), s, relu_mask=True)
    assert relerr(gx.float(), g * s[:, :, None, None] * (xr > 0)) < 2e-2
    assert relerr(Q, (g * xr).sum((2, 3))) < 2e-2


def test_modulation_coefficients(K):
    n, cin, cout = 3, 64, 128
    w = rnd(cout, cin, 3, 3, seed=14)
    s = rnd(n, cin, seed=15) + 1
    alpha = 1 / math.sqrt(cin * 9)
    q = K.weight_sqsum(w, alpha)
    assert relerr(q, ((w * alpha) ** 2).sum((2, 3))) < 1e-5
    si = K.demod(s, q)
    wts = (w * alpha)[None] * s[:, None, :, None, None]
    ref = torch.rsqrt((wts**2).sum((2, 3, 4)) + 1e-8)
    assert relerr(si, ref) < 1e-5
    # backward of the coefficient path against autograd on the dense identity
    P = rnd(n, cout, seed=16)
    Q = rnd(n, cin, seed=17)
    s2 = s.clone().requires_grad_(True)
    w2 = w.clone().requires_grad_(True)
    q2 = ((w2 * alpha) ** 2).sum((2, 3))
    sig = torch.rsqrt((s2**2) @ q2.t() + 1e-8)
    # dL/dsigma_inv = P / sigma_inv  (P = sum dy*y), plus the direct term Q on s
    loss = (sig * (P / sig.detach())).sum() + (s2 * Q).sum()
    gs, gw = torch.autograd.grad(loss, (s2, w2))
    dw = torch.zeros_like(w)
    ds = K.mod_bwd(w, alpha, s, si, q, P, Q, dw)
    assert relerr(ds, gs) < 1e-4
    assert relerr(dw, gw) < 1e-4


NORM_SHAPES = [(2, 64, 12, 20), (2, 128, 9, 7), (3, 1, 16, 10), (1, 512, 5, 5)]


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", NORM_SHAPES)
@pytest.mark.parametrize("act", [0, 1, 2, 3])
def test_norm_act_fwd_bwd(K, shape, dtype, act):
    n, c, h, w = shape
    x = (rnd(*shape, seed=18) * 2 + 0.5).to(dtype).float().requires_grad_(True)
    res = rnd(*shape, seed=19).to(dtype).float().requires_grad_(True)
    halo = 1 if act != 3 else 3
    fn = [lambda t: t, F.relu, lambda t: F.leaky_relu(t, 0.2), torch.tanh][act]
    yn = F.instance_norm(x, eps=1e-5)
    y = fn(yn) + res
    yp = F.pad(y, (halo,) * 4, mode="reflect")
    g = rnd(*yp.shape, seed=20).to(dtype).float()
    gx_ref, gres_ref = torch.autograd.grad(yp, (x, res), g)

    xt = nhwc(x.detach(), dtype)
    stats = K.instnorm_stats(xt)
    m = x.detach().mean((2, 3))
    v = x.detach().var((2, 3), correction=0)
    assert relerr(stats[..., 0], m) < 1e-4 and relerr(stats[..., 1], torch.rsqrt(v + 1e-5)) < 1e-4
    out = K.norm_act(xt, stats, act, residual=nhwc(res.detach(), dtype), y_halo=halo)
    assert relerr(K.padded_view(out, halo).float(), yp.detach()) < TOL[dtype]
    gp = nhwc(g, dtype)  # padded-size gradient; interior view of it:
    gint = gp[:, :, halo : halo + h, halo : halo + w]
    gx, gres = K.norm_act_bwd(gint, xt, stats, act, g_halo=halo, want_gres=True)
    assert relerr(gres.float(), gres_ref) < TOL[dtype]
    tol = TOL[dtype] * (3 if dtype == torch.float32 else 2)
    assert relerr(gx.float(), gx_ref) < tol


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(2, 64, 16, 24), (2, 64, 31, 15), (1, 128, 7, 9), (2, 1, 33, 20),
                                   (2, 64, 2, 2), (1, 64, 3, 5)])
def test_down_up(K, shape, dtype):
    from oracle import reference_port as rp  # the resampling restatement pinned by the goldens

    kern = rp._smooth_kernel().cuda()
    x = rnd(*shape, seed=21).to(dtype).float().requires_grad_(True)
    # DownSample with fused norm + lrelu prologue
    t = F.leaky_relu(F.instance_norm(x, eps=1e-5), 0.2)
    t.retain_grad()
    y = rp.down_sample(t, kern)
    g = rnd(*y.shape, seed=22).to(dtype).float()
    (gt_ref,) = torch.autograd.grad(y, t, g, retain_graph=True)
    xt = nhwc(x.detach(), dtype)
    stats = K.instnorm_stats(xt)
    out = K.down(xt, stats, K.ACT_LRELU)
    assert relerr(out.float(), y.detach()) < TOL[dtype], "down fwd"
    ga = K.down_bwd(nhwc(g, dtype), shape[2:])
    assert relerr(ga.float(), gt_ref) < TOL[dtype], "down bwd"
    # fused: stencil transpose + activation + instance-norm backward in one call
    if shape[2] * shape[3] > 4:
        (gx_ref,) = torch.autograd.grad(y, x, g, retain_graph=True)
        gx = K.norm_act_bwd(nhwc(g, dtype), xt, stats, K.ACT_LRELU, g_down=True)
        assert relerr(gx.float(), gx_ref) < TOL[dtype] * 3, "fused down+norm bwd"
    # UpSample
    x2 = rnd(*shape, seed=23).to(dtype).float().requires_grad_(True)
    y2 = rp.up_sample(x2, kern)
    g2 = rnd(*y2.shape, seed=24).to(dtype).float()
    (gx2_ref,) = torch.autograd.grad(y2, x2, g2)
    out2 = K.up(nhwc(x2.detach(), dtype))
    assert relerr(out2.float(), y2.detach()) < TOL[dtype], "up fwd"
    gx2 = K.up_bwd(nhwc(g2, dtype))
    assert relerr(gx2.float(), gx2_ref) < TOL[dtype], "up bwd"


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_mod_side_passes(K, dtype):
    n, c, h, w = 2, 64, 10, 12
    out = F.relu(rnd(n, c, h, w, seed=25)).to(dtype).float()
    res = rnd(n, c, h, w, seed=26).to(dtype).float()
    gpad = rnd(n, c, h + 2, w + 2, seed=27).to(dtype).float()
    g2 = rnd(n, c, h, w, seed=28).to(dtype).float()
    # fold of a reflect-padded gradient = autograd of F.pad
    z = torch.zeros(n, c, h, w, device="cuda", requires_grad=True)
    (fold,) = torch.autograd.grad(F.pad(z, (1,) * 4, mode="reflect"), z, gpad)
    ga = fold + g2
    gp = nhwc(gpad, dtype)
    gint = gp[:, :, 1 : 1 + h, 1 : 1 + w]
    gy, P = K.mod_out(gint, nhwc(out, dtype), g_halo=1, g2=nhwc(g2, dtype), act=K.ACT_RELU)
    gy_ref = ga * (out > 0)
    assert relerr(gy.float(), gy_ref) < TOL[dtype]
    assert relerr(P, (gy_ref * out).sum((2, 3))) < TOL[dtype]
    _, P2 = K.mod_out(gint, nhwc(out, dtype), g_halo=1, res=nhwc(res, dtype), materialise=False)
    assert relerr(P2, (fold * (out - res)).sum((2, 3))) < TOL[dtype]
    s = torch.rand(n, c, device="cuda") + 0.5
    x = rnd(n, c, h, w, seed=29).to(dtype).float()
    gx, Q = K.mod_in(gint, nhwc(x, dtype), s, g_halo=1, gadd=nhwc(g2, dtype))
    assert relerr(Q, (fold * x).sum((2, 3))) < TOL[dtype]
    assert relerr(gx.float(), fold * s[:, :, None, None] + g2) < TOL[dtype]
    assert relerr(K.channel_sum(nhwc(x, dtype)), x.sum((0, 2, 3))) < TOL[dtype]
    assert relerr(K.avgpool(nhwc(x, dtype)), x.mean((2, 3))) < TOL[dtype]
    gv = rnd(n, c, seed=30)
    assert relerr(K.avgpool_bwd(gv, (n, c, h, w), dtype).float(),
                  (gv / (h * w))[:, :, None, None].expand(n, c, h, w)) < TOL[dtype]


def test_losses(K):
    s = rnd(4, 1, 13, 9, seed=31).contiguous(memory_format=torch.channels_last)
    out, g = K.loss_lsgan(s, 1.0, scale=0.5)
    sr = s.clone().requires_grad_(True)
    ref = F.mse_loss(sr, torch.ones_like(sr))
    (gr,) = torch.autograd.grad(0.5 * ref, sr)
    assert abs(out[0].item() - ref.item()) < 1e-5 * max(1, abs(ref.item()))
    assert abs(out[1].item() - torch.sign(s * 2 - 1).mean().item()) < 1e-6
    assert relerr(g, gr) < 1e-5
    a, b = rnd(2, 1, 16, 12, seed=32), rnd(2, 1, 16, 12, seed=33)
    out, g = K.loss_l1(a, b, scale=5.0)
    ar = a.clone().requires_grad_(True)
    ref = F.l1_loss(ar, b)
    (gr,) = torch.autograd.grad(5 * ref, ar)
    assert abs(out[0].item() - ref.item()) < 1e-5 and relerr(g, gr) < 1e-6
    for dtype in (torch.float32, torch.bfloat16):
        f1 = rnd(2, 64, 8, 8, seed=34).to(dtype).float()
        f2 = rnd(2, 64, 8, 8, seed=35).to(dtype).float()
        h = torch.tensor([0.11, 0.19], device="cuda")
        o = torch.zeros(1, device="cuda")
        g1, g2 = K.loss_path(nhwc(f1, dtype), nhwc(f2, dtype), h, 0.25, 0.1, o)
        f1r = f1.clone().requires_grad_(True)
        ref = 0.25 * (((f1r - f2) / h[:, None, None, None]) ** 2).mean()
        (gr,) = torch.autograd.grad(0.1 * ref, f1r)
        assert abs(o.item() - ref.item()) < 1e-4 * abs(ref.item())
        assert relerr(g1.float(), gr) < TOL[dtype] and relerr(g2.float(), -gr) < TOL[dtype]
        xm = nhwc(f1, dtype, 1)
        mo = K.moments(xm)
        assert abs(mo[0].item() - f1.sum().item()) < 1e-3 * f1.abs().sum().item()
        assert abs(mo[1].item() - (f1**2).sum().item()) < 1e-4 * (f1**2).sum().item()
        coef = torch.tensor([0.3, -0.7], device="cuda")
        ag = K.affine_grad(xm, coef)
        assert relerr(ag.float(), 0.3 - 0.7 * f1) < TOL[dtype]


def test_adam_matches_torch(K):
    n = 10007
    p0 = rnd(n, seed=36)
    p = p0.clone()
    m = torch.zeros_like(p)
    v = torch.zeros_like(p)
    step = torch.zeros(1, dtype=torch.int32, device="cuda")
    ref = torch.nn.Parameter(p0.clone())
    opt = torch.optim.Adam([ref], lr=2e-3, betas=(0.5, 0.99))
    for it in range(3):
        g = rnd(n, seed=40 + it)
        step += 1
        K.adam(p, g, m, v, step, 2e-3, 0.5, 0.99)
        ref.grad = g.clone()
        opt.step()
    assert relerr(p, ref.detach()) < 1e-6
    assert (p - ref.detach()).abs().max().item() < 1e-6


def test_synth_uniform(K):
    out = torch.empty(1 << 20, device="cuda")
    K.synth_uniform(out, 42, 1)
    assert -1.0 <= out.min().item() and out.max().item() < 1.0
    assert abs(out.mean().item()) < 5e-3 and abs(out.var().item() - 1 / 3) < 5e-3
    again = torch.empty(1 << 20, device="cuda")
    K.synth_uniform(again, 42, 1)
    assert torch.equal(out, again)
    other = torch.empty(1 << 20, device="cuda")
    K.synth_uniform(other, 42, 2)
    assert not torch.equal(out, other)


def test_cast_add(K):
    x = rnd(2, 64, 5, 7, seed=50)
    a = nhwc(x, torch.float32, 1)
    b = K.cast(a, torch.bfloat16)
    assert torch.equal(b.float(), x.bfloat16().float())
    K.add_(a, nhwc(x, torch.float32))
    assert relerr(a, 2 * x) < 1e-6


def test_weight_pack_multi_matches_single():
    """otm_weight_pack_multi: many shared packs in one launch == otm_weight_pack one by one
    (bit-exact: same arithmetic, same rounding)."""
    from one_to_many_gan_b200 import kernels as K

    g = torch.Generator(device="cuda").manual_seed(11)
    shapes = [(128, 128, 3, 3), (64, 128, 3, 3), (128, 64, 4, 4), (256, 128, 4, 4), (64, 8, 7, 7)]
    for dtype in (torch.bfloat16, torch.float32):
        jobs = []
        for i, sh in enumerate(shapes * 15):  # 75 jobs: more than one kernel-parameter table
            w = torch.randn(*sh, generator=g, device="cuda")
            jobs.append((w, 0.1 + 0.01 * i, bool(i % 2)))
        outs = K.weight_pack_multi(jobs, dtype)
        for (w, alpha, tr), out in zip(jobs, outs):
            ref = K.weight_pack(w, alpha, dtype, transpose=tr)
            assert out.shape == ref.shape
            assert torch.equal(out, ref)
