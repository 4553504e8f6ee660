"""Thin Python wrappers over the C-ABI: allocate outputs with torch (plumbing), describe the
tensors, call the kernel on the current CUDA stream.  No arithmetic happens here."""

from __future__ import annotations

import ctypes as C

import torch

from . import _lib as L
from ._lib import ACT_LRELU, ACT_NONE, ACT_RELU, ACT_TANH, PATH_AUTO, PATH_SIMT, PATH_TC  # noqa: F401

_byref = C.byref


def alloc(n: int, c: int, h: int, w: int, dtype, device, halo: int = 0, zero: bool = False):
    """NHWC buffer with `halo` extra pixels on every side; returns the INTERIOR view as a
    logically-NCHW tensor (channel stride 1)."""
    mk = torch.zeros if zero else torch.empty
    buf = mk((n, h + 2 * halo, w + 2 * halo, c), dtype=dtype, device=device)
    if halo:
        buf = buf[:, halo : halo + h, halo : halo + w, :]
    return buf.permute(0, 3, 1, 2)


def padded_view(t: torch.Tensor, halo: int) -> torch.Tensor:
    """The halo-inclusive view [n,c,h+2p,w+2p] of an interior view produced by alloc()."""
    if halo == 0:
        return t
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    off = t.storage_offset() - halo * (sh + sw)
    if off < 0:
        raise ValueError("tensor has no materialised halo")
    return t.as_strided((n, c, h + 2 * halo, w + 2 * halo), (sn, sc, sh, sw), off)


def conv_fwd(x, wpack, cout, kh, kw, pad, *, x_halo=0, y_halo=0, alpha=1.0, row_scale=None,
             bias=None, act=ACT_NONE, residual=None, out_dtype=None, per_sample=False,
             path=PATH_AUTO, out=None, post_scale=None, want_stats=False, eps=1e-5):
    """want_stats: also return the InstanceNorm statistics (mean, rstd) [n, cout, 2] of the output
    -- accumulated in the conv epilogue when the launch supports it, else by a separate pass."""
    n, cin, h, w = x.shape
    ho, wo = h + 2 * pad - kh + 1, w + 2 * pad - kw + 1
    if out is None:
        out = alloc(n, cout, ho, wo, out_dtype or x.dtype, x.device, y_halo)
    a = L.ConvFwdArgs()
    a.x = L.tdesc(x)
    a.x_halo = x_halo
    a.wpack = L.ptr(wpack)
    a.w_batch_stride = cout * kh * kw * cin if per_sample else 0
    a.kh, a.kw, a.pad = kh, kw, pad
    a.y = L.tdesc(out)
    a.y_halo = y_halo
    a.alpha = alpha
    a.row_scale = L.ptr(row_scale)
    a.bias = L.ptr(bias)
    a.act = act
    a.residual = L.tdesc(residual)
    a.path = path
    a.post_scale = L.ptr(post_scale)
    if not want_stats:
        L.check(L.lib.otm_conv_fwd(_byref(a), L.stream_ptr()), "otm_conv_fwd")
        return out
    if y_halo == 0 and L.lib.otm_conv_fwd_fuses_stats(_byref(a)):
        sums = torch.empty((n, cout, 2), dtype=torch.float32, device=x.device)
        stats = torch.empty((n, cout, 2), dtype=torch.float32, device=x.device)
        a.stat_sums = L.ptr(sums)
        L.check(L.lib.otm_conv_fwd(_byref(a), L.stream_ptr()), "otm_conv_fwd")
        L.check(L.lib.otm_instnorm_finalize(L.ptr(sums), L.ptr(stats), n * cout, ho * wo, eps,
                                            L.stream_ptr()), "otm_instnorm_finalize")
        return out, stats
    L.check(L.lib.otm_conv_fwd(_byref(a), L.stream_ptr()), "otm_conv_fwd")
    return out, instnorm_stats(out, eps)


def conv_fwd_dot(x, wpack, cout, pad, aux):
    """3x3 conv_fwd (shared pack) that also returns dot[n, o] = sum_hw y * aux from its epilogue
    (residual_mode 2), or None when this launch cannot fuse it."""
    n, cin, h, w = x.shape
    out = alloc(n, cout, h + 2 * pad - 2, w + 2 * pad - 2, x.dtype, x.device)
    a = L.ConvFwdArgs()
    a.x = L.tdesc(x)
    a.wpack = L.ptr(wpack)
    a.kh, a.kw, a.pad = 3, 3, pad
    a.y = L.tdesc(out)
    a.alpha = 1.0
    a.act = ACT_NONE
    a.path = PATH_AUTO
    if x.dtype != torch.bfloat16 or not L.lib.otm_conv_fwd_fuses_gate(_byref(a)):
        return None
    dot = torch.empty((n, cout), dtype=torch.float32, device=x.device)
    a.residual = L.tdesc(aux)
    a.residual_mode = 2
    a.dot_sums = L.ptr(dot)
    L.check(L.lib.otm_conv_fwd(_byref(a), L.stream_ptr()), "otm_conv_fwd")
    return out, dot


def _dgrad_reflect_args(g, wpack_t, cin, per_sample, out):
    a = L.ConvFwdArgs()
    a.x = L.tdesc(g)
    a.x_halo = 0
    a.wpack = L.ptr(wpack_t)
    a.w_batch_stride = cin * 9 * g.shape[1] if per_sample else 0
    a.kh, a.kw, a.pad = 3, 3, 1
    a.y = L.tdesc(out)
    a.y_halo = 0
    a.alpha = 1.0
    a.act = ACT_NONE
    a.path = PATH_AUTO
    return a


def dgrad_reflect_ok(g, cin) -> bool:
    """Can conv_dgrad_reflect run for this upstream gradient?  (bf16 on the tcgen05 path, channel
    counts the border kernel tiles, an image the reflect pad is defined on)"""
    n, k, h, w = g.shape
    return g.dtype == torch.bfloat16 and k % 64 == 0 and cin % 64 == 0 and h >= 3 and w >= 3


def conv_dgrad_reflect(g, wpack_t, cin, *, per_sample=False, residual=None, gate=None,
                       row_scale=None, post_scale=None, want_dot=False):
    """Gradient of conv3x3(ReflectionPad2d(1)(x)) w.r.t. x, [n, cin, H, W], WITHOUT the padded
    (H+2) x (W+2) intermediate: the zero-padded "same" dgrad on the tcgen05 path plus
    otm_conv_reflect_border for the halo ring (reference blocks.py:21-27,49-56 backward).
    wpack_t: the dgrad pack (weight_pack(..., transpose=True)).
      residual: added to the result (the block's skip gradient);
      gate: (style-scaled) ReLU output the gradient is taken w.r.t. -- result = (gate != 0) * row_scale *
            post_scale * dgrad, and with want_dot also dot[n, c] = sum_hw dgrad * gate
            (the otm_mod_in pass of a modulated conv, done in the dgrad epilogue).
    Returns y or (y, dot).  The caller checks dgrad_reflect_ok / dgrad_reflect_fuses_gate."""
    n, k, h, w = g.shape
    out = alloc(n, cin, h, w, g.dtype, g.device)
    a = _dgrad_reflect_args(g, wpack_t, cin, per_sample, out)
    a.row_scale = L.ptr(row_scale)
    a.post_scale = L.ptr(post_scale)
    dot = None
    if gate is not None:
        a.residual = L.tdesc(gate)
        a.residual_mode = 1
        if want_dot:
            dot = torch.empty((n, cin), dtype=torch.float32, device=g.device)
            a.dot_sums = L.ptr(dot)
    else:
        a.residual = L.tdesc(residual)
    L.check(L.lib.otm_conv_fwd(_byref(a), L.stream_ptr()), "otm_conv_fwd")
    b = L.ConvReflectBorderArgs()
    b.dy = L.tdesc(g)
    b.wpack = L.ptr(wpack_t)
    b.w_batch_stride = a.w_batch_stride
    b.y = L.tdesc(out)
    b.row_scale = L.ptr(row_scale)
    b.post_scale = L.ptr(post_scale)
    b.gate = L.tdesc(gate)
    b.dot_sums = L.ptr(dot)
    L.check(L.lib.otm_conv_reflect_border(_byref(b), L.stream_ptr()), "otm_conv_reflect_border")
    return (out, dot) if want_dot else out


def dgrad_reflect_fuses_gate(g, wpack_t, cin, per_sample=False) -> bool:
    """Would the main launch of conv_dgrad_reflect run on the kernel that implements the gate?"""
    if not dgrad_reflect_ok(g, cin):
        return False
    n, k, h, w = g.shape
    a = _dgrad_reflect_args(g, wpack_t, cin, per_sample, g)
    # only the geometry / dtype / alignment of y decide the kernel: describe a dense y from g
    a.y.c, a.y.sw, a.y.sh, a.y.sn = cin, cin, cin * w, cin * w * h
    return bool(L.lib.otm_conv_fwd_fuses_gate(_byref(a)))


def conv_wgrad(x, dy, dw, kh, kw, pad, *, x_halo=0, alpha=1.0, rs=None, cs=None, path=PATH_AUTO,
               use_ws=True, wfwd=None, P=None, wfwd_per_sample=True, Q=None):
    """P (zeroed [n, Cout] fp32) + wfwd (forward pack: per-sample, or the shared one when x
    already carries the modulation): also accumulate the demodulation term sum_hw dy*y in the
    epilogue (check wgrad_fuses_P first)."""
    a = L.ConvWgradArgs()
    a.x = L.tdesc(x)
    a.x_halo = x_halo
    a.dy = L.tdesc(dy)
    a.kh, a.kw, a.pad = kh, kw, pad
    a.dw = L.ptr(dw)
    a.alpha = alpha
    a.rs = L.ptr(rs)
    a.cs = L.ptr(cs)
    a.path = path
    ws = None
    if use_ws:
        nbytes = L.lib.otm_conv_wgrad_workspace_bytes(_byref(a))
        if nbytes:
            ws = torch.empty(nbytes // 4, dtype=torch.float32, device=dw.device)
    a.ws = L.ptr(ws)
    a.wfwd = L.ptr(wfwd)
    a.P = L.ptr(P)
    a.Q = L.ptr(Q)
    a.wfwd_batch_stride = dw.numel() if (wfwd is not None and wfwd_per_sample) else 0
    L.check(L.lib.otm_conv_wgrad(_byref(a), L.stream_ptr()), "otm_conv_wgrad")
    return dw


def _lib_uses_tc_fwd(x, wpack, cout, k, pad, x_halo=0) -> bool:
    """Would otm_conv_fwd take the tcgen05 path for this (bf16) input?  (tests)"""
    n, cin, h, w = x.shape
    y = alloc(n, cout, h + 2 * pad - k + 1, w + 2 * pad - k + 1, x.dtype, x.device)
    a = L.ConvFwdArgs()
    a.x = L.tdesc(x)
    a.x_halo = x_halo
    a.wpack = L.ptr(wpack)
    a.kh, a.kw, a.pad = k, k, pad
    a.y = L.tdesc(y)
    a.path = PATH_AUTO
    return bool(L.lib.otm_conv_fwd_uses_tcgen05(_byref(a)))


def _lib_uses_tc_wgrad(x, dy, k, pad, x_halo=0) -> bool:
    a = L.ConvWgradArgs()
    a.x = L.tdesc(x)
    a.x_halo = x_halo
    a.dy = L.tdesc(dy)
    a.kh, a.kw, a.pad = k, k, pad
    a.path = PATH_AUTO
    return bool(L.lib.otm_conv_wgrad_uses_tcgen05(_byref(a)))


def wgrad_fuses_P(x, dy, kh, kw, pad, x_halo=0) -> bool:
    a = L.ConvWgradArgs()
    a.x = L.tdesc(x)
    a.x_halo = x_halo
    a.dy = L.tdesc(dy)
    a.kh, a.kw, a.pad = kh, kw, pad
    a.path = PATH_AUTO
    return bool(L.lib.otm_conv_wgrad_fuses_P(_byref(a)))


def wgrad_fuses_Q(x, dy, kh, kw, pad, x_halo=0) -> bool:
    """Can conv_wgrad(..., Q=) reduce the direct style-gradient term in its epilogue (rs=None)?"""
    a = L.ConvWgradArgs()
    a.x = L.tdesc(x)
    a.x_halo = x_halo
    a.dy = L.tdesc(dy)
    a.kh, a.kw, a.pad = kh, kw, pad
    a.path = PATH_AUTO
    return bool(L.lib.otm_conv_wgrad_fuses_Q(_byref(a)))


def weight_pack(w, alpha, dtype, *, cs=None, rs=None, nb=1, transpose=False):
    cout, cin, kh, kw = w.shape
    out = torch.empty((nb, cin if transpose else cout, kh, kw, cout if transpose else cin),
                      dtype=dtype, device=w.device)
    a = L.WeightPackArgs()
    a.w = L.ptr(w)
    a.cout, a.cin, a.kh, a.kw = cout, cin, kh, kw
    a.alpha = alpha
    a.cs = L.ptr(cs)
    a.rs = L.ptr(rs)
    a.nb = nb
    a.transpose = int(transpose)
    a.out = L.ptr(out)
    a.out_dtype = L.dtype_code(out)
    L.check(L.lib.otm_weight_pack(_byref(a), L.stream_ptr()), "otm_weight_pack")
    return out


def weight_pack_multi_ok(w, transpose: bool) -> bool:
    cout, cin = w.shape[:2]
    return (cout if transpose else cin) % 8 == 0


def weight_pack_multi(jobs, dtype):
    """jobs: [(w [Cout,Cin,kh,kw] fp32, alpha, transpose)] -> the shared packs (as weight_pack
    returns them, [1, ...]) built by ONE launch into one buffer."""
    es = 2 if dtype == torch.bfloat16 else 4
    offs, total = [], 0
    for w, _, _ in jobs:
        offs.append(total)
        total += (w.numel() * es + 255) // 256 * 256  # 256-byte aligned slices
    dev = jobs[0][0].device
    flat = torch.empty(total, dtype=torch.uint8, device=dev)
    arr = (L.WeightPackArgs * len(jobs))()
    outs = []
    for k, (w, alpha, transpose) in enumerate(jobs):
        cout, cin, kh, kw = w.shape
        shape = (1, cin if transpose else cout, kh, kw, cout if transpose else cin)
        out = flat[offs[k] : offs[k] + w.numel() * es].view(dtype).view(shape)
        a = arr[k]
        a.w = L.ptr(w)
        a.cout, a.cin, a.kh, a.kw = cout, cin, kh, kw
        a.alpha = alpha
        a.nb = 1
        a.transpose = int(transpose)
        a.out = L.ptr(out)
        a.out_dtype = L.dtype_code(out)
        outs.append(out)
    L.check(L.lib.otm_weight_pack_multi(arr, len(jobs), L.stream_ptr()), "otm_weight_pack_multi")
    return outs


def weight_sqsum(w, alpha):
    cout, cin, kh, kw = w.shape
    q = torch.empty((cout, cin), dtype=torch.float32, device=w.device)
    L.check(L.lib.otm_weight_sqsum(L.ptr(w), cout, cin, kh * kw, alpha, L.ptr(q), L.stream_ptr()),
            "otm_weight_sqsum")
    return q


def demod(s, q, eps=1e-8):
    nb, cin = s.shape
    cout = q.shape[0]
    out = torch.empty((nb, cout), dtype=torch.float32, device=s.device)
    L.check(L.lib.otm_demod(L.ptr(s), L.ptr(q), nb, cout, cin, eps, L.ptr(out), L.stream_ptr()),
            "otm_demod")
    return out


def mod_bwd(w, alpha, s, sigma_inv, q, P, Q, dw, q_scaled=False):
    cout, cin, kh, kw = w.shape
    ds = torch.empty_like(s)
    a = L.ModBwdArgs()
    a.w = L.ptr(w)
    a.cout, a.cin, a.taps = cout, cin, kh * kw
    a.alpha = alpha
    a.s, a.sigma_inv, a.q, a.P, a.Q = L.ptr(s), L.ptr(sigma_inv), L.ptr(q), L.ptr(P), L.ptr(Q)
    a.nb = s.shape[0]
    a.ds = L.ptr(ds)
    a.dw = L.ptr(dw)
    a.q_scaled = int(q_scaled)
    L.check(L.lib.otm_mod_bwd(_byref(a), L.stream_ptr()), "otm_mod_bwd")
    return ds


def instnorm_stats(x, eps=1e-5, out=None):
    n, c = x.shape[:2]
    ws = torch.empty((n, c, 2), dtype=torch.float32, device=x.device)
    stats = out if out is not None else torch.empty((n, c, 2), dtype=torch.float32, device=x.device)
    d = L.tdesc(x)
    L.check(L.lib.otm_instnorm_stats(_byref(d), eps, L.ptr(ws), L.ptr(stats), L.stream_ptr()),
            "otm_instnorm_stats")
    return stats


def norm_act(x, stats=None, act=ACT_NONE, residual=None, y_halo=0, out=None):
    n, c, h, w = x.shape
    if out is None:
        out = alloc(n, c, h, w, x.dtype, x.device, y_halo)
    a = L.NormActArgs()
    a.x = L.tdesc(x)
    a.stats = L.ptr(stats)
    a.act = act
    a.residual = L.tdesc(residual)
    a.y = L.tdesc(out)
    a.y_halo = y_halo
    L.check(L.lib.otm_norm_act(_byref(a), L.stream_ptr()), "otm_norm_act")
    return out


def norm_act_bwd(g, x, stats=None, act=ACT_NONE, *, g_halo=0, g2=None, want_gres=False,
                 g_down=False):
    """g: gradient w.r.t. the forward output; if g_halo>0 it is the INTERIOR view of the
    gradient w.r.t. the reflect-padded output.  g_down: g is the half-resolution gradient of
    DownSample(act(norm(x))); the stencil transpose is applied on load."""
    n, c, h, w = x.shape if g_down else g.shape
    gx = alloc(n, c, h, w, g.dtype, g.device)
    gres = alloc(n, c, h, w, g.dtype, g.device) if want_gres else None
    sums = torch.empty((n, c, 2), dtype=torch.float32, device=g.device) if stats is not None else None
    a = L.NormActBwdArgs()
    a.g = L.tdesc(g)
    a.g_halo = g_halo
    a.g2 = L.tdesc(g2)
    a.x = L.tdesc(x)
    a.stats = L.ptr(stats)
    a.act = act
    a.gx = L.tdesc(gx)
    a.gres = L.tdesc(gres)
    a.sums = L.ptr(sums)
    a.g_down = int(g_down)
    L.check(L.lib.otm_norm_act_bwd(_byref(a), L.stream_ptr()), "otm_norm_act_bwd")
    return (gx, gres) if want_gres else gx


def norm_act_bwd_bwd(g, gg, x, stats, act, want_dg=True, want_dx=True):
    """Second derivative of norm_act (R1 penalty): returns (dL/dg, dL/dx) given gg = dL/d(gx)."""
    n, c, h, w = x.shape
    dg = alloc(n, c, h, w, x.dtype, x.device) if want_dg else None
    dx = alloc(n, c, h, w, x.dtype, x.device) if want_dx else None
    sums = torch.empty((n, c, 5), dtype=torch.float32, device=x.device)
    tg, tgg, tx, tdg, tdx = L.tdesc(g), L.tdesc(gg), L.tdesc(x), L.tdesc(dg), L.tdesc(dx)
    L.check(L.lib.otm_norm_act_bwd_bwd(_byref(tg), _byref(tgg), _byref(tx), L.ptr(stats), act, _byref(tdg),
                                       _byref(tdx), L.ptr(sums), L.stream_ptr()), "otm_norm_act_bwd_bwd")
    return dg, dx


def down(x, stats=None, act=ACT_NONE, y_halo=0, out=None):
    n, c, h, w = x.shape
    if out is None:
        out = alloc(n, c, h // 2, w // 2, x.dtype, x.device, y_halo)
    a = L.DownArgs()
    a.x = L.tdesc(x)
    a.stats = L.ptr(stats)
    a.act = act
    a.y = L.tdesc(out)
    a.y_halo = y_halo
    L.check(L.lib.otm_down(_byref(a), L.stream_ptr()), "otm_down")
    return out


def down_bwd(g, in_hw, g_halo=0):
    n, c = g.shape[:2]
    ga = alloc(n, c, in_hw[0], in_hw[1], g.dtype, g.device)
    dg, dga = L.tdesc(g), L.tdesc(ga)
    L.check(L.lib.otm_down_bwd(_byref(dg), g_halo, _byref(dga), L.stream_ptr()), "otm_down_bwd")
    return ga


def up(x, y_halo=0, scale=None):
    """scale: [n, c] fp32 factor on the result (the consumer modconv's style scale)."""
    n, c, h, w = x.shape
    out = alloc(n, c, 2 * h, 2 * w, x.dtype, x.device, y_halo)
    dx, dy = L.tdesc(x), L.tdesc(out)
    L.check(L.lib.otm_up(_byref(dx), _byref(dy), y_halo, L.ptr(scale), L.stream_ptr()), "otm_up")
    return out


def up_bwd(g, g_halo=0, scale=None):
    n, c, h, w = g.shape
    gx = alloc(n, c, h // 2, w // 2, g.dtype, g.device)
    dg, dgx = L.tdesc(g), L.tdesc(gx)
    L.check(L.lib.otm_up_bwd(_byref(dg), g_halo, _byref(dgx), L.ptr(scale), L.stream_ptr()), "otm_up_bwd")
    return gx


def mod_out(g, out, *, g_halo=0, g2=None, res=None, act=ACT_NONE, materialise=True, gy_scale=None):
    n, c, h, w = out.shape
    gy = alloc(n, c, h, w, out.dtype, out.device) if materialise else None
    P = torch.empty((n, c), dtype=torch.float32, device=out.device)
    a = L.ModOutArgs()
    a.g = L.tdesc(g)
    a.g_halo = g_halo
    a.g2 = L.tdesc(g2)
    a.out = L.tdesc(out)
    a.res = L.tdesc(res)
    a.act = act
    a.gy = L.tdesc(gy)
    a.P = L.ptr(P)
    a.gy_scale = L.ptr(gy_scale)
    L.check(L.lib.otm_mod_out(_byref(a), L.stream_ptr()), "otm_mod_out")
    return gy, P


def mod_in(g, x, s, *, g_halo=0, gadd=None, relu_mask=False, gx_scale=None, want_gx=True):
    """want_gx=False: only the reduction Q[n,c] = sum_hw fold(g) * x."""
    n, c, h, w = x.shape
    gx = alloc(n, c, h, w, x.dtype, x.device) if want_gx else None
    Q = torch.empty((n, c), dtype=torch.float32, device=x.device)
    a = L.ModInArgs()
    a.g = L.tdesc(g)
    a.g_halo = g_halo
    a.x = L.tdesc(x)
    a.s = L.ptr(s)
    a.gadd = L.tdesc(gadd)
    a.gx = L.tdesc(gx)
    a.Q = L.ptr(Q)
    a.relu_mask = int(relu_mask)
    a.gx_scale = L.ptr(gx_scale)
    L.check(L.lib.otm_mod_in(_byref(a), L.stream_ptr()), "otm_mod_in")
    return gx, Q


def channel_sum(g, out=None):
    """out given: ACCUMULATE into it (a bias gradient view of the arena)."""
    acc = out is not None
    if out is None:
        out = torch.empty((g.shape[1],), dtype=torch.float32, device=g.device)
    d = L.tdesc(g)
    L.check(L.lib.otm_channel_sum(_byref(d), L.ptr(out), int(acc), L.stream_ptr()), "otm_channel_sum")
    return out


def avgpool(x):
    out = torch.empty(x.shape[:2], dtype=torch.float32, device=x.device)
    d = L.tdesc(x)
    L.check(L.lib.otm_avgpool(_byref(d), L.ptr(out), L.stream_ptr()), "otm_avgpool")
    return out


def avgpool_bwd(g, shape, dtype):
    n, c, h, w = shape
    gx = alloc(n, c, h, w, dtype, g.device)
    d = L.tdesc(gx)
    L.check(L.lib.otm_avgpool_bwd(L.ptr(g), _byref(d), L.stream_ptr()), "otm_avgpool_bwd")
    return gx


def loss_lsgan(x, target, scale=1.0, want_grad=True):
    out = torch.empty((2,), dtype=torch.float32, device=x.device)
    grad = alloc(*_nchw(x), x.dtype, x.device) if want_grad else None
    dx, dg = L.tdesc(x), L.tdesc(grad)
    L.check(L.lib.otm_loss_lsgan(_byref(dx), target, scale, L.ptr(out), _byref(dg), L.stream_ptr()),
            "otm_loss_lsgan")
    return out, grad


def _nchw(x):
    n, c, h, w = x.shape
    return n, c, h, w


def loss_l1(a, b, scale=1.0, want_grad=True):
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    grad = alloc(*_nchw(a), a.dtype, a.device) if want_grad else None
    da, db, dg = L.tdesc(a), L.tdesc(b), L.tdesc(grad)
    L.check(L.lib.otm_loss_l1(_byref(da), _byref(db), scale, L.ptr(out), _byref(dg), L.stream_ptr()),
            "otm_loss_l1")
    return out, grad


def moments(x):
    out = torch.empty((2,), dtype=torch.float32, device=x.device)
    d = L.tdesc(x)
    L.check(L.lib.otm_moments(_byref(d), L.ptr(out), L.stream_ptr()), "otm_moments")
    return out


def affine_grad(x, coef, grad=None):
    accumulate = grad is not None
    if grad is None:
        grad = alloc(*_nchw(x), x.dtype, x.device)
    dx, dg = L.tdesc(x), L.tdesc(grad)
    L.check(L.lib.otm_affine_grad(_byref(dx), L.ptr(coef), _byref(dg), int(accumulate), L.stream_ptr()),
            "otm_affine_grad")
    return grad


def loss_path(f1, f2, h, weight, scale, out, want_grad=True, g1=None, g2=None):
    if want_grad and g1 is None:
        g1 = alloc(*_nchw(f1), f1.dtype, f1.device)
        g2 = alloc(*_nchw(f1), f1.dtype, f1.device)
    d1, d2, dg1, dg2 = L.tdesc(f1), L.tdesc(f2), L.tdesc(g1), L.tdesc(g2)
    L.check(L.lib.otm_loss_path(_byref(d1), _byref(d2), L.ptr(h), weight, scale, L.ptr(out),
                                _byref(dg1), _byref(dg2), L.stream_ptr()), "otm_loss_path")
    return g1, g2


def _rows(t):
    """(pointer, row stride) of a [n, k] fp32 tensor whose last stride is 1 (n may be a stride-0
    broadcast)."""
    if t.dtype != torch.float32 or t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError(f"expected fp32 [n, k] rows, got {t.dtype} {tuple(t.shape)} {t.stride()}")
    return L.ptr(t), t.stride(0)


def linear_jobs(specs):
    """specs: dicts with x [n,k], w [o,k], bias|None, and y (forward) or dy + dw/dbias/dx
    (backward).  Returns a ctypes array of otm_linear_job."""
    arr = (L.LinearJob * len(specs))()
    for j, sp in zip(arr, specs):
        x, w = sp["x"], sp["w"]
        j.x, j.x_row_stride = _rows(x)
        j.w = L.ptr(w)
        j.bias = L.ptr(sp.get("bias"))
        j.y = L.ptr(sp.get("y"))
        j.dy = L.ptr(sp.get("dy"))
        j.dw = L.ptr(sp.get("dw"))
        j.dbias = L.ptr(sp.get("dbias"))
        dx = sp.get("dx")
        if dx is not None:
            j.dx, j.dx_row_stride = _rows(dx)
        j.n, j.k, j.o = x.shape[0], x.shape[1], w.shape[0]
    return arr


def linear_fwd(specs):
    for i in range(0, len(specs), L.MAX_LINEAR_JOBS):
        chunk = specs[i : i + L.MAX_LINEAR_JOBS]
        L.check(L.lib.otm_linear_fwd(linear_jobs(chunk), len(chunk), L.stream_ptr()), "otm_linear_fwd")


def linear_bwd(specs):
    for i in range(0, len(specs), L.MAX_LINEAR_JOBS):
        chunk = specs[i : i + L.MAX_LINEAR_JOBS]
        L.check(L.lib.otm_linear_bwd(linear_jobs(chunk), len(chunk), L.stream_ptr()), "otm_linear_bwd")


def _mapping_args(z1, z2, cross, weights, biases, n_blocks, d, d_const):
    a = L.MappingArgs()
    feats = z1.shape[1]
    if z1.dtype != torch.float32 or not z1.is_contiguous() or (z2 is not None and not z2.is_contiguous()):
        raise ValueError("mapping: z must be contiguous fp32 [batch, features]")
    a.z1, a.z2, a.cross = L.ptr(z1), L.ptr(z2), L.ptr(cross)
    for i, (w, b) in enumerate(zip(weights, biases)):
        a.w[i], a.b[i] = L.ptr(w), L.ptr(b)
    a.features, a.n_layers, a.batch, a.n_blocks = feats, len(weights), z1.shape[0], n_blocks
    for j in range(2):
        a.d[j] = L.ptr(d[j]) if d[j] is not None else None
        a.d_const[j] = float(d_const[j])
    return a


def mapping_fwd(z1, z2, cross, weights, biases, n_blocks, d=(None, None), d_const=(1.0, 1.0),
                n_out=1):
    """-> list of n_out tensors [n_blocks, batch, features]."""
    a = _mapping_args(z1, z2, cross, weights, biases, n_blocks, d, d_const)
    outs = [torch.empty((n_blocks, z1.shape[0], z1.shape[1]), dtype=torch.float32, device=z1.device)
            for _ in range(n_out)]
    for j, o in enumerate(outs):
        a.out[j] = L.ptr(o)
    L.check(L.lib.otm_mapping_fwd(_byref(a), L.stream_ptr()), "otm_mapping_fwd")
    return outs


def mapping_bwd(z1, z2, cross, weights, biases, n_blocks, d, d_const, douts, dws, dbs):
    """Accumulates the parameter gradients into dws / dbs."""
    a = _mapping_args(z1, z2, cross, weights, biases, n_blocks, d, d_const)
    for i, (dw, db) in enumerate(zip(dws, dbs)):
        a.dw[i], a.db[i] = L.ptr(dw), L.ptr(db)
    for j, g in enumerate(douts):
        a.dout[j] = L.ptr(g) if g is not None else None
    L.check(L.lib.otm_mapping_bwd(_byref(a), L.stream_ptr()), "otm_mapping_bwd")


def loss_style_cycle(a, b, ratio=0.2, scale=1.0, want_grad=(True, True)):
    """a, b: fp32 [batch, features] rows.  Returns (loss[1], da|None, db|None)."""
    pa, sa = _rows(a)
    pb, sb = _rows(b)
    out = torch.empty((1,), dtype=torch.float32, device=a.device)
    da = torch.empty(a.shape, dtype=torch.float32, device=a.device) if want_grad[0] else None
    db = torch.empty(b.shape, dtype=torch.float32, device=a.device) if want_grad[1] else None
    L.check(L.lib.otm_loss_style_cycle(pa, sa, pb, sb, a.shape[0], a.shape[1], ratio, scale, L.ptr(out),
                                       L.ptr(da), L.ptr(db), L.stream_ptr()), "otm_loss_style_cycle")
    return out, da, db


def adam(param, grad, m, v, step_dev, lr, beta1, beta2, eps=1e-8, grad_scale=1.0):
    a = L.AdamArgs()
    a.param, a.grad, a.m, a.v = L.ptr(param), L.ptr(grad), L.ptr(m), L.ptr(v)
    a.n = param.numel()
    a.lr, a.beta1, a.beta2, a.eps, a.grad_scale = lr, beta1, beta2, eps, grad_scale
    a.step = L.ptr(step_dev)
    L.check(L.lib.otm_adam(_byref(a), L.stream_ptr()), "otm_adam")


def synth_uniform(out, seed, stream_id, offset=0):
    L.check(L.lib.otm_synth_uniform(L.ptr(out), out.numel(), seed, stream_id, offset, L.stream_ptr()),
            "otm_synth_uniform")
    return out


def gather_batch(data, idx, flip, out=None):
    """data: uint8 [N,C,H,W] on the device; idx int64 [B]; flip uint8 [B] or None -> fp32 [B,C,H,W]
    in [-1, 1]."""
    n, c, h, w = data.shape
    if data.dtype != torch.uint8 or not data.is_contiguous():
        raise ValueError("gather_batch: data must be a contiguous uint8 [N,C,H,W] tensor")
    if out is None:
        out = torch.empty((idx.shape[0], c, h, w), dtype=torch.float32, device=data.device)
    L.check(L.lib.otm_gather_batch(L.ptr(data), n, c, h, w, L.ptr(idx), L.ptr(flip), idx.shape[0],
                                   L.ptr(out), L.stream_ptr()), "otm_gather_batch")
    return out


def cast(x, dtype, out=None):
    if out is None:
        out = alloc(*_nchw(x), dtype, x.device)
    dx, dy = L.tdesc(x), L.tdesc(out)
    L.check(L.lib.otm_cast(_byref(dx), _byref(dy), L.stream_ptr()), "otm_cast")
    return out


def add_(dst, src):
    dd, ds = L.tdesc(dst), L.tdesc(src)
    L.check(L.lib.otm_add_inplace(_byref(dd), _byref(ds), L.stream_ptr()), "otm_add_inplace")
    return dst


def launch_count() -> int:
    return int(L.lib.otm_launch_count())
