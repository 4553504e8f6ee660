"""Host-side check of the decomposition behind otm_conv_reflect_border (csrc/conv_border.cu):

    d/dx conv3x3(ReflectionPad2d(1)(x))  ==  "same" zero-padded dgrad  +  halo-ring correction

The ring is restated here with the kernel's own index logic -- four border lines, the pack tap
each line sees, the line-coordinate shift, the reflected target -- in plain torch on the CPU, and
compared with autograd through F.pad(mode="reflect") + F.conv2d (reference blocks.py:21-27).
The CUDA kernel itself is checked against the same autograd result in test_kernels_gpu.py."""

import pytest
import torch
import torch.nn.functional as F


def dgrad_pack(w):
    """otm_weight_pack(transpose=1): out[i][2-r][2-s][o] = w[o][i][r][s]."""
    return w.flip(2, 3).permute(1, 2, 3, 0).contiguous()  # [Cin][r'][s'][Cout]


def ring_correction(dy, wp):
    """dy [K,H,W], wp [C][3][3][K] (the dgrad pack) -> correction [C,H,W] added to the same-dgrad."""
    K_, H, W = dy.shape
    C = wp.shape[0]
    out = torch.zeros(C, H, W, dtype=dy.dtype)
    for line in range(4):
        horiz = line < 2
        L = W + 2 if horiz else H
        ext = W if horiz else H
        shift = 0 if horiz else 1
        for pos in range(L):
            val = torch.zeros(C, dtype=dy.dtype)
            for t in range(3):
                q = pos + t - 2 + shift  # line coordinate read by tap t
                if not 0 <= q < ext:
                    continue
                hh = 0 if line == 0 else H - 1 if line == 1 else q
                ww = 0 if line == 2 else W - 1 if line == 3 else q
                tap = (6 + t, t, 3 * t + 2, 3 * t)[line]
                val += wp[:, tap // 3, tap % 3, :] @ dy[:, hh, ww]
            if horiz:
                b = pos - 1
                ta = 1 if line == 0 else H - 2
                tb = 1 if b < 0 else (W - 2 if b >= W else b)
            else:
                ta, tb = pos, (1 if line == 2 else W - 2)
            out[:, ta, tb] += val
    return out


@pytest.mark.parametrize("H,W", [(8, 8), (5, 9), (3, 3), (3, 7)])
def test_same_dgrad_plus_ring_equals_reflect_pad_backward(H, W):
    torch.manual_seed(0)
    cin, cout = 4, 6
    x = torch.randn(1, cin, H, W, dtype=torch.float64, requires_grad=True)
    w = torch.randn(cout, cin, 3, 3, dtype=torch.float64)
    y = F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), w)
    dy = torch.randn_like(y)
    (gx_ref,) = torch.autograd.grad(y, x, dy)
    wp = dgrad_pack(w)  # [cin][3][3][cout]
    # main launch: y'[o'] = sum_{r',s',i'} dy[h+r'-1, w+s'-1, i'] wp[o'][r'][s'][i'], zero padding 1
    same = F.conv2d(dy, wp.permute(0, 3, 1, 2), padding=1)
    got = same[0] + ring_correction(dy[0], wp)
    assert torch.allclose(got, gx_ref[0], atol=1e-12)


def test_style_gradient_terms_are_reductions_of_the_per_sample_weight_gradient():
    """The identities behind otm_conv_wgrad_args.P / .Q, in fp64 against
    autograd on the reference's formulation of a modulated conv (layers.py:145-182):
        u[n,o] = sigma[n,o] * conv(reflpad(s[n,i] * x[n,i]), cW)[o]
        G_n[o,i,t] = sum_hw dy[n,o,hw] * xpad[n,i,hw+t]        (what the wgrad kernel accumulates)
        Q[n,i] = sum_{o,t} sigma[n,o] * G_n[o,i,t] * cW[o,i,t]  == dL/ds[n,i] at fixed sigma
        P[n,o] = sigma[n,o] * sum_{i,t} G_n[o,i,t] * cW[o,i,t] * s[n,i]  == sum_hw dy * u
    so neither needs a pass over the activations once G_n sits in the accumulator."""
    torch.manual_seed(1)
    n, cin, cout, H, W = 3, 4, 5, 6, 7
    x = torch.randn(n, cin, H, W, dtype=torch.float64)
    w = torch.randn(cout, cin, 3, 3, dtype=torch.float64)
    s = torch.randn(n, cin, dtype=torch.float64, requires_grad=True)
    sigma = torch.rand(n, cout, dtype=torch.float64) + 0.5
    dy = torch.randn(n, cout, H, W, dtype=torch.float64)
    xp = F.pad(x, (1, 1, 1, 1), mode="reflect")
    u = F.conv2d(xp * s[:, :, None, None], w) * sigma[:, :, None, None]
    (ds,) = torch.autograd.grad(u, s, dy)
    G = torch.stack([torch.nn.grad.conv2d_weight(xp[b : b + 1], w.shape, dy[b : b + 1]) for b in range(n)])
    Q = torch.einsum("no,noikl,oikl->ni", sigma, G, w)
    P = sigma * torch.einsum("noikl,oikl,ni->no", G, w, s.detach())
    assert torch.allclose(Q, ds, atol=1e-10)
    assert torch.allclose(P, (dy * u.detach()).sum((2, 3)), atol=1e-10)
