"""Micro-benchmark of the conv kernels on the layer shapes of BASELINE configs (CUDA events,
L2 flushed between timed launches).  Usage: python tools/bench_conv.py [--cfg 128|256]"""
import argparse
import json
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from one_to_many_gan_b200 import kernels as K  # noqa: E402

SHAPES = {
    "128": [  # name, cin, cout, k, pad, halo, H, W, B, per_sample
        ("res3x3_128", 128, 128, 3, 1, 1, 64, 64, 32, False),
        ("mod3x3_128", 128, 128, 3, 1, 1, 64, 64, 32, True),
        ("enc3x3_64_128", 64, 128, 3, 1, 0, 128, 128, 32, False),
        ("up3x3_128_64", 128, 64, 3, 1, 0, 128, 128, 32, True),
        ("d4x4_64_128", 64, 128, 4, 1, 0, 63, 63, 32, False),
        ("d4x4_128_256", 128, 256, 4, 1, 0, 31, 31, 32, False),
        ("d4x4_256_512", 256, 512, 4, 1, 0, 15, 15, 32, False),
    ],
    "256": [
        ("res3x3_256", 256, 256, 3, 1, 1, 64, 64, 32, False),
        ("mod3x3_256", 256, 256, 3, 1, 1, 64, 64, 32, True),
        ("enc3x3_128_256", 128, 256, 3, 1, 0, 128, 128, 32, False),
        ("up3x3_256_128", 256, 128, 3, 1, 0, 128, 128, 32, True),
        ("up3x3_128_64", 128, 64, 3, 1, 0, 256, 256, 32, True),
    ],
}


def timeit(fn, flush, iters=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        flush.zero_()
        s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        s.record()
        fn()
        e.record()
        torch.cuda.synchronize()
        ts.append(s.elapsed_time(e))
    ts.sort()
    return ts[len(ts) // 2]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--cfg", default="128")
    args = ap.parse_args()
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for name, cin, cout, k, pad, halo, H, W, B, ps in SHAPES[args.cfg]:
        x = K.alloc(B, cin, H, W, torch.bfloat16, dev, halo, zero=True)
        K.padded_view(x, halo).normal_()
        w = torch.randn(cout, cin, k, k, device=dev)
        alpha = 1 / math.sqrt(cin * k * k)
        s = torch.rand(B, cin, device=dev) + 0.5
        rs = torch.rand(B, cout, device=dev) + 0.5
        wp = K.weight_pack(w, alpha, torch.bfloat16, cs=s if ps else None, nb=B if ps else 1)
        ho, wo = H + 2 * pad - k + 1, W + 2 * pad - k + 1
        y = K.alloc(B, cout, ho, wo, torch.bfloat16, dev)
        flops = 2.0 * B * ho * wo * cout * cin * k * k
        t_f = timeit(lambda: K.conv_fwd(x, wp, cout, k, k, pad, x_halo=halo, per_sample=ps,
                                        row_scale=rs if ps else None, act=K.ACT_RELU, out=y), flush)
        dy = torch.randn(B, ho, wo, cout, device=dev).bfloat16().permute(0, 3, 1, 2)
        wpt = K.weight_pack(w, alpha, torch.bfloat16, rs=rs if ps else None, nb=B if ps else 1,
                            transpose=True)
        pd = k - 1 - (pad - halo)
        gx = K.alloc(B, cin, H + 2 * halo, W + 2 * halo, torch.bfloat16, dev)
        t_d = timeit(lambda: K.conv_fwd(dy, wpt, cin, k, k, pd, per_sample=ps, out=gx), flush)
        dw = torch.zeros_like(w)
        t_w = timeit(lambda: K.conv_wgrad(x, dy, dw, k, k, pad, x_halo=halo, alpha=alpha,
                                          rs=rs if ps else None, cs=s if ps else None), flush)
        row = dict(layer=name, gflop=round(flops / 1e9, 1),
                   fwd_ms=round(t_f, 4), fwd_tflops=round(flops / t_f / 1e9, 1),
                   dgrad_ms=round(t_d, 4), dgrad_tflops=round(flops / t_d / 1e9, 1),
                   wgrad_ms=round(t_w, 4), wgrad_tflops=round(flops / t_w / 1e9, 1))
        rows.append(row)
        print(json.dumps(row), flush=True)


if __name__ == "__main__":
    main()
