"""Time the dominant launches of the bench workload alone (CUDA events, L2 flushed): the 3x3
128->128 conv at 64x64, n = 96, in the forms the iteration launches.  With `ncu`:
    ncu --set full --clock-control none --import-source on -k regex:rr2t -c 1 python tools/bench_one.py shared"""
import math
import statistics
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from one_to_many_gan_b200 import kernels as K  # noqa: E402

only = sys.argv[1] if len(sys.argv) > 1 else None
dev = "cuda"
n, c, hw = 96, 128, 64
x = K.alloc(n, c, hw, hw, torch.bfloat16, dev, 1, zero=True)
K.padded_view(x, 1).normal_()
res = K.alloc(n, c, hw, hw, torch.bfloat16, dev, 1, zero=True)
res.normal_()
w = torch.randn(c, c, 3, 3, device=dev)
s = torch.rand(n, c, device=dev) + 0.5
sig = torch.rand(n, c, device=dev) + 0.5
alpha = 1 / math.sqrt(c * 9)
wp = K.weight_pack(w, alpha, torch.bfloat16, cs=s, nb=n)
wp1 = K.weight_pack(w, alpha, torch.bfloat16)
y = K.alloc(n, c, hw, hw, torch.bfloat16, dev, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flops = 2.0 * n * hw * hw * c * c * 9
forms = {
    "shared": lambda: K.conv_fwd(x, wp1, c, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig, residual=res, out=y),
    "per_sample": lambda: K.conv_fwd(x, wp, c, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig, act=K.ACT_RELU,
                                     post_scale=s, per_sample=True, out=y),
    "plain": lambda: K.conv_fwd(x, wp1, c, 3, 3, 1, x_halo=1, out=y),
}
if only:
    forms[only]()
    torch.cuda.synchronize()
    print("ran", only)
    sys.exit(0)
for name, fn in forms.items():
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(8):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    t = statistics.median(ts)
    print(f"{name}: {t * 1e3:.1f} us {flops / t / 1e9:.0f} TFLOP/s")
