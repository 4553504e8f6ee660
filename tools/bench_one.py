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
wpt = K.weight_pack(w, alpha, torch.bfloat16, rs=sig, nb=n, transpose=True)
wpt1 = K.weight_pack(w, alpha, torch.bfloat16, transpose=True)
y = K.alloc(n, c, hw, hw, torch.bfloat16, dev, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flops = 2.0 * n * hw * hw * c * c * 9
forms = {
    "shared": lambda: K.conv_fwd(x, wp1, c, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig, residual=res, out=y),
    "per_sample": lambda: K.conv_fwd(x, wp, c, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig, act=K.ACT_RELU,
                                     post_scale=s, per_sample=True, out=y),
    "plain": lambda: K.conv_fwd(x, wp1, c, 3, 3, 1, x_halo=1, out=y),
    # backward of a ModulatedResnetBlock (reflect-pad dgrad = same-size launch + halo-ring launch):
    # conv2's dgrad with the fused input-side pass (per-sample packs, gate, dot) ...
    "dgrad_gate": lambda: K.conv_dgrad_reflect(x, wpt, c, per_sample=True, gate=res, row_scale=s,
                                               post_scale=sig, want_dot=True),
    # ... conv1's dgrad (shared pack, s1 as row scale, skip gradient as residual)
    "dgrad_residual": lambda: K.conv_dgrad_reflect(x, wpt1, c, residual=res, row_scale=s),
    # the round-2 form of the same gradient: padded 66x66 dgrad (folded by the next pass)
    "dgrad_padded": lambda: K.conv_fwd(x, wpt1, c, 3, 3, 2),
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
