"""Time the dominant launch (modulated 3x3 128->128 @64x64, n=96) under the current env knobs."""
import math, statistics, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from one_to_many_gan_b200 import kernels as K
dev = "cuda"
n, c, hw = 96, int(sys.argv[1]) if len(sys.argv) > 1 else 128, 64
x = K.alloc(n, c, hw, hw, torch.bfloat16, dev, 1, zero=True); K.padded_view(x, 1).normal_()
w = torch.randn(c, c, 3, 3, device=dev); s = torch.rand(n, c, device=dev) + 0.5; sig = torch.rand(n, c, device=dev) + 0.5
wp = K.weight_pack(w, 1 / math.sqrt(c * 9), torch.bfloat16, cs=s, nb=n)
wp1 = K.weight_pack(w, 1 / math.sqrt(c * 9), torch.bfloat16)
y = K.alloc(n, c, hw, hw, torch.bfloat16, dev, 1)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
flops = 2.0 * n * hw * hw * c * c * 9
def run(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(8):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
t1 = run(lambda: K.conv_fwd(x, wp, c, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig, act=K.ACT_RELU, per_sample=True, out=y))
t2 = run(lambda: K.conv_fwd(x, wp1, c, 3, 3, 1, x_halo=1, y_halo=0, out=y))
print(f"modulated+halo: {t1*1e3:.1f} us {flops/t1/1e9:.0f} TF/s | shared plain: {t2*1e3:.1f} us {flops/t2/1e9:.0f} TF/s")
for name, kw, wpk in [
    ("per-sample only", dict(per_sample=True), wp),
    ("per-sample+scale", dict(per_sample=True, row_scale=sig), wp),
    ("per-sample+scale+relu", dict(per_sample=True, row_scale=sig, act=K.ACT_RELU), wp),
    ("shared+halo", dict(y_halo=1), wp1),
]:
    t = run(lambda: K.conv_fwd(x, wpk, c, 3, 3, 1, x_halo=1, out=y, **kw))
    print(f"  {name}: {t*1e3:.1f} us {flops/t/1e9:.0f} TF/s")
