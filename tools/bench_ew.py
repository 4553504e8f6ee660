"""GB/s of the HBM-bound passes on the big tensors of the 128x128/b32 iteration."""
import statistics, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from one_to_many_gan_b200 import kernels as K
dev = "cuda"; flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(fn):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(6):
        flush.zero_(); a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
def mk(n, c, h, w, halo=0):
    t = K.alloc(n, c, h, w, torch.bfloat16, dev, halo, zero=True); K.padded_view(t, halo).normal_(); return t
for (n, c, h, w) in [(64, 128, 128, 128), (64, 128, 64, 64), (160, 128, 64, 64)]:
    x = mk(n, c, h, w); g = mk(n, c, h, w); gp = mk(n, c, h, w, 1)
    st = K.instnorm_stats(x); mb = n * c * h * w * 2 / 1e6
    y = torch.empty_like(x)
    t = run(lambda: y.copy_(x)); print(f"[{n},{c},{h},{w}] torch copy   {t*1e3:7.1f} us {2*mb/t/1e3:6.2f} TB/s (1R+1W, calibration)")
    t = run(lambda: torch.add(x, g, out=y)); print(f"[{n},{c},{h},{w}] torch add    {t*1e3:7.1f} us {3*mb/t/1e3:6.2f} TB/s (2R+1W, calibration)")
    t = run(lambda: K.instnorm_stats(x)); print(f"[{n},{c},{h},{w}] stats        {t*1e3:7.1f} us {mb/t/1e3:6.2f} TB/s (1R)")
    t = run(lambda: K.norm_act(x, st, K.ACT_RELU, y_halo=1)); print(f"[{n},{c},{h},{w}] norm_act     {t*1e3:7.1f} us {2*mb/t/1e3:6.2f} TB/s (1R+1W)")
    t = run(lambda: K.norm_act_bwd(g, x, st, K.ACT_RELU)); print(f"[{n},{c},{h},{w}] norm_act_bwd {t*1e3:7.1f} us {5*mb/t/1e3:6.2f} TB/s (4R+1W, two kernels)")
    t = run(lambda: K.norm_act_bwd(gp, None, None, K.ACT_NONE, g_halo=1, g2=g)); print(f"[{n},{c},{h},{w}] fold+add     {t*1e3:7.1f} us {3*mb/t/1e3:6.2f} TB/s (2R+1W)")
    s = torch.rand(n, c, device=dev)
    t = run(lambda: K.mod_in(gp, x, s, g_halo=1, gadd=g, relu_mask=True)); print(f"[{n},{c},{h},{w}] mod_in       {t*1e3:7.1f} us {4*mb/t/1e3:6.2f} TB/s (3R+1W)")
    t = run(lambda: K.down(x, st, K.ACT_RELU, 1)); print(f"[{n},{c},{h},{w}] down         {t*1e3:7.1f} us {1.25*mb/t/1e3:6.2f} TB/s (1R+.25W)")
    t = run(lambda: K.up(x)); print(f"[{n},{c},{h},{w}] up           {t*1e3:7.1f} us {5*mb/t/1e3:6.2f} TB/s (1R+4W)")
