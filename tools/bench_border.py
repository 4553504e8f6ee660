"""Time otm_conv_reflect_border alone at the bench's launch shapes (CUDA events, L2 flushed)."""
import math
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from one_to_many_gan_b200 import _lib as L  # noqa: E402
from one_to_many_gan_b200 import kernels as K  # noqa: E402

dev = torch.device("cuda", 0)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for n, k, c, h, w, per_sample, gate in [(96, 128, 128, 64, 64, False, False), (96, 128, 128, 64, 64, True, True),
                                        (64, 128, 128, 64, 64, False, False),
                                        (96, 128, 128, 128, 128, False, False), (96, 128, 128, 128, 128, True, True)]:
    g = K.alloc(n, k, h, w, torch.bfloat16, dev).normal_()
    wt = torch.randn(k, c, 3, 3, device=dev) / math.sqrt(9 * k)
    rs = torch.rand(n, k, device=dev) + 0.5
    wp = K.weight_pack(wt, 1.0, torch.bfloat16, rs=rs if per_sample else None, nb=n if per_sample else 1,
                       transpose=True)
    y = K.alloc(n, c, h, w, torch.bfloat16, dev).zero_()
    ht = K.alloc(n, c, h, w, torch.bfloat16, dev, 1).normal_() if gate else None
    dot = torch.zeros(n, c, device=dev) if gate else None
    b = L.ConvReflectBorderArgs()
    b.dy, b.wpack, b.y = L.tdesc(g), L.ptr(wp), L.tdesc(y)
    b.w_batch_stride = 9 * k * c if per_sample else 0
    b.gate, b.dot_sums = L.tdesc(ht), L.ptr(dot)
    ts = []
    for it in range(6):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.check(L.lib.otm_conv_reflect_border(K._byref(b), L.stream_ptr()), "border")
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    print(f"n={n} {k}->{c} @{h}x{w} per_sample={per_sample} gate={gate}: {min(ts[1:]):.1f} us (median {sorted(ts[1:])[2]:.1f})")
