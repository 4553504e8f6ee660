"""One InstanceNorm + ReLU backward on [64,128,64,64] bf16 (the bench's `roofline_hbm` launch) for
ncu: the row_stream_kernel launches are, in order, StatsRowOp, NabRowOp<1> (reductions) and
NabRowOp<0> (apply).  Usage: ncu --set full -k regex:row_stream --launch-skip 2 -c 1 python tools/prof_stream.py"""
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from one_to_many_gan_b200 import kernels as K  # noqa: E402

x = K.alloc(64, 128, 64, 64, torch.bfloat16, "cuda", 0, zero=True)
x.normal_()
g = torch.randn_like(x)
st = K.instnorm_stats(x)
torch.cuda.synchronize()
K.norm_act_bwd(g, x, st, K.ACT_RELU)
torch.cuda.synchronize()
print("done")
