import sys, math, torch
sys.path.insert(0, "/root/repo")
from one_to_many_gan_b200 import kernels as K
dev="cuda"
torch.manual_seed(0)
for (n,cin,cout,h,w,k,pad,halo) in [(32,128,128,64,64,3,1,1),(8,64,128,128,128,3,1,0),(16,128,256,31,31,4,1,0),(32,64,128,63,63,4,1,0)]:
    x = K.alloc(n,cin,h,w,torch.bfloat16,dev,halo,zero=True); K.padded_view(x,halo).normal_()
    wt = torch.randn(cout,cin,k,k,device=dev)
    bias = torch.randn(cout,device=dev)
    wp = K.weight_pack(wt, 1/math.sqrt(cin*k*k), torch.bfloat16)
    y, st = K.conv_fwd(x, wp, cout, k, k, pad, x_halo=halo, bias=bias, want_stats=True)
    ref = K.instnorm_stats(y)
    y2 = K.conv_fwd(x, wp, cout, k, k, pad, x_halo=halo, bias=bias)
    torch.cuda.synchronize()
    a=K.ConvFwdArgs if hasattr(K,'ConvFwdArgs') else None
    dm = (st[...,0]-ref[...,0]).abs().max().item(); dr = ((st[...,1]-ref[...,1]).abs()/ref[...,1]).max().item()
    print((n,cin,cout,h,w,k), "mean diff", dm, "rstd rel diff", dr, "y equal", torch.equal(y,y2))
    assert dm < 2e-3 and dr < 2e-3
print("ok")
