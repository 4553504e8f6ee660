import math, statistics, sys
from pathlib import Path
import torch
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
from one_to_many_gan_b200 import kernels as K
dev="cuda"; flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def run(fn):
    fn(); torch.cuda.synchronize(); ts=[]
    for _ in range(8):
        flush.zero_(); a,b=torch.cuda.Event(enable_timing=True),torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    return statistics.median(ts)
for (n,cin,cout,hw,halo) in [(96,128,128,64,1),(64,128,128,64,1),(64,64,128,128,0),(160,128,64,128,0)]:
    x = K.alloc(n,cin,hw,hw,torch.bfloat16,dev,halo,zero=True); K.padded_view(x,halo).normal_()
    dy = torch.randn(n,hw,hw,cout,device=dev).bfloat16().permute(0,3,1,2)
    dw = torch.zeros(cout,cin,3,3,device=dev); rs=torch.rand(n,cout,device=dev); cs=torch.rand(n,cin,device=dev)
    t = run(lambda: K.conv_wgrad(x,dy,dw,3,3,1,x_halo=halo,alpha=0.1,rs=rs,cs=cs))
    fl = 2.0*n*hw*hw*cin*cout*9
    print(f"wgrad n={n} {cin}->{cout} @{hw}: {t*1e3:.1f} us {fl/t/1e9:.0f} TF/s")
