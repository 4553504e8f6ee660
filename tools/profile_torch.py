"""Per-kernel GPU time of one eager training iteration via torch.profiler (CUPTI): fast,
concurrent-safe, warm caches.  Usage: python tools/profile_torch.py [--batch B] [--top N]"""
import argparse
import sys
from collections import defaultdict
from pathlib import Path

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from one_to_many_gan_b200.synthetic import SyntheticImages  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=bench.BATCH)
ap.add_argument("--top", type=int, default=45)
ap.add_argument("--iters", type=int, default=2)
ap.add_argument("--size", type=int, default=0, help="square image size (default: the bench config)")
args = ap.parse_args()
bench.BATCH = args.batch
if args.size:
    bench.IMAGE = (args.size, args.size)
    bench.CONFIG["data"]["image_size"] = [args.size, args.size]
bench.CONFIG["training"]["batch_size"] = args.batch
dev = torch.device("cuda", 0)
step = bench.build_trainer(dev, 0, use_graph=False)
prints = SyntheticImages(args.batch, 1, bench.IMAGE, dev, seed=42, stream_id=0)
marks = SyntheticImages(args.batch, 1, bench.IMAGE, dev, seed=42, stream_id=1)
for _ in range(3):
    step(prints, marks)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(args.iters):
        step(prints, marks)
    torch.cuda.synchronize()
tot = defaultdict(lambda: [0, 0.0])
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.split("(")[0]
        tot[name][0] += 1
        tot[name][1] += ev.device_time_total if hasattr(ev, "device_time_total") else ev.cuda_time_total
total = sum(v[1] for v in tot.values())
print(f"total GPU kernel time per iteration {total / 1e3 / args.iters:.3f} ms, "
      f"{sum(v[0] for v in tot.values()) // args.iters} launches")
for name, (n, us) in sorted(tot.items(), key=lambda kv: -kv[1][1])[: args.top]:
    print(f"{us / 1e3 / args.iters:9.3f} ms {100 * us / total:5.1f}% {n // args.iters:5d}x  {name[:105]}")
