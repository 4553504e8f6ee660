"""Run warm-up iterations of the bench workload, then ONE iteration inside a
cudaProfilerStart/Stop range (use with `ncu --profile-from-start off`)."""
import argparse
import sys
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from one_to_many_gan_b200.synthetic import SyntheticImages  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("--warm", type=int, default=2)
ap.add_argument("--batch", type=int, default=bench.BATCH)
args = ap.parse_args()
bench.BATCH = args.batch
bench.CONFIG["training"]["batch_size"] = args.batch
dev = torch.device("cuda", 0)
step = bench.build_trainer(dev, 0, use_graph=False)
prints = SyntheticImages(args.batch, 1, bench.IMAGE, dev, seed=42, stream_id=0)
marks = SyntheticImages(args.batch, 1, bench.IMAGE, dev, seed=42, stream_id=1)
for _ in range(args.warm):
    step(prints, marks)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
out = step(prints, marks)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled step losses", out)
