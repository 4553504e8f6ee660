"""Run a few iterations of the engine at another BASELINE config (size, batch) and report
time per iteration and peak memory.
Usage: python tools/run_config.py H W BATCH [iters [styles_per_input [r1_gamma]]]"""
import sys
import time
from pathlib import Path

import torch

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import bench  # noqa: E402
from one_to_many_gan_b200.synthetic import SyntheticImages  # noqa: E402

H, W, B = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 4
K_STYLES = int(sys.argv[5]) if len(sys.argv) > 5 else 1   # BASELINE config 4: styles per input
R1_GAMMA = float(sys.argv[6]) if len(sys.argv) > 6 else 0.0  # BASELINE config 5: R1 penalty
bench.IMAGE = (H, W)
bench.BATCH = B
bench.CONFIG["training"]["batch_size"] = B
bench.CONFIG["data"]["image_size"] = [H, W]
bench.CONFIG["training"]["styles_per_input"] = K_STYLES
bench.CONFIG["optimisation"]["r1_gamma"] = R1_GAMMA
dev = torch.device("cuda", 0)
step = bench.build_trainer(dev, 0, use_graph=True)
prints = SyntheticImages(B, 1, (H, W), dev, seed=42, stream_id=0)
marks = SyntheticImages(B, 1, (H, W), dev, seed=42, stream_id=1)
out = None
for i in range(iters):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    out = step(prints, marks)
    torch.cuda.synchronize()
    print(f"iter {i}: {1e3 * (time.perf_counter() - t0):.1f} ms", flush=True)
print({k: round(v, 4) for k, v in out.items()})
print(f"{H}x{W} batch {B}: peak memory {torch.cuda.max_memory_allocated() / 2**30:.1f} GiB, "
      f"{B / (time.perf_counter() - t0):.1f} img/s last iteration")
