"""2-GPU probe of the data-parallel engine modes (run under torchrun with a `timeout`):
prints per-iteration losses, dumps all thread stacks if an iteration takes longer than 60 s."""
import faulthandler
import os
import random
import sys
import time
from pathlib import Path

import torch
import torch.distributed as dist

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
faulthandler.enable()

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from one_to_many_gan_b200 import builder  # noqa: E402
from one_to_many_gan_b200.engine import TrainIteration  # noqa: E402
from one_to_many_gan_b200.optim import FlatAdam  # noqa: E402

size, batch = (64, 64), 2
cfg = {"training": {"batch_size": batch, "image_buffer_size": 10},
       "optimisation": {"style_cycle_loss_lambda": 5.0, "identity_loss_lambda": 5.0,
                        "reconstruction_loss_lambda": 5.0, "kl_loss_lambda": 0.01,
                        "path_loss_lambda": 0.1, "path_loss_jacobian_granularity": [0.1, 0.2]},
       "architecture": {"add_latent_noise": False},
       "data": {"image_size": list(size), "image_channels": 1}}
torch.manual_seed(42)
dt = torch.bfloat16 if os.environ.get("PROBE_BF16", "1") == "1" else torch.float32
D = builder.Discriminator(1, act_dtype=dt).to(dev)
G = builder.Generator(1, 6, size, 32, 5, act_dtype=dt).to(dev)
M = builder.MappingNetwork(6, 2, 0.9).to(dev)
S = builder.StyleExtractor(1, 6, act_dtype=dt).to(dev)
opts = [FlatAdam(D.parameters(), 2e-3, (0.5, 0.99)), FlatAdam(G.parameters(), 2e-3, (0.5, 0.99)),
        FlatAdam(M.parameters(), 2e-5, (0.5, 0.99)), FlatAdam(S.parameters(), 2e-3, (0.5, 0.99))]
eng = TrainIteration(cfg, dev, D, G, M, S, *opts, use_graph=os.environ.get("PROBE_GRAPH", "1") == "1", warmup=1)
print(f"[rank {rank}] mode={eng.mode} graph={eng.use_graph}", flush=True)
torch.manual_seed(100 + rank)
random.seed(100 + rank)
for it in range(4):
    faulthandler.dump_traceback_later(60, exit=True)
    g = torch.Generator().manual_seed(1000 * rank + it)
    eng.load_inputs(*[(torch.rand(batch, 1, *size, generator=g) * 2 - 1).to(dev) for _ in range(4)])
    t0 = time.time()
    out = eng.run(h=torch.tensor([0.12, 0.18]))
    faulthandler.cancel_dump_traceback_later()
    print(f"[rank {rank}] it {it} {time.time() - t0:.2f}s total_gen={out['total_gen']:.4f}", flush=True)
flat = torch.cat([o.param_arena for o in opts])
other = flat.clone()
dist.broadcast(other, src=0)
print(f"[rank {rank}] replicas equal: {torch.equal(other, flat)} checksum {flat.double().sum().item():.6f}", flush=True)
from one_to_many_gan_b200.optim import shutdown_process_group  # noqa: E402

t0 = time.time()
eng.close()
shutdown_process_group()
print(f"[rank {rank}] clean shutdown in {time.time() - t0:.1f}s", flush=True)
