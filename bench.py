"""Headline benchmark: G+D training iterations per second (images/sec) of the one-to-many GAN.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference]
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Workload (BASELINE.json configs[1]): default architecture at 128x128, batch 32 per GPU, bf16
activations / fp32 accumulation, synthetic U(-1,1) images, random-init weights (seed 42).
One step = `discriminator_step` + `generator_step` (reference src/core/training.py:71,136).
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is produced."""

from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

IMAGE = (128, 128)
BATCH = 32
GFLOP_PER_IMAGE = 301.7  # dense-conv algorithmic FLOPs per image per iteration, SURVEY.md §8(d)
CONFIG = {
    "training": {"batch_size": BATCH, "style_mixing_prob": 0.9, "random_seed": 42,
                 "image_buffer_size": 100},
    "optimisation": {
        "style_cycle_loss_lambda": 5.0, "identity_loss_lambda": 5.0,
        "reconstruction_loss_lambda": 5.0, "kl_loss_lambda": 0.01, "path_loss_lambda": 0.1,
        "path_loss_jacobian_granularity": [0.1, 0.2], "learning_rate": 2e-3,
        "mapping_network_learning_rate": 2e-5, "adam_betas": [0.5, 0.99],
    },
    "architecture": {"w_dim": 6, "add_latent_noise": False, "min_latent_resolution": 64,
                     "n_resnet_blocks": 7, "mapping_network_layers": 2},
    "data": {"image_size": list(IMAGE), "image_channels": 1},
}
WORKLOAD = "G+D train iteration, default arch, 128x128, batch 32/GPU, bf16 act / fp32 accum"
# dense-conv algorithmic GFLOP per image per iteration for the extra configs
GFLOP_256 = 1441.2
# dram bytes (read + write) of the roofline launch: profiles/r2_ncu_conv_tc_fwd_rr2t_residual.md
NCU_TRAFFIC_RESIDUAL = 282.1e6


def base_config(world):
    """The `config` object both arms print (the driver compares them)."""
    return {"workload": WORKLOAD, "global_batch": BATCH * world, "parallelism": f"dp{world}",
            "l2": "activations per step (>1 GB) exceed the 126 MB L2"}


def make_config(image, batch):
    import copy

    cfg = copy.deepcopy(CONFIG)
    cfg["training"]["batch_size"] = batch
    cfg["data"]["image_size"] = list(image)
    return cfg


def peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        d = json.loads(p.read_text())
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows: list[list[str]] = []
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-i", str(index), "-lms", "200"],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for nm, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": max(mx) if mx else None, "reasons": sorted(reasons),
                "samples": len(sm)}


class HostBatches:
    """Pinned host batches: the e2e leg copies its inputs host->device every step."""

    def __init__(self, shape, seed, n_distinct=4):
        g = torch.Generator().manual_seed(seed)
        self.pool = [(torch.rand(*shape, generator=g) * 2 - 1).pin_memory() for _ in range(n_distinct)]
        self.i = 0

    def __iter__(self):
        return self

    def __next__(self):
        t = self.pool[self.i % len(self.pool)]
        self.i += 1
        return t


def build_trainer(device, rank, use_graph=True, config=None, return_engine=False):
    """The four networks + optimisers + the CUDA-graph iteration engine (public API:
    one_to_many_gan_b200.engine.TrainIteration)."""
    from one_to_many_gan_b200 import builder
    from one_to_many_gan_b200.engine import TrainIteration
    from one_to_many_gan_b200.optim import FlatAdam

    config = config or CONFIG
    image = tuple(config["data"]["image_size"])
    torch.manual_seed(42)
    arch = config["architecture"]
    dt = torch.bfloat16
    D = builder.Discriminator(1, act_dtype=dt).to(device)
    G = builder.Generator(1, arch["w_dim"], image, arch["min_latent_resolution"],
                          arch["n_resnet_blocks"], act_dtype=dt).to(device)
    M = builder.MappingNetwork(arch["w_dim"], arch["mapping_network_layers"], 0.9).to(device)
    S = builder.StyleExtractor(1, arch["w_dim"], act_dtype=dt).to(device)
    o = config["optimisation"]
    betas = tuple(o["adam_betas"])
    oD = FlatAdam(D.parameters(), o["learning_rate"], betas)
    oG = FlatAdam(G.parameters(), o["learning_rate"], betas)
    oM = FlatAdam(M.parameters(), o["mapping_network_learning_rate"], betas)
    oS = FlatAdam(S.parameters(), o["learning_rate"], betas)
    torch.manual_seed(1234 + rank)  # per-rank style / theta draws
    import random

    random.seed(1234 + rank)
    eng = TrainIteration(config, device, D, G, M, S, oD, oG, oM, oS, use_graph=use_graph, warmup=2)

    def step(prints, marks):
        # one iteration consumes two shoeprint and two shoemark batches (D step, then G step)
        eng.load_inputs(next(prints), next(marks), next(prints), next(marks))
        return eng.run(sync_losses=True)  # includes the device->host read of the 10 scalars

    return (step, eng) if return_engine else step


def time_steps(step, prints, marks, steps, warmup, dist_on, device):
    import torch.distributed as dist

    for _ in range(warmup):
        step(prints, marks)
    if dist_on:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    last = None
    for _ in range(steps):
        last = step(prints, marks)
    e1.record()
    torch.cuda.synchronize()
    if dist_on:
        dist.barrier()
    ms = e0.elapsed_time(e1)
    if dist_on:
        t = torch.tensor([ms], device=device)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.item()
    return ms, last


def dominant_kernel_roofline(device):
    """The 3x3 128->128 conv at 64x64 is ~69 % of the dense MACs of this workload
    and `conv_tc_fwd_rr2t_kernel<3,2,4>` runs all of them (forward and dgrad, 62 launches per
    iteration).  Its heaviest form is timed alone -- CUDA events on the launching stream, L2
    flushed between launches -- against the measured burst peak: conv2 of a ModulatedResnetBlock
    (shared weights, demodulation row scale, residual add, reflect halo) at n = 96 = the 3B decode
    batch.  The other two forms it is launched in are timed the same way and reported next to it
    with their launch counts: conv1 of the block (per-sample weight packs, ReLU, the next conv's
    style scale as post-activation scale, halo) and the bare shared-weight form (dgrad)."""
    import math

    from one_to_many_gan_b200 import kernels as K

    n, c, hw = BATCH * 3, 128, 64  # the 3B decode batch the G step actually launches
    x = K.alloc(n, c, hw, hw, torch.bfloat16, device, 1, zero=True)
    K.padded_view(x, 1).normal_()
    res = K.alloc(n, c, hw, hw, torch.bfloat16, device, 1, zero=True)
    res.normal_()
    w = torch.randn(c, c, 3, 3, device=device)
    s = torch.rand(n, c, device=device) + 0.5
    sig = torch.rand(n, c, device=device) + 0.5
    alpha = 1 / math.sqrt(c * 9)
    wp_shared = K.weight_pack(w, alpha, torch.bfloat16)
    wp_sample = K.weight_pack(w, alpha, torch.bfloat16, cs=s, nb=n)
    y = K.alloc(n, c, hw, hw, torch.bfloat16, device, 1)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    flops = 2.0 * n * hw * hw * c * c * 9

    def residual():  # conv2: y = x_res + sigma_inv * conv(h~, cW)
        K.conv_fwd(x, wp_shared, c, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig, residual=res, out=y)

    def per_sample():  # conv1: h~ = s2 * relu(sigma_inv * conv(x, cW * s1))
        K.conv_fwd(x, wp_sample, c, 3, 3, 1, x_halo=1, y_halo=1, row_scale=sig, act=K.ACT_RELU,
                   post_scale=s, per_sample=True, out=y)

    def plain():  # shared pack, no epilogue work (encoder ResnetBlock convs, plain dgrads)
        K.conv_fwd(x, wp_shared, c, 3, 3, 1, x_halo=1, out=y)

    wpt_sample = K.weight_pack(w, alpha, torch.bfloat16, rs=sig, nb=n, transpose=True)
    wpt_shared = K.weight_pack(w, alpha, torch.bfloat16, transpose=True)

    def dgrad_gate():  # conv2's dgrad through ReflectionPad2d(1): gate + dot epilogue, + ring launch
        K.conv_dgrad_reflect(x, wpt_sample, c, per_sample=True, gate=res, row_scale=s, post_scale=sig,
                             want_dot=True)

    def dgrad_residual():  # conv1's dgrad: s1 row scale, skip gradient as residual, + ring launch
        K.conv_dgrad_reflect(x, wpt_shared, c, residual=res, row_scale=s)

    def timed(fn):
        fn()
        torch.cuda.synchronize()
        ts = []
        for _ in range(10):
            flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.mean(ts)

    pk, how = peaks()
    peak = pk["bf16_tflops"]
    ms = timed(residual)
    ach = flops / ms / 1e9
    forms = {}
    for name, fn, cnt in (("per_sample_weights_relu_poststyle_halo", per_sample, 12),
                          ("shared_weights_bare", plain, 18),
                          ("reflect_dgrad_gate_dot_per_sample_plus_ring_launch", dgrad_gate, 8),
                          ("reflect_dgrad_rowscale_residual_plus_ring_launch", dgrad_residual, 12)):
        t = timed(fn)
        forms[name] = {"achieved": round(flops / t / 1e9, 1), "frac": round(flops / t / 1e9 / peak, 4),
                       "launch_ms": round(t, 4), "launches_per_iteration": cnt}
    # dram__bytes_read.sum + dram__bytes_write.sum of this launch from `ncu --set full`
    # (profiles/r2_ncu_conv_tc_fwd_rr2t_residual.md); algorithmic bytes = x (107 MB) + residual
    # (101 MB) + y (107 MB) + weights (0.3 MB) = 315 MB.
    return {"bound": "tensor",
            "kernel": "conv_tc_fwd_rr2t_kernel<3,2,4>: shared-weight 3x3 128->128 @64x64, n=96 "
                      "(+demodulation row scale, residual add, reflect halo)",
            "achieved": round(ach, 1), "peak": peak, "peak_source": how + " burst",
            "unit": "TFLOP/s", "frac": round(ach / peak, 4), "traffic": NCU_TRAFFIC_RESIDUAL,
            "launch_ms": round(ms, 4), "flops_per_launch": flops, "launches_per_iteration": 12,
            "other_forms": forms}


def hbm_kernel_roofline(device):
    """The heaviest HBM-bound stage of the iteration -- the InstanceNorm + ReLU backward of a
    res-block conv, [64,128,64,64] bf16 (otm_norm_act_bwd: per-(n,c) reductions, then the apply)
    -- timed alone (CUDA events, L2 flushed) against the measured copy bandwidth.  Algorithmic
    bytes: reduction 2R + apply 2R + 1W of 67.1 MB each."""
    import statistics

    from one_to_many_gan_b200 import kernels as K

    n, c, hw = 64, 128, 64
    x = K.alloc(n, c, hw, hw, torch.bfloat16, device, 0, zero=True)
    x.normal_()
    g = torch.randn_like(x)
    st = K.instnorm_stats(x)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=device)
    nbytes = 5.0 * x.numel() * 2

    def launch():
        K.norm_act_bwd(g, x, st, K.ACT_RELU)

    launch()
    torch.cuda.synchronize()
    ts = []
    for _ in range(10):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        launch()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ms = statistics.mean(ts)
    pk, how = peaks()
    ach = nbytes / ms / 1e6
    return {"bound": "hbm", "kernel": "otm_norm_act_bwd (NormActBwdReduceF + NormActBwdApplyF), "
                                       "[64,128,64,64] bf16, InstanceNorm + ReLU backward",
            "achieved": round(ach, 1), "peak": pk["hbm_gbs"], "peak_source": how, "unit": "GB/s",
            "frac": round(ach / pk["hbm_gbs"], 4), "traffic": None, "launch_ms": round(ms, 4),
            "bytes_per_launch": nbytes}


def _oracle_trainer(batch, device="cpu"):
    from oracle import reference_port as rp

    arch = rp.Arch(image_size=IMAGE)
    tr = rp.Trainer(arch, rp.Hyper(batch_size=batch), rp.init_all(arch, 42), device=device)

    def one(i):
        b = [rp.synthetic_batch(batch, arch, 10 * k + i).to(device) for k in range(1, 5)]
        tr.discriminator_step(b[0], b[1])
        tr.generator_step(b[2], b[3])

    return one


def cpu_baseline(batches=(2, 8, BATCH)):
    """The oracle port (oracle/reference_port.py) of the same iteration on the host cores:
    one warm-up iteration (batch 2: thread pools, MKL-DNN primitive caches), then ONE timed D+G
    iteration per batch size -- the last one is the workload's own batch; the smaller ones show
    how the CPU's per-image cost depends on the batch.  `value` is the workload-batch number."""
    torch.set_num_threads(os.cpu_count() or 1)
    _oracle_trainer(2)(0)  # warm-up
    by_batch = {}
    for b in batches:
        one = _oracle_trainer(b)
        t0 = time.perf_counter()
        one(100)
        by_batch[str(b)] = round(b / (time.perf_counter() - t0), 4)
    return {"value": by_batch[str(batches[-1])], "unit": "images/sec", "cores": torch.get_num_threads(),
            "kind": "port", "by_batch": by_batch,
            "sample": f"1 warm-up (batch 2) + 1 timed D+G iteration each at batch {list(batches)}, "
                      f"128x128, fp32, torch CPU; value = batch {batches[-1]} (the workload's)"}


def gpu_baseline(device, iters=10, warm=2):
    """The 'existing Blackwell kernel' bar: the SAME iteration as eager
    PyTorch on this B200 -- the oracle port's functional torch ops (cuDNN / cuBLAS / ATen of
    torch 2.11) at the workload's batch 32, 128x128 -- in the two modes a user of the reference
    would run: (i) fp32 with TF32 allowed, as reference train.py:67-68 sets it; (ii) bf16 autocast
    with channels_last images and weights.  CUDA events, `warm` untimed + `iters` timed."""
    out = {}
    dev = str(device)
    saved = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32,
             torch.get_float32_matmul_precision(), torch.backends.cudnn.benchmark)
    try:
        torch.backends.cudnn.allow_tf32 = True
        torch.backends.cuda.matmul.allow_tf32 = True
        torch.set_float32_matmul_precision("medium")
        for mode in ("tf32", "bf16"):
            from oracle import reference_port as rp

            arch = rp.Arch(image_size=IMAGE)
            tr = rp.Trainer(arch, rp.Hyper(batch_size=BATCH), rp.init_all(arch, 42), device=dev)
            if mode == "bf16":
                for net in tr.params.values():
                    for k, v in net.items():
                        if v.dim() == 4:
                            net[k] = v.contiguous(memory_format=torch.channels_last)
            bat = [rp.synthetic_batch(BATCH, arch, 7 + k).to(device) for k in range(4)]
            if mode == "bf16":
                bat = [b.contiguous(memory_format=torch.channels_last) for b in bat]

            def one():
                with torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
                    tr.discriminator_step(bat[0], bat[1])
                    tr.generator_step(bat[2], bat[3])

            for _ in range(warm):
                one()
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for _ in range(iters):
                one()
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / iters
            out[mode] = {"value": round(BATCH / ms * 1e3, 2), "ms_per_step": round(ms, 2)}
            del tr, bat
            torch.cuda.empty_cache()
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = saved[0], saved[1]
        torch.set_float32_matmul_precision(saved[2])
    out.update({"unit": "images/sec", "kind": "port on cuda:0 (eager torch %s, cuDNN %s)" % (
        torch.__version__, torch.backends.cudnn.version()),
        "sample": f"{warm} warm-up + {iters} timed D+G iterations, 128x128, batch {BATCH}; tf32 = fp32 "
                  "storage with TF32 convs/matmuls (reference train.py:67-68); bf16 = autocast + "
                  "channels_last.  The port skips two things the reference wastes (the G autograd "
                  "graph of the D step, D's weight gradients in the G step), so this bar is not "
                  "lower than the reference's own eager run."})
    return out


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU implementation of the path (oracle port; the
    reference itself is pure Python/PyTorch and is pinned to the port by tests/golden) on all
    host cores.  Each step is a bounded sample of the workload: one D+G iteration at batch 8
    (the workload's batch is 32; the CPU's img/s at batch 2 / 8 / 32 is in the b200 arm's
    `cpu_baseline.by_batch`)."""
    if rank != 0:
        return
    sample_batch = 8
    torch.set_num_threads(os.cpu_count() or 1)
    one = _oracle_trainer(sample_batch)
    for i in range(args.warmup):
        one(i)
    t0 = time.perf_counter()
    for i in range(args.steps):
        one(1000 + i)
    dt = time.perf_counter() - t0
    val = sample_batch * args.steps / dt
    sample = (f"each step = one D+G iteration at 128x128 on batch {sample_batch} (bounded sample of the "
              f"batch-{BATCH} workload), fp32, torch CPU, {torch.get_num_threads()} threads")
    line = {
        "impl": "reference", "metric": "G+D train-step images/sec", "value": round(val, 4),
        "unit": "images/sec", "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": round(1000 * dt / args.steps, 2), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": base_config(world),
        "cpu_baseline": {"value": round(val, 4), "unit": "images/sec", "cores": torch.get_num_threads(),
                         "kind": "port", "sample": sample},
        "e2e": {"value": round(val, 4), "unit": "images/sec", "h2d_bytes_per_step": 0,
                "d2h_bytes_per_step": 0},
    }
    if torch.cuda.is_available() and os.environ.get("OTM_REF_GPU", "1") == "1":
        try:  # the eager-PyTorch-on-B200 bar, reported next to the CPU number
            line["gpu_baseline"] = gpu_baseline(torch.device("cuda", 0), iters=5, warm=2)
        except Exception as e:  # noqa: BLE001 - the CPU arm must still print its line
            line["gpu_baseline"] = {"error": str(e)[:200]}
    print(json.dumps(line), flush=True)


def extra_config(device, rank, world, dist_on, image, batch, gflop, steps=5):
    """Another BASELINE config through the same engine (device-resident inputs, CUDA events,
    max over ranks): reported under `configs`, not the headline."""
    from one_to_many_gan_b200.synthetic import SyntheticImages

    cfg = make_config(image, batch)
    step, eng = build_trainer(device, rank, True, cfg, return_engine=True)
    prints = SyntheticImages(batch, 1, image, device, seed=42, rank=rank, stream_id=0)
    marks = SyntheticImages(batch, 1, image, device, seed=42, rank=rank, stream_id=1)
    ms, last = time_steps(step, prints, marks, steps, 4, dist_on, device)
    eng.close()
    value = world * batch * steps / (ms / 1e3)
    pk, _ = peaks()
    tf = gflop * value / world / 1e3
    return {"workload": f"{image[0]}x{image[1]}, batch {batch}/GPU, bf16", "value": round(value, 2),
            "unit": "images/sec", "ms_per_step": round(ms / steps, 3), "steps": steps,
            "conv_tflops_per_gpu": round(tf, 1),
            "conv_frac_of_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 4),
            "peak_mem_gib": round(torch.cuda.max_memory_allocated() / 2**30, 1),
            "losses_finite": all(v == v for v in last.values())}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))

    if args.impl == "reference":
        run_reference(args, rank, world)
        return

    import torch.distributed as dist

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    dist_on = world > 1
    verbose = os.environ.get("OTM_BENCH_VERBOSE", "0") == "1"

    def note(msg):
        if verbose:
            print(f"[bench rank {rank}] {msg}", file=sys.stderr, flush=True)

    if dist_on:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)
        note("process group up")

    from one_to_many_gan_b200 import kernels as K
    from one_to_many_gan_b200.synthetic import SyntheticImages

    warmup = max(args.warmup, 3)
    use_graph = os.environ.get("OTM_NO_GRAPH", "0") != "1"
    step, eng = build_trainer(device, rank, use_graph, return_engine=True)
    prints = SyntheticImages(BATCH, 1, IMAGE, device, seed=42, rank=rank, stream_id=0)
    marks = SyntheticImages(BATCH, 1, IMAGE, device, seed=42, rank=rank, stream_id=1)

    # ---- leg 1: inputs resident in HBM (synthetic generator on device) ---------------------
    sampler = ClockSampler(local) if rank == 0 else None
    # kernels per iteration: counted on the eager warm-up iterations (a replayed graph launches
    # the same kernels without going through the library's host entry points)
    l0 = K.launch_count()
    step(prints, marks)
    launches = K.launch_count() - l0
    note("first eager iteration done")
    ms, last = time_steps(step, prints, marks, args.steps, warmup, dist_on, device)
    clocks = sampler.stop() if sampler else None
    value = world * BATCH * args.steps / (ms / 1e3)
    note(f"device-resident leg done: {ms / args.steps:.2f} ms/step")

    # ---- leg 2: end to end through the public step API with pinned HOST batches -------------
    shape = (BATCH, 1, *IMAGE)
    hp, hm = HostBatches(shape, 1 + rank), HostBatches(shape, 1001 + rank)
    ms_e2e, _ = time_steps(step, hp, hm, args.steps, 1, dist_on, device)
    e2e_value = world * BATCH * args.steps / (ms_e2e / 1e3)
    h2d = 4 * BATCH * IMAGE[0] * IMAGE[1] * 4  # 2 batches per half-step x 2 half-steps, fp32
    d2h = 10 * 4  # the 3 + 7 logged scalars

    if rank == 0:
        pk, how = peaks()
        roof = dominant_kernel_roofline(device)
        conv_tflops = GFLOP_PER_IMAGE * value / world / 1e3
        line = {
            "metric": "G+D train-step images/sec", "value": round(value, 2), "unit": "images/sec",
            "n_gpus": world, "steps": args.steps, "warmup": warmup,
            "ms_per_step": round(ms / args.steps, 3), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": base_config(world),
            "execution": "one CUDA graph per iteration" if use_graph else "eager",
            "losses": {k: round(v, 5) for k, v in last.items()},
            "clocks": clocks,
            "e2e": {"value": round(e2e_value, 2), "unit": "images/sec", "h2d_bytes_per_step": h2d,
                    "d2h_bytes_per_step": d2h, "ms_per_step": round(ms_e2e / args.steps, 3)},
            "gpu_launches": int(launches * args.steps),
            "roofline": roof,
            "roofline_hbm": hbm_kernel_roofline(device),
            "conv_step_roofline": {
                "achieved_tflops_per_gpu": round(conv_tflops, 1), "peak": pk["bf16_tflops_sustained"],
                "peak_source": how + " sustained",
                "frac": round(conv_tflops / pk["bf16_tflops_sustained"], 4),
                "note": "algorithmic dense-conv FLOPs/image (SURVEY 8d) x images/s / sustained bf16 peak",
            },
        }
    # ---- BASELINE config 3 (256x256, batch 32 per GPU) through the same engine, every N -------
    eng.close()  # captured NCCL collectives must be released before the process group goes away
    del step, eng
    torch.cuda.empty_cache()
    extra = None
    if not args.no_extra_configs:
        extra = [extra_config(device, rank, world, dist_on, (256, 256), 32, GFLOP_256)]
        torch.cuda.empty_cache()
    if rank == 0:
        if extra:
            line["configs"] = extra
        if world == 1 and not args.no_gpu_baseline:
            line["gpu_baseline"] = gpu_baseline(device)
        if world == 1 and not args.no_cpu_baseline:
            line["cpu_baseline"] = cpu_baseline()
        print(json.dumps(line), flush=True)
    if dist_on:
        from one_to_many_gan_b200.optim import shutdown_process_group

        shutdown_process_group()


if __name__ == "__main__":
    main()
