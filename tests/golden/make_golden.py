"""Generate golden vectors by running the UNMODIFIED reference in this container.

    PYTHONDONTWRITEBYTECODE=1 python tests/golden/make_golden.py

Imports `src.model.*` / `src.core.training` from /root/reference (read-only),
with a 6-line identity stand-in for the un-installed third-party `ada` package
(exact while ADA p == 0, reference loss.py:22,28,33), builds D, G, M, S and the
four Adams exactly as reference train.py:35,72-116 does, feeds seeded synthetic
U(-1,1) batches, runs `discriminator_step` + `generator_step` for a few
iterations and records losses, outputs and per-tensor fingerprints of gradients
and updated weights.  The fixtures (`tests/golden/*.pt`) are committed; the GPU
box never sees /root/reference.
"""

from __future__ import annotations

import os
import random
import sys
import tempfile
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
REF = Path("/root/reference")

CASES = {
    # name: (image_size, min_latent, n_resnet_blocks, batch, iterations)
    "a_32x32_down1": ((32, 32), 16, 3, 2, 3),
    "b_64x64_default": ((64, 64), 64, 7, 4, 3),
    "c_32x48_down2": ((32, 48), 8, 2, 2, 2),
    # BASELINE configs[0] literally: 64x64 (smallest designed resolution), batch 4, 20 iterations
    "d_64x64_config1_20it": ((64, 64), 64, 7, 4, 20),
}


def fingerprint(t: torch.Tensor) -> torch.Tensor:
    """[sum, abs-sum, l2, 16 strided samples] in float64."""
    f = t.detach().double().reshape(-1)
    n = f.numel()
    idx = torch.linspace(0, n - 1, 16).long()
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), f.norm()]), f[idx]])


def batches(shape, seed):
    g = torch.Generator().manual_seed(seed)
    while True:
        yield torch.rand(*shape, generator=g) * 2 - 1


def run_case(name, image_size, min_latent, n_res, batch, iters):
    from src.core.training import ImageBuffer, discriminator_step, generator_step
    from src.data.config import load_config
    from src.model.builder import Discriminator, Generator, MappingNetwork, StyleExtractor
    from src.model.loss import ADAp
    from ada import AdaptiveDiscriminatorAugmentation

    cfg = load_config(REF / "config.toml")
    cfg["training"]["batch_size"] = batch
    cfg["data"]["image_size"] = list(image_size)
    cfg["architecture"]["min_latent_resolution"] = min_latent
    cfg["architecture"]["n_resnet_blocks"] = n_res
    seed = cfg["training"]["random_seed"]
    torch.manual_seed(seed)
    random.seed(seed)
    dev = torch.device("cpu")
    D = Discriminator(input_nc=1)
    G = Generator(1, cfg["architecture"]["w_dim"], tuple(image_size), min_latent, n_res)
    M = MappingNetwork(cfg["architecture"]["w_dim"], 2, cfg["training"]["style_mixing_prob"])
    S = StyleExtractor(1, cfg["architecture"]["w_dim"])
    lr, betas = cfg["optimisation"]["learning_rate"], cfg["optimisation"]["adam_betas"]
    oD = torch.optim.Adam(D.parameters(), lr=lr, betas=betas)
    oG = torch.optim.Adam(G.parameters(), lr=lr, betas=betas)
    oM = torch.optim.Adam(
        M.parameters(), lr=cfg["optimisation"]["mapping_network_learning_rate"], betas=betas
    )
    oS = torch.optim.Adam(S.parameters(), lr=lr, betas=betas)
    ada = AdaptiveDiscriminatorAugmentation()
    ada_p = ADAp(256, cfg["ada"]["ada_adjustment_size"], batch, 0.6)
    buf = ImageBuffer(cfg["training"]["image_buffer_size"])
    shape = (batch, 1, *image_size)
    prints, marks = batches(shape, 1000), batches(shape, 2000)

    out = {
        "meta": dict(
            image_size=tuple(image_size), min_latent=min_latent, n_res=n_res, batch=batch,
            iters=iters, seed=seed, print_seed=1000, mark_seed=2000,
            n_style_blocks=G.n_style_blocks,
        ),
        "init_fp": {
            n: {k: fingerprint(v) for k, v in mod.state_dict().items()}
            for n, mod in (("D", D), ("G", G), ("M", M), ("S", S))
        },
    }

    # forward-only known answers on fixed inputs and a fixed style (no RNG draws)
    gx = torch.Generator().manual_seed(7)
    x = torch.rand(*shape, generator=gx) * 2 - 1
    w = torch.rand(G.n_style_blocks, batch, cfg["architecture"]["w_dim"], generator=gx)
    with torch.no_grad():
        z = G.encode(x)
        y = G.decode(z, w)
        feats = G.extract(z, w)
        out["forward"] = {
            "x_seed": 7,
            "latent_fp": fingerprint(z),
            "g_out": y.clone() if y.numel() <= 8192 else fingerprint(y),
            "g_out_fp": fingerprint(y),
            "feat_fp": [fingerprint(f) for f in feats],
            "d_out": D(x).clone(),
            "s_out": S(x).clone(),
            "m_out": M(w[0]).clone(),
        }

    losses = []
    for it in range(iters):
        d = discriminator_step(cfg, dev, D, G, M, oD, prints, marks, buf, ada, ada_p)
        if it == 0:
            out["d_grad_fp"] = {k: fingerprint(p.grad) for k, p in D.named_parameters()}
        g = generator_step(cfg, dev, G, D, M, S, oG, oM, oS, prints, marks, ada)
        if it == 0:
            out["g_grad_fp"] = {
                n: {k: fingerprint(p.grad) for k, p in mod.named_parameters()}
                for n, mod in (("G", G), ("M", M), ("S", S))
            }
            out["after1_fp"] = {
                n: {k: fingerprint(v) for k, v in mod.state_dict().items()}
                for n, mod in (("D", D), ("G", G), ("M", M), ("S", S))
            }
        losses.append([d[0], d[1][0], d[1][1], g[0], *g[1]])
    out["losses"] = torch.tensor(losses, dtype=torch.float64)
    out["final_fp"] = {
        n: {k: fingerprint(v) for k, v in mod.state_dict().items()}
        for n, mod in (("D", D), ("G", G), ("M", M), ("S", S))
    }
    if iters >= 10:
        # Long runs are chaotic: from iteration ~3 on, ANY two fp32 executions (and the
        # reference against its own fp64 arithmetic) only agree statistically.  Record the
        # port's fp64 run of the same replay so tests can gate each iteration against the
        # measured fp32-vs-fp64 divergence instead of a made-up tolerance.
        out["losses_port_fp64"] = port_losses(image_size, min_latent, n_res, batch, iters, seed)
    torch.save(out, HERE / f"{name}.pt")
    print(name, "losses[0] =", [f"{v:.6g}" for v in losses[0]])


def port_losses(image_size, min_latent, n_res, batch, iters, seed):
    sys.path.insert(0, str(HERE.parents[1]))
    from oracle import reference_port as rp

    arch = rp.Arch(image_size=tuple(image_size), min_latent_resolution=min_latent,
                   n_resnet_blocks=n_res)
    params = rp.init_all(arch, seed)
    tr = rp.Trainer(arch, rp.Hyper(batch_size=batch), params, dtype=torch.float64)
    shape = (batch, 1, *image_size)
    prints, marks = batches(shape, 1000), batches(shape, 2000)
    rows = []
    for _ in range(iters):
        d = tr.discriminator_step(next(prints), next(marks))
        g = tr.generator_step(next(prints), next(marks))
        rows.append([d[0], d[1][0], d[1][1], g[0], *g[1]])
    return torch.tensor(rows, dtype=torch.float64)


def main():
    os.environ["PYTHONDONTWRITEBYTECODE"] = "1"
    sys.dont_write_bytecode = True
    with tempfile.TemporaryDirectory() as shim:
        Path(shim, "ada.py").write_text(
            "import torch\n"
            "class AdaptiveDiscriminatorAugmentation(torch.nn.Module):\n"
            "    def __init__(self, **kw):\n        super().__init__()\n"
            "    def set_p(self, p):\n        self.p = p\n"
            "    def forward(self, x):\n        return x\n"
        )
        sys.path[:0] = [shim, str(REF)]
        torch.set_num_threads(os.cpu_count() or 1)
        only = sys.argv[1:]
        for name, args in CASES.items():
            if not only or name in only:
                run_case(name, *args)


if __name__ == "__main__":
    main()
