"""Entry point with the command line and config keys of the reference's train.py:
    python train.py [config.toml]

ONE config flag selects the execution path: `[training] backend = "b200"` (default) runs the
hand-written sm_100a path of this repository; `backend = "torch"` runs the reference's own
PyTorch modules and step functions (imported from the reference checkout on `sys.path` or at
`[training] reference_path`) inside the same loop.  Further optional keys (defaults = the
reference's behaviour): `[training] precision = "fp32" | "bf16"`, `synthetic_data = true`
(on-device Philox batches instead of the image folders), `execution = "graph" | "eager"`,
`resume = true | "<file>"`, `styles_per_input`; `[optimisation] r1_gamma`;
`[architecture] start_filters`; `[ada] allow_identity`.
Under `torchrun` the b200 backend trains data-parallel: one process per GPU, gradients
all-reduced over NCCL (the reference is single-GPU, train.py:61-65)."""

from __future__ import annotations

import itertools
import os
import random
import sys

import numpy as np
import torch


def _seed_everything(seed: int) -> None:
    """reference train.py:35-37"""
    torch.manual_seed(seed)
    np.random.default_rng(seed)
    random.seed(seed)
    if torch.cuda.is_available():
        torch.cuda.manual_seed_all(seed)


def _data_iterators(config, device, rank: int, world: int):
    from one_to_many_gan_b200.synthetic import SyntheticImages

    data, seed, batch = config["data"], config["training"]["random_seed"], config["training"]["batch_size"]
    if config["training"]["synthetic_data"]:
        mk = lambda sid: SyntheticImages(batch, data["image_channels"], data["image_size"], device,  # noqa: E731
                                         seed=seed, rank=rank, stream_id=sid)
        return mk(0), mk(1)
    from one_to_many_gan_b200.datasets import image_folder_loader

    # ONE generator shared by both loaders, like the reference's `dataloader_g`
    # (train.py:56-58,139,153): the two folders are shuffled independently; every rank draws the
    # same permutations and takes its own slice of each
    gen = torch.Generator().manual_seed(seed)
    marks = image_folder_loader(config, "shoemark_data_dir", gen, rank, world, device)
    prints = image_folder_loader(config, "shoeprint_data_dir", gen, rank, world, device)
    return prints, marks  # DeviceImages cycles over epochs itself


def main_torch(config):
    """backend = "torch": the reference's modules and step functions, unmodified, in this loop."""
    ref = config["training"].get("reference_path")
    if ref and str(ref) not in sys.path:
        sys.path.insert(0, str(ref))
    try:
        from src.core.training import ImageBuffer, discriminator_step, generator_step
        from src.model.builder import Discriminator, Generator, MappingNetwork, StyleExtractor
        from src.model.loss import ADAp
    except ImportError as e:
        raise SystemExit(
            "backend 'torch' runs the reference's own PyTorch path: put the reference checkout (and "
            f"its `ada` dependency) on PYTHONPATH or set [training] reference_path ({e})") from e
    try:
        from ada import AdaptiveDiscriminatorAugmentation as Ada
    except ImportError as e:
        raise SystemExit(f"backend 'torch' needs the reference's pytorch-ada dependency ({e})") from e
    from one_to_many_gan_b200.checkpoint import RunLog

    t, arch, data, o = config["training"], config["architecture"], config["data"], config["optimisation"]
    _seed_everything(t["random_seed"])
    device = torch.device(f"cuda:{t['gpu_number']}" if torch.cuda.is_available() else "cpu")
    torch.set_float32_matmul_precision("medium")  # reference train.py:67-68
    torch.backends.cudnn.allow_tf32 = True
    D = Discriminator(input_nc=data["image_channels"]).to(device)
    G = Generator(data["image_channels"], arch["w_dim"], tuple(data["image_size"]),
                  arch["min_latent_resolution"], arch["n_resnet_blocks"]).to(device)
    M = MappingNetwork(arch["w_dim"], arch["mapping_network_layers"], t["style_mixing_prob"]).to(device)
    S = StyleExtractor(data["image_channels"], arch["w_dim"]).to(device)
    betas = tuple(o["adam_betas"])
    oD = torch.optim.Adam(D.parameters(), lr=o["learning_rate"], betas=betas)
    oG = torch.optim.Adam(G.parameters(), lr=o["learning_rate"], betas=betas)
    oM = torch.optim.Adam(M.parameters(), lr=o["mapping_network_learning_rate"], betas=betas)
    oS = torch.optim.Adam(S.parameters(), lr=o["learning_rate"], betas=betas)
    if t["synthetic_data"]:  # plain torch: this branch must also run where the b200 library cannot
        def synth(seed):
            g = torch.Generator().manual_seed(seed)
            while True:
                yield torch.rand(t["batch_size"], data["image_channels"], *data["image_size"], generator=g) * 2 - 1

        prints, marks = synth(t["random_seed"]), synth(t["random_seed"] + 1)
    else:  # the reference's own dataset + DataLoaders (train.py:120-169)
        from src.data.datasets import ShoeDataset
        from torchvision import transforms

        tf = transforms.Compose([transforms.Resize(data["image_size"]), transforms.ToTensor(),
                                 transforms.Normalize((0.5,), (0.5,))])
        gen = torch.Generator().manual_seed(t["random_seed"])

        def loader(key):
            ds = ShoeDataset(data[key], mode="train", transform=tf)
            return itertools.cycle(torch.utils.data.DataLoader(
                ds, batch_size=t["batch_size"], shuffle=True, num_workers=8, drop_last=True,
                pin_memory=True, generator=gen))

        marks, prints = loader("shoemark_data_dir"), loader("shoeprint_data_dir")
    buf = ImageBuffer(t["image_buffer_size"])
    aug = ("xflip", "rotate90", "xint", "scale", "rotate", "aniso", "xfrac", "brightness", "contrast",
           "lumaflip", "hue", "saturation")  # every pipeline stage on, reference train.py:175-188
    ada = Ada(**dict.fromkeys(aug, 1)).to(device)
    ada_p = ADAp(config["ada"]["ada_overfitting_measurement_n_images"], config["ada"]["ada_adjustment_size"],
                 t["batch_size"], config["ada"]["discriminator_real_acc_target"])
    log = RunLog(t["training_steps"])
    for step in range(t["training_steps"]):
        ada.set_p(ada_p())
        d, (ra, fa) = discriminator_step(config, device, D, G, M, oD, prints, marks, buf, ada, ada_p)
        g, (gan, rec, idt, kl, path, style) = generator_step(config, device, G, D, M, S, oG, oM, oS,
                                                             prints, marks, ada)
        log.record(disc=d, sign_real=ra, sign_fake=fa, total_gen=g, gan=gan, rec=rec, idt=idt, kl=kl,
                   path=path, style=style, ada_p=ada_p())
        if (step + 1) % config["evaluation"]["log_interval"] == 0 or step + 1 == t["training_steps"]:
            print(log.line(step + 1), flush=True)


def main(config_path: str):
    from one_to_many_gan_b200.config import act_dtype, load_config

    config = load_config(config_path)
    backend = config["training"]["backend"]
    if backend == "torch":
        return main_torch(config)
    if backend != "b200":
        raise SystemExit(f"[training] backend must be 'b200' or 'torch', got {backend!r}")

    from one_to_many_gan_b200 import builder, training
    from one_to_many_gan_b200.checkpoint import (RunLog, checkpoint_path, latest_checkpoint,
                                                 load_checkpoint, save_checkpoint)
    from one_to_many_gan_b200.engine import TrainIteration
    from one_to_many_gan_b200.optim import FlatAdam

    # ---- distributed (reference has none; torchrun-style env) --------------------------------
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", str(config["training"]["gpu_number"])))
    if not torch.cuda.is_available():
        raise SystemExit("the b200 backend needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    device = torch.device("cuda", local)
    if world > 1:
        import torch.distributed as dist

        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=device)

    seed = config["training"]["random_seed"]
    _seed_everything(seed)

    # ---- models, in the reference's construction (= RNG draw) order: D, G, M, S -------------
    dt = act_dtype(config)
    arch, data = config["architecture"], config["data"]
    discriminator = builder.Discriminator(input_nc=data["image_channels"], act_dtype=dt).to(device)
    generator = builder.Generator(
        input_nc=data["image_channels"], w_dim=arch["w_dim"], image_size=data["image_size"],
        min_latent_resolution=arch["min_latent_resolution"], n_resnet_blocks=arch["n_resnet_blocks"],
        start_filters=arch["start_filters"], act_dtype=dt,
    ).to(device)
    mapping_network = builder.MappingNetwork(
        features=arch["w_dim"], n_layers=arch["mapping_network_layers"],
        style_mixing_prob=config["training"]["style_mixing_prob"],
    ).to(device)
    style_extractor = builder.StyleExtractor(
        input_nc=data["image_channels"], w_dim=arch["w_dim"], act_dtype=dt
    ).to(device)
    nets = {"generator": generator, "discriminator": discriminator,
            "mapping_network": mapping_network, "style_extractor": style_extractor}

    o = config["optimisation"]
    betas = tuple(o["adam_betas"])
    opts = {n: FlatAdam(m.parameters(), o["mapping_network_learning_rate" if n == "mapping_network"
                                          else "learning_rate"], betas) for n, m in nets.items()}

    batch = config["training"]["batch_size"]
    if rank != 0:  # decorrelate the per-rank style / theta / pool draws
        torch.manual_seed(seed + rank)
        random.seed(seed + rank)
    shoeprint_iter, shoemark_iter = _data_iterators(config, device, rank, world)

    ada_p = training.ADAp(
        ada_e=config["ada"]["ada_overfitting_measurement_n_images"],
        ada_adjustment_size=config["ada"]["ada_adjustment_size"],
        batch_size=batch,
        discriminator_overfitting_target=config["ada"]["discriminator_real_acc_target"],
    )
    eng = TrainIteration(config, device, discriminator, generator, mapping_network, style_extractor,
                         opts["discriminator"], opts["generator"], opts["mapping_network"],
                         opts["style_extractor"],
                         use_graph=config["training"].get("execution", "graph") == "graph")

    # ---- resume (the loader the reference lacks) ------------------------------------------------
    start = 0
    resume = config["training"].get("resume", False)
    if resume:
        path = latest_checkpoint(config) if resume is True else resume
        if path is not None:
            st = load_checkpoint(path, nets=nets, optimisers=opts, ada_p=ada_p, device=device,
                                 restore_rng=(rank == 0 and world == 1))
            eng.load_pool(st["pool_images"])
            start = st["step"]
            for it in (shoeprint_iter, shoemark_iter):  # synthetic streams are counter-based
                if hasattr(it, "step"):
                    it.step = 2 * start
            if rank == 0:
                print(f"resumed from {path} at step {start}", flush=True)

    steps = config["training"]["training_steps"]
    log = RunLog(steps)
    allow_identity = bool(config["ada"].get("allow_identity", False))
    for step in range(start, steps):
        p = ada_p()
        if p > 0 and not allow_identity:
            # The augmentation pipeline (third-party pytorch-ada) is outside this path: the b200
            # backend trains WITHOUT augmentation.  Never report or checkpoint a p that is not applied.
            raise SystemExit(
                f"step {step}: the ADA controller asks for p = {p:.4g} but the b200 backend has no "
                "augmentation pipeline; set `[ada] allow_identity = true` to train un-augmented "
                "(p is then logged and saved as 0), or use backend = 'torch'")
        eng.load_inputs(next(shoeprint_iter), next(shoemark_iter), next(shoeprint_iter), next(shoemark_iter))
        out = eng.run(h="host")  # every random draw of the iteration comes from the host generators
        ada_p.update_p(torch.tensor(out["sign_real"]))
        if allow_identity:
            ada_p.p = torch.zeros(())  # what is applied is what is logged and saved
        log.record(**out, ada_p=0.0 if allow_identity else p)

        last = (step + 1) == steps
        if rank == 0 and ((step + 1) % config["evaluation"]["log_interval"] == 0 or last):
            line = log.line(step + 1)
            print(line, flush=True)
            log_dir = config["training"]["checkpoint_directory"] / config["training"]["training_run"]
            log_dir.mkdir(parents=True, exist_ok=True)
            with (log_dir / "log").open("a") as f:
                f.write(line + "\n")
        if rank == 0 and ((step + 1) % config["evaluation"]["checkpoint_interval"] == 0 or last):
            save_checkpoint(checkpoint_path(config, step + 1), nets=nets, optimisers=opts, ada_p=ada_p,
                            pool_images=eng.pool_images(), pool_size=eng.pool_size, step=step + 1)
    if world > 1:
        from one_to_many_gan_b200.optim import shutdown_process_group

        eng.close()
        shutdown_process_group()
    return eng


if __name__ == "__main__":
    main("config.toml" if len(sys.argv) < 2 or sys.argv[1] == "" else sys.argv[1])
