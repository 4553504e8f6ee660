"""Entry point with the command line and config keys of the reference's train.py:
    python train.py [config.toml]

One new optional key selects the execution path: `[training] backend = "b200"` (default)
runs the hand-written sm_100a path of this repository.  Further optional keys:
`[training] precision = "fp32" | "bf16"`, `[training] synthetic_data = true` (on-device
Philox batches instead of the image folders), `[architecture] start_filters`.
Under `torchrun` the same script trains data-parallel: one process per GPU, gradients
all-reduced over NCCL (the reference is single-GPU, train.py:61-65)."""

from __future__ import annotations

import itertools
import os
import random
import sys

import numpy as np
import torch


def main(config_path: str):
    from one_to_many_gan_b200 import builder, training
    from one_to_many_gan_b200.config import act_dtype, load_config
    from one_to_many_gan_b200.evaluation import Logger, model_checkpoint
    from one_to_many_gan_b200.optim import FlatAdam
    from one_to_many_gan_b200.synthetic import SyntheticImages

    config = load_config(config_path)
    if config["training"]["backend"] != "b200":
        raise SystemExit(
            f"backend {config['training']['backend']!r}: this repository ships only the 'b200' path; "
            "run the reference's own train.py for its PyTorch path"
        )

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

    # ---- seeding (reference train.py:35-37) --------------------------------------------------
    seed = config["training"]["random_seed"]
    torch.manual_seed(seed)
    np.random.default_rng(seed)
    random.seed(seed)
    torch.cuda.manual_seed_all(seed)

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

    o = config["optimisation"]
    betas = tuple(o["adam_betas"])
    discriminator_optimiser = FlatAdam(discriminator.parameters(), o["learning_rate"], betas)
    generator_optimiser = FlatAdam(generator.parameters(), o["learning_rate"], betas)
    mapping_network_optimiser = FlatAdam(
        mapping_network.parameters(), o["mapping_network_learning_rate"], betas
    )
    style_extractor_optimiser = FlatAdam(style_extractor.parameters(), o["learning_rate"], betas)

    # ---- data ---------------------------------------------------------------------------------
    batch = config["training"]["batch_size"]
    if rank != 0:  # decorrelate the per-rank style / theta draws
        torch.manual_seed(seed + rank)
        random.seed(seed + rank)
    if config["training"]["synthetic_data"]:
        shoeprint_iter = SyntheticImages(batch, data["image_channels"], data["image_size"], device,
                                         seed=seed, rank=rank, stream_id=0)
        shoemark_iter = SyntheticImages(batch, data["image_channels"], data["image_size"], device,
                                        seed=seed, rank=rank, stream_id=1)
    else:
        from one_to_many_gan_b200.datasets import image_folder_loader

        shoemark_iter = itertools.cycle(image_folder_loader(config, "shoemark_data_dir", seed))
        shoeprint_iter = itertools.cycle(image_folder_loader(config, "shoeprint_data_dir", seed))

    image_buffer = training.ImageBuffer(config["training"]["image_buffer_size"])
    ada = training.IdentityAugment().to(device)
    ada_p = training.ADAp(
        ada_e=config["ada"]["ada_overfitting_measurement_n_images"],
        ada_adjustment_size=config["ada"]["ada_adjustment_size"],
        batch_size=batch,
        discriminator_overfitting_target=config["ada"]["discriminator_real_acc_target"],
    )
    logger = Logger(config["training"]["training_steps"])

    steps = config["training"]["training_steps"]
    for step in range(steps):
        p = ada_p()
        # the augmentation pipeline itself (pytorch-ada) is outside the hot path: p stays 0
        ada.set_p(0.0)
        logger.log_ada_ps.append(p)
        disc_loss, (real_acc, fake_acc) = training.discriminator_step(
            config, device, discriminator, generator, mapping_network, discriminator_optimiser,
            shoeprint_iter, shoemark_iter, image_buffer, ada, ada_p,
        )
        logger.log_total_disc_losses.append(disc_loss)
        logger.log_disc_real_accs.append(real_acc)
        logger.log_disc_fake_accs.append(fake_acc)
        total, (gan, rec, idt, kl, path, style) = training.generator_step(
            config, device, generator, discriminator, mapping_network, style_extractor,
            generator_optimiser, mapping_network_optimiser, style_extractor_optimiser,
            shoeprint_iter, shoemark_iter, ada,
        )
        logger.log_total_gen_losses.append(total)
        logger.log_gan_losses.append(gan)
        logger.log_rec_losses.append(rec)
        logger.log_idt_losses.append(idt)
        logger.log_kl_losses.append(kl)
        logger.log_path_losses.append(path)
        logger.log_style_losses.append(style)

        last = (step + 1) == steps
        if rank == 0 and ((step + 1) % config["evaluation"]["log_interval"] == 0 or last):
            log = logger.print(step + 1)
            print(log, flush=True)
            log_dir = config["training"]["checkpoint_directory"] / config["training"]["training_run"]
            log_dir.mkdir(parents=True, exist_ok=True)
            with (log_dir / "log").open("a") as f:
                f.write(log + "\n")
        if rank == 0 and ((step + 1) % config["evaluation"]["checkpoint_interval"] == 0 or last):
            model_checkpoint(
                step, config, generator, discriminator, mapping_network, style_extractor,
                generator_optimiser, discriminator_optimiser, mapping_network_optimiser,
                style_extractor_optimiser, ada_p, image_buffer,
            )
    if world > 1:
        import torch.distributed as dist

        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main("config.toml" if len(sys.argv) < 2 or sys.argv[1] == "" else sys.argv[1])
