"""The two pieces of the reference's src/core/evaluation.py the training loop itself needs:
the running-mean Logger (:269-308) and the checkpoint writer (:227-263, same dict keys so
checkpoints interchange with the reference).  FID/KID and image grids are evaluation-only and
out of scope for the hot path."""

from __future__ import annotations

import numpy as np
import torch


class Logger:
    def __init__(self, training_steps: int):
        self.training_steps = training_steps
        self.initialise_trackers()

    def initialise_trackers(self):
        self.log_total_disc_losses = []
        self.log_disc_real_accs = []
        self.log_disc_fake_accs = []
        self.log_total_gen_losses = []
        self.log_gan_losses = []
        self.log_idt_losses = []
        self.log_rec_losses = []
        self.log_kl_losses = []
        self.log_path_losses = []
        self.log_style_losses = []
        self.log_ada_ps = []

    def print(self, step: int):
        m = lambda v: f"{np.mean(v):.6g}"  # noqa: E731
        string = (
            f"Step: {step}/{self.training_steps}, "
            f"D loss: {m(self.log_total_disc_losses)}, "
            f"D real/fake acc: {m(self.log_disc_real_accs)}/{m(self.log_disc_fake_accs)}, "
            f"Total G loss: {m(self.log_total_gen_losses)}, "
            f"Gan loss {m(self.log_gan_losses)}, "
            f"Idt loss {m(self.log_idt_losses)}, "
            f"Rec loss {m(self.log_rec_losses)}, "
            f"KL loss {m(self.log_kl_losses)}, "
            f"Path loss {m(self.log_path_losses)}, "
            f"Style loss: {m(self.log_style_losses)}, "
            f"ADA: {m(self.log_ada_ps)}, "
        )
        self.initialise_trackers()
        return string


def model_checkpoint(step, config, generator, discriminator, mapping_network, style_extractor,
                     generator_optimiser, discriminator_optimiser, mapping_network_optimiser,
                     style_extractor_optimiser, ada_p, image_buffer):
    d = config["training"]["checkpoint_directory"] / config["training"]["training_run"] / "models"
    d.mkdir(parents=True, exist_ok=True)
    torch.save(
        {
            "generator_state_dict": generator.state_dict(),
            "generator_optim_state_dict": generator_optimiser.state_dict(),
            "discriminator_state_dict": discriminator.state_dict(),
            "discriminator_optim_state_dict": discriminator_optimiser.state_dict(),
            "mapping_network_state_dict": mapping_network.state_dict(),
            "mapping_network_optim_state_dict": mapping_network_optimiser.state_dict(),
            "style_extractor_state_dict": style_extractor.state_dict(),
            "style_extractor_optim_state_dict": style_extractor_optimiser.state_dict(),
            "ada_p": ada_p(),
            "image_buffer_images": image_buffer.images,
            "image_buffer_size": image_buffer.buffer_size,
        },
        d / f"{step + 1}.tar",
    )
