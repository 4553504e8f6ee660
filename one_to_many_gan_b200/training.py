"""The two halves of one training iteration, with the signatures and return values of the
reference's src/core/training.py (`discriminator_step` :71-128, `generator_step` :136-257),
executing on the B200 kernels.

What is kept: the order of host RNG draws, every loss definition and weight,
the ImageBuffer / ADAp host logic, the (float, tuple-of-floats) return values.
What is scheduled differently (same mathematics):
  * the generator forward inside the D step runs without building an autograd graph (the
    reference builds one and discards it, training.py:98);
  * fake and real images go through the discriminator as ONE 2B batch, the three image
    decodes (reconstruction, identity, translation) as ONE 3B batch and the two path-length
    extractions as ONE 2B batch (InstanceNorm and the style modulation are per-sample, so the
    results are identical);
  * discriminator weight gradients are not computed in the G step (the reference computes and
    then zeroes them, training.py:245 vs :88);
  * every loss weight is folded into the kernel that writes its backward seed, and the 3+7
    logged scalars come back in one device->host copy instead of ten."""

from __future__ import annotations

import random
from collections.abc import Iterator

import torch
import torch.nn.functional as F

from . import ops
from .builder import Discriminator, Generator, MappingNetwork, StyleExtractor


class ImageBuffer:
    """History pool of generated images (reference training.py:22-65)."""

    def __init__(self, buffer_size: int):
        self.buffer_size = buffer_size
        if self.buffer_size < 1:
            raise ValueError
        self.num_imgs = 0
        self.images: list[torch.Tensor] = []

    def __call__(self, images: torch.Tensor):
        out = []
        for image in images:
            image = torch.unsqueeze(image.detach(), 0)
            if self.num_imgs < self.buffer_size:
                self.num_imgs += 1
                self.images.append(image)
                out.append(image)
            else:
                if random.uniform(0, 1) > 0.5:
                    k = random.randint(0, self.buffer_size - 1)
                    out.append(self.images[k].clone())
                    self.images[k] = image
                else:
                    out.append(image)
        return torch.cat(out, 0)


class ADAp:
    """Adaptive-augmentation probability controller (reference loss.py:11-52, host side)."""

    def __init__(self, ada_e: float, ada_adjustment_size: float, batch_size: int,
                 discriminator_overfitting_target: float):
        self.n_batches = ada_e // batch_size
        self.ada_adjustment = ada_adjustment_size * ada_e
        self.overfitting_target = discriminator_overfitting_target
        self.p = torch.zeros(())
        self.curr_batch = 0
        self.mean_real_scores: list[torch.Tensor] = []

    def update_p(self, mean_score: torch.Tensor):
        if self.curr_batch == self.n_batches:
            self.mean_real_scores.append(mean_score)
            mean_sign = torch.mean(torch.stack(self.mean_real_scores)).cpu()
            if mean_sign < self.overfitting_target:
                self.p -= self.ada_adjustment
            elif mean_sign > self.overfitting_target:
                self.p += self.ada_adjustment
            self.curr_batch = 0
            self.mean_real_scores = []
            self.p = F.relu(self.p, inplace=True)
        self.curr_batch += 1
        self.mean_real_scores.append(mean_score)

    def __call__(self) -> float:
        return self.p.item()


class IdentityAugment(torch.nn.Module):
    """Stand-in with the call surface of pytorch-ada's AdaptiveDiscriminatorAugmentation
    (reference train.py:175-188,206).  Exact while ADA p == 0; the augmentation pipeline itself
    is out of scope for the hot path."""

    def __init__(self, **_kw):
        super().__init__()
        self.p = 0.0

    def set_p(self, p):
        if p != 0:
            raise NotImplementedError("non-zero ADA probability needs the pytorch-ada pipeline")
        self.p = p

    def forward(self, x):
        return x


def style_cycle_loss_func(original_w, reconstructed_w, *, normalise=True, cos_l2_ratio: float = 0.2):
    """Reference loss.py:60-75 on [B, w_dim] tensors: one kernel writes the scalar and both
    backward seeds (ops.style_cycle)."""
    if not normalise:
        raise ValueError("the B200 path implements normalise=True (all the reference uses)")
    return ops.style_cycle(original_w, reconstructed_w, 1.0, cos_l2_ratio)[0].reshape(())


def _floats(*tensors) -> list[float]:
    """One device->host copy for all logged scalars."""
    return torch.stack([t.detach().reshape(()).float() for t in tensors]).cpu().tolist()


# ---------------------------------------------------------------------------
# device-side cores shared by the reference-signature step functions below and by the
# CUDA-graph engine (engine.py).  No host synchronisation, no host RNG in here.
# ---------------------------------------------------------------------------
def r1_gamma(config) -> float:
    """`[optimisation] r1_gamma` (not a reference key; default 0 = the reference): weight of the
    R1 gradient penalty on real images, BASELINE config 5."""
    return float(config.get("optimisation", {}).get("r1_gamma", 0.0))


def discriminator_losses(discriminator, fake, real, r1_gamma: float = 0.0):
    """(reference training.py:107-117) one 2B discriminator pass; returns
    (disc_loss, sign_real, sign_fake) as 1-element tensors.  r1_gamma > 0 adds
    gamma/2 * E||grad_x D(real)||^2 (r1.py) to disc_loss."""
    batch = real.shape[0]
    scores = discriminator(torch.cat([fake, real], dim=0))
    fake_scores, real_scores = scores[:batch], scores[batch:]
    # disc_loss = (mse(real, 1) + mse(fake, 0)) / 2
    real_loss, sign_real, _ = ops.lsgan(real_scores, 1.0, 0.5)
    fake_loss, sign_fake, _ = ops.lsgan(fake_scores, 0.0, 0.5)
    loss = real_loss + fake_loss
    if r1_gamma > 0:
        from . import r1

        loss = loss + r1.r1_penalty(discriminator, real, r1_gamma)
    return loss, sign_real, -sign_fake


def styles_per_input(config) -> int:
    """`[training] styles_per_input = K` (not a reference key; default 1 = the reference):
    BASELINE config 4, one input -> K sampled outputs per step."""
    k = int(config["training"].get("styles_per_input", 1))
    if k < 1:
        raise ValueError("styles_per_input must be >= 1")
    return k


def generator_losses(config, generator, discriminator, style_extractor, prints, marks,
                     reconstruct_w, translation_w, w1, w2, cent_fin_diff_h, ada=None,
                     latent_noise=None, before_discriminator=None, on_latent_grad=None):
    """(reference training.py:158-243) all generator-side losses; each lambda is folded into the
    kernel that writes the term's backward seed.  Returns (total, gan, rec, idt, kl, path, style)
    as 1-element tensors: `total` is the weighted sum that is differentiated, the six terms are
    the RAW (unweighted, detached) values the reference logs (training.py:250-257).

    With K = translation_w.shape[1] / B > 1 styles per input, every sampled-style pass
    (translation decode, D and S on the translations, both path-length extractions) runs on the
    shoeprint latents broadcast to K styles (image (b, k) at index b*K + k, the
    `latent.expand(K, ...)` of evaluation.py:172-177); the latent gradient is the sum over the
    K decodes.  Reconstruction (w = 0) and identity stay at B.

    Two scheduling hooks for the data-parallel engine (both optional, no effect on the maths):
    `before_discriminator()` is called right before the first use of the discriminator (the
    deferred Adam(D) of the D step goes there); `on_latent_grad()` is called by autograd when the
    gradient of the latents is complete, i.e. when every decoder / style-extractor gradient is
    final and only the encoder backward remains."""
    opt = config["optimisation"]
    batch = prints.shape[0]
    nb = generator.n_style_blocks
    n_sty = translation_w.shape[1] // batch
    if translation_w.shape[1] != batch * n_sty or w1.shape[1] != batch * n_sty:
        raise ValueError("style batch must be a multiple of the image batch")
    combined_latents = generator.encode(torch.cat([prints, marks], dim=0))
    if on_latent_grad is not None and combined_latents.requires_grad:
        def _latents_ready(_g):
            on_latent_grad()

        combined_latents.register_hook(_latents_ready)
    kl_loss, kl_raw = ops.kl(combined_latents, opt["kl_loss_lambda"])
    if config["architecture"]["add_latent_noise"]:
        # (training.py:166-167) the caller pre-draws the noise so the device generator is
        # consumed in the reference's order: latent noise BEFORE the finite-difference step h
        if latent_noise is None:
            latent_noise = torch.randn_like(combined_latents)
        combined_latents = combined_latents + latent_noise.to(combined_latents.dtype)
    shoeprint_latent, shoemark_latent = combined_latents.chunk(2, dim=0)

    real_shoemark_w = style_extractor(marks)
    identity_w = real_shoemark_w.expand(nb, *real_shoemark_w.shape)

    # reconstruction / identity / translation as one 3B decode
    # (K = 1 and a halo-carrying latent: the batches are stacked as copies of the padded NHWC
    # buffer, ops.stack_halo; otherwise plain torch.cat, which decode() re-lays out)
    stacked = n_sty == 1 and ops.halo_of(combined_latents) > 0
    shoeprint_latent_k = (shoeprint_latent if n_sty == 1
                          else shoeprint_latent.repeat_interleave(n_sty, dim=0))
    if stacked:
        dec_latents = ops.stack_halo(combined_latents, [(0, 2 * batch), (0, batch)])
    else:
        dec_latents = torch.cat([shoeprint_latent, shoemark_latent, shoeprint_latent_k], dim=0)
    dec_w = torch.cat([reconstruct_w, identity_w, translation_w], dim=1)
    images = generator.decode(dec_latents, dec_w)
    reconstruction_loss, rec_raw = ops.l1(images[:batch], prints, opt["reconstruction_loss_lambda"])
    identity_loss, idt_raw = ops.l1(images[batch : 2 * batch], marks, opt["identity_loss_lambda"])
    generated_shoemarks = images[2 * batch :]

    # GAN loss (the discriminator's own weight gradients are not needed here)
    if before_discriminator is not None:
        before_discriminator()
    d_params = [p for p in discriminator.parameters() if p.requires_grad]
    for p in d_params:
        p.requires_grad_(False)
    try:
        fake_in = generated_shoemarks if ada is None else ada(generated_shoemarks)
        fake_shoemark_scores = discriminator(fake_in)
    finally:
        for p in d_params:
            p.requires_grad_(True)
    gan_loss, _, gan_raw = ops.lsgan(fake_shoemark_scores, 1.0, 1.0)

    reconstructed_w = style_extractor(generated_shoemarks)
    style_loss, style_raw = ops.style_cycle(translation_w[-1], reconstructed_w,
                                            opt["style_cycle_loss_lambda"])

    # path length: two extractions of the same latent as one 2B batch
    ext_latents = (ops.stack_halo(combined_latents, [(0, batch), (0, batch)]) if stacked
                   else torch.cat([shoeprint_latent_k, shoeprint_latent_k], dim=0))
    feats = generator.extract(ext_latents, torch.cat([w1, w2], dim=1))
    path_loss, path_raw = ops.path(feats, cent_fin_diff_h, opt["path_loss_lambda"])

    total = gan_loss + identity_loss + reconstruction_loss + kl_loss + path_loss + style_loss
    return total, gan_raw, rec_raw, idt_raw, kl_raw, path_raw, style_raw


def backward_unit(loss):
    """loss.backward() with the loss-seed fast path (each term's lambda is already folded in)."""
    ops.UNIT_LOSS_GRADS = True
    ops.DIRECT_WEIGHT_GRADS = True  # wgrad kernels accumulate straight into the gradient arenas
    try:
        loss.backward()
    finally:
        ops.UNIT_LOSS_GRADS = False
        ops.DIRECT_WEIGHT_GRADS = False


def discriminator_step(
    config,
    device: torch.device,
    discriminator: Discriminator,
    generator: Generator,
    mapping_network: MappingNetwork,
    discriminator_optimiser,
    shoeprint_iter: Iterator[torch.Tensor],
    shoemark_iter: Iterator[torch.Tensor],
    image_buffer: ImageBuffer,
    ada,
    ada_p: ADAp,
):
    """Take a step with the discriminator and return (loss, (real confidence, fake confidence))."""
    discriminator_optimiser.zero_grad()
    batch = config["training"]["batch_size"]

    shoeprint_images = next(shoeprint_iter).to(device)
    w = mapping_network.get_single_w(
        batch_size=batch, n_gen_blocks=generator.n_style_blocks, device=device, domain_variable=1
    )
    with torch.no_grad():
        generated_shoemarks = generator(shoeprint_images, w)
    buffered_shoemarks = image_buffer(generated_shoemarks)
    augmented_fake = ada(buffered_shoemarks)
    real_shoemarks = next(shoemark_iter).to(device)
    augmented_real = ada(real_shoemarks)

    disc_loss, sign_real, sign_fake = discriminator_losses(discriminator, augmented_fake, augmented_real,
                                                           r1_gamma(config))

    ada_p.update_p(sign_real.reshape(()))

    backward_unit(disc_loss)
    discriminator_optimiser.step()

    loss, s_real, s_fake = _floats(disc_loss, sign_real, sign_fake)
    return loss, (s_real, s_fake)


def generator_step(
    config,
    device: torch.device,
    generator: Generator,
    discriminator: Discriminator,
    mapping_network: MappingNetwork,
    style_extractor: StyleExtractor,
    generator_optimiser,
    mapping_network_optimiser,
    style_extractor_optimiser,
    shoeprint_iter: Iterator[torch.Tensor],
    shoemark_iter: Iterator[torch.Tensor],
    ada,
    *,
    cent_fin_diff_h: torch.Tensor | None = None,
):
    """Take a step with the generator and return (total, (gan, rec, idt, kl, path, style)).

    `cent_fin_diff_h` (optional, not in the reference signature) injects the finite-difference
    step the reference draws from the device generator (training.py:216-223) so seeded parity
    tests can replay it: a tensor, or a callable theta -> h invoked at the reference's position
    in the draw order (a CPU run of the reference draws h from the HOST generator there)."""
    generator_optimiser.zero_grad()
    mapping_network_optimiser.zero_grad()
    style_extractor_optimiser.zero_grad()
    opt = config["optimisation"]
    batch = config["training"]["batch_size"]
    nb = generator.n_style_blocks

    real_shoeprint_images = next(shoeprint_iter).to(device)
    real_shoemark_images = next(shoemark_iter).to(device)

    # --- styles, in the reference's host-RNG draw order ----------------------
    reconstruct_w = mapping_network.get_single_w(
        batch_size=batch, n_gen_blocks=nb, device=device, domain_variable=0
    )
    style_batch = batch * styles_per_input(config)
    translation_w = mapping_network.get_single_w(
        batch_size=style_batch, n_gen_blocks=nb, device=device, domain_variable=1
    )
    latent_noise = None
    if config["architecture"]["add_latent_noise"]:  # device draw, before h (training.py:166 vs :216)
        b2, c, lh, lw = generator.latent_shape((2 * batch, *real_shoeprint_images.shape[1:]))
        latent_noise = torch.randn((b2, c, lh, lw), device=device)  # NCHW element order, like randn_like
    theta_host = torch.rand(style_batch)
    theta = theta_host.to(device)
    if cent_fin_diff_h is None:
        lo, hi = opt["path_loss_jacobian_granularity"]
        cent_fin_diff_h = torch.ones_like(theta).uniform_(lo, hi)
    else:
        if callable(cent_fin_diff_h):
            cent_fin_diff_h = cent_fin_diff_h(theta_host)
        cent_fin_diff_h = cent_fin_diff_h.to(device=device, dtype=torch.float32)
    d1 = (theta + cent_fin_diff_h / 2).clamp(0, 1)
    d2 = (theta - cent_fin_diff_h / 2).clamp(0, 1)
    w1, w2 = mapping_network.get_two_w(
        batch_size=style_batch, n_gen_blocks=nb, device=device, domain_variables=(d1, d2)
    )

    losses = generator_losses(config, generator, discriminator, style_extractor,
                              real_shoeprint_images, real_shoemark_images, reconstruct_w,
                              translation_w, w1, w2, cent_fin_diff_h, ada, latent_noise)
    backward_unit(losses[0])
    generator_optimiser.step()
    mapping_network_optimiser.step()
    style_extractor_optimiser.step()

    vals = _floats(*losses)
    return vals[0], tuple(vals[1:])
