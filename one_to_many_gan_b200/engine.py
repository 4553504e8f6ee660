"""One full training iteration (D step + G step) as a replayable CUDA graph.

The reference drives its iteration from Python with ten `.cpu().item()` syncs, host RNG
draws interleaved with device work and a Python image pool (train.py:204-251,
training.py:22-65).  At 128x128 / batch 32 an iteration is ~2,800 kernel launches, so the host
would bound the GPU.  This engine keeps the SAME mathematics and the SAME host-RNG draw
order, but separates the iteration into

  host part   : draw the style noise / mixing coin / crossover / theta (host generator) and the
                image-pool decisions (python `random`) -> one pinned staging buffer,
  device part : everything else, static shapes, no host sync -> captured once with
                torch.cuda.graph and replayed.

Differences from the eager step functions that make the device part static (numerically
identical): style mixing is a `where(block < crossover, s1, s2)` instead of expand+cat
(builder.py:115-132); the image pool is a device tensor indexed by host-decided slots."""

from __future__ import annotations

import gc
import os
import random

import torch

from . import ops, training
from .optim import all_reduce_buckets


class PoolIndexer:
    """Host half of the image history pool (reference ImageBuffer, training.py:22-65): consumes
    python `random` exactly like the reference and resolves every per-image decision of a batch
    to gather/scatter indices for the device half.

    decide(B) -> (src[B], dst[B], sto[B]):
      returned image j = sources[src[j]], sources = [pool slots 0..P-1, scratch slot P, new 0..B-1]
      after the batch   pool[dst[i]] = new[sto[i]]  (padding writes go to the scratch slot P)."""

    def __init__(self, pool_size: int):
        if pool_size < 1:
            raise ValueError
        self.size = pool_size
        self.count = 0

    def decide(self, batch: int):
        P = self.size
        new0 = P + 1
        content: dict[int, int] = {}  # slot -> index of the new image written this batch
        src = []
        for j in range(batch):
            if self.count < P:
                content[self.count] = j
                self.count += 1
                src.append(new0 + j)
            elif random.uniform(0, 1) > 0.5:
                k = random.randint(0, P - 1)
                src.append(new0 + content[k] if k in content else k)
                content[k] = j
            else:
                src.append(new0 + j)
        dst = list(content.keys())  # last writer wins
        sto = list(content.values())
        while len(dst) < batch:
            dst.append(P)
            sto.append(0)
        return src, dst, sto


class TrainIteration:
    LOSS_NAMES = ("disc", "sign_real", "sign_fake", "total_gen", "gan", "rec", "idt", "kl", "path",
                  "style")

    def __init__(self, config, device, discriminator, generator, mapping_network, style_extractor,
                 opt_d, opt_g, opt_m, opt_s, *, use_graph: bool = True, warmup: int = 2):
        self.cfg = config
        self.dev = torch.device(device)
        self.D, self.G, self.M, self.S = discriminator, generator, mapping_network, style_extractor
        self.oD, self.oG, self.oM, self.oS = opt_d, opt_g, opt_m, opt_s
        self.B = config["training"]["batch_size"]
        self.nb = generator.n_style_blocks
        self.wd = mapping_network.d_latent
        self.mix_p = mapping_network.style_mixing_prob
        size = config["data"]["image_size"]
        ch = config["data"].get("image_channels", 1)
        B, wd = self.B, self.wd
        # styles per input (BASELINE config 4): the G step's sampled-style passes run at B*K
        self.K = training.styles_per_input(config)
        BK = B * self.K
        # static inputs: [d_prints, d_marks, g_prints, g_marks]
        self.x = torch.zeros(4, B, ch, size[0], size[1], device=self.dev)
        # host-drawn randomness: 3 style draws x (z1, z2) -- slot 0 (D step) at batch B, slots
        # 1-2 (G step) at batch B*K -- then theta and the optional injected h (B*K each)
        self.style_batch = (B, BK, BK)
        self.style_base = (0, 2 * B * wd, 2 * B * wd + 2 * BK * wd)
        self.theta_base = 2 * B * wd + 4 * BK * wd
        self.n_rng = self.theta_base + 2 * BK
        # pinned staging is double-buffered: with sync_losses=False the GPU may still be reading
        # iteration i's buffers (asynchronous H2D) while the host draws iteration i+1
        self._rng_hosts = [torch.zeros(self.n_rng, dtype=torch.float32).pin_memory() for _ in range(2)]
        self._staged = [None, None]  # event recorded after each buffer's H2D copies
        self.rng_host = self._rng_hosts[0]
        self.rng_dev = torch.zeros(self.n_rng, dtype=torch.float32, device=self.dev)
        # integer controls: 3 crossovers, B pool sources, B pool destinations, B stored images
        self.n_idx = 3 + 3 * B
        self._idx_hosts = [torch.zeros(self.n_idx, dtype=torch.int64).pin_memory() for _ in range(2)]
        self.idx_host = self._idx_hosts[0]
        self.idx_dev = torch.zeros(self.n_idx, dtype=torch.int64, device=self.dev)
        self.pool_size = config["training"].get("image_buffer_size", 100)
        if self.pool_size < 1:
            raise ValueError
        self.pool = torch.zeros(self.pool_size + 1, ch, size[0], size[1], device=self.dev)
        self.pool_index = PoolIndexer(self.pool_size)
        self.losses = torch.zeros(len(self.LOSS_NAMES), dtype=torch.float32, device=self.dev)
        self.losses_host = torch.zeros(len(self.LOSS_NAMES), dtype=torch.float32).pin_memory()
        self.inject_h = False
        self.use_graph = use_graph
        # "full": the whole iteration (NCCL included) is ONE graph and the gradient exchange
        # overlaps compute; "segmented": three graphs with blocking NCCL between them
        self.mode = "segmented" if os.environ.get("OTM_DDP_CAPTURE", "1") == "0" else "full"
        self.g_dec_start = opt_g.offset_of(next(generator.decoder.parameters()))
        self.on_decoder_grads_final = None  # test hook
        self.graph = None
        self._warm_left = warmup
        self.packs = ops.PackRecorder()  # shared weight packs of a phase in one launch
        self.iterations = 0

    # ---------------------------------------------------------------- host part
    def _draw_style(self, slot: int):
        """builder.py:106-132 draw order: rand(()), [randint, randn, randn] | [randn]."""
        B, wd = self.style_batch[slot], self.wd
        base = self.style_base[slot]
        if torch.rand(()).lt(self.mix_p):
            cross = int(torch.randint(0, self.nb, ()))
            z1 = torch.randn(B, wd)
            z2 = torch.randn(B, wd)
        else:
            cross = self.nb
            z1 = torch.randn(B, wd)
            z2 = z1
        self.rng_host[base : base + B * wd] = z1.reshape(-1)
        self.rng_host[base + B * wd : base + 2 * B * wd] = z2.reshape(-1)
        self.idx_host[slot] = cross

    def _pool_decisions(self):
        src, dst, sto = self.pool_index.decide(self.B)
        o, B = 3, self.B
        self.idx_host[o : o + B] = torch.tensor(src)
        self.idx_host[o + B : o + 2 * B] = torch.tensor(dst)
        self.idx_host[o + 2 * B : o + 3 * B] = torch.tensor(sto)

    def _sample_host(self, h):
        BK = self.B * self.K
        self._draw_style(0)              # D step: get_single_w(d=1)
        self._pool_decisions()           # ImageBuffer
        self._draw_style(1)              # G step: translation w
        tbase = self.theta_base
        self.rng_host[tbase : tbase + BK] = torch.rand(BK)  # theta
        if isinstance(h, str):  # "host": drawn HERE from the host generator, exactly where a CPU run
            lo, hi = self.cfg["optimisation"]["path_loss_jacobian_granularity"]  # of the reference
            h = torch.ones(BK).uniform_(lo, hi)                                  # draws it (:216-223)
        if h is not None:
            self.rng_host[tbase + BK : tbase + 2 * BK] = h.float().cpu()
        self._draw_style(2)              # G step: get_two_w

    # ---------------------------------------------------------------- device part
    def _style(self, slot: int, d=(None, None), n_out: int = 1):
        """get_single_w(d=1) / get_two_w (reference builder.py:51-132) as ONE launch: mapping
        network on both host-drawn z, style mixing at the host-drawn crossover, lerp(0, s, d)."""
        B, wd, nb = self.style_batch[slot], self.wd, self.nb
        base = self.style_base[slot]
        z1 = self.rng_dev[base : base + B * wd].view(B, wd)
        z2 = self.rng_dev[base + B * wd : base + 2 * B * wd].view(B, wd)
        return ops.mapping(self.M.linears(), z1, z2, self.idx_dev[slot : slot + 1], n_blocks=nb, d=d,
                           n_out=n_out)

    # The device part of one iteration, in stream order (data-parallel exchanges marked *):
    #   D forward / backward
    #   * all-reduce(D grads) launched asynchronously ...
    #   G step forward up to the generated images (encode, style extractor, 3B decode, L1 terms)
    #   ... waited for here: Adam(D) runs right before the G step first uses the discriminator
    #   rest of the G step forward, then backward; when autograd reaches the encoder (hook on
    #   the latents) every decoder / style-extractor gradient is final:
    #   * all-reduce(G decoder slice, S) launched asynchronously under the encoder backward
    #   * all-reduce(G encoder slice, M) after backward; wait; Adam(G), Adam(M), Adam(S)
    # NCCL collectives are capture-safe, so the WHOLE iteration is one CUDA graph (mode "full");
    # mode "segmented" keeps NCCL outside three captured segments (OTM_DDP_CAPTURE=0).
    def _d_step_and_zero(self):
        B = self.B
        ops.invalidate_packs()
        self.packs.phase("iteration")  # every shared pack of G / D / S in one launch
        self.oD.zero_grad()
        with torch.no_grad():
            (w,) = self._style(0)
            generated = self.G(self.x[0], w)
            o = 3
            sources = torch.cat([self.pool, generated], dim=0)
            fake = sources.index_select(0, self.idx_dev[o : o + B])
            self.pool.index_copy_(0, self.idx_dev[o + B : o + 2 * B],
                                  generated.index_select(0, self.idx_dev[o + 2 * B : o + 3 * B]))
        disc_loss, sign_real, sign_fake = training.discriminator_losses(
            self.D, fake, self.x[1], training.r1_gamma(self.cfg))
        training.backward_unit(disc_loss)
        self.losses[0:3].copy_(torch.cat([v.detach().reshape(1).float()
                                          for v in (disc_loss, sign_real, sign_fake)]))

    def _update_d(self):
        """Adam(D): the reference runs it at the end of the D step (training.py:123); nothing
        before the G step's discriminator forward reads D, so it is deferred to there and D's
        all-reduce hides behind the G step's encode / decode."""
        self.oD.wait_all_reduce()
        self.oD.step(reduced=True)
        self.packs.phase("after_adam_d")  # D's packs went stale: rebuild them in one launch

    def _decoder_grads_final(self):
        """Autograd is about to run the encoder backward: decoder, to_style and style-extractor
        gradients are complete (verified by test_engine_gpu.test_bucket_gradients_are_final)."""
        all_reduce_buckets([(self.oG, self.g_dec_start, None), (self.oS, 0, None)])
        if self.on_decoder_grads_final is not None:
            self.on_decoder_grads_final()

    def _g_step(self, update_d, overlap: bool = True):
        cfg, B = self.cfg, self.B
        opt = cfg["optimisation"]
        self.oG.zero_grad()
        self.oM.zero_grad()
        self.oS.zero_grad()
        zero = self.M.shoeprint_style_vector
        reconstruct_w = zero.expand(self.nb, B, self.wd)
        (translation_w,) = self._style(1)
        BK = B * self.K
        tbase = self.theta_base
        theta = self.rng_dev[tbase : tbase + BK]
        latent_noise = None
        if cfg["architecture"]["add_latent_noise"]:  # device draw BEFORE h (training.py:166 vs :216)
            latent_noise = torch.randn(self.G.latent_shape((2 * B, *self.x.shape[2:])), device=self.dev)
        if self.inject_h:
            h = self.rng_dev[tbase + BK : tbase + 2 * BK]
        else:
            lo, hi = opt["path_loss_jacobian_granularity"]
            h = torch.empty(BK, device=self.dev).uniform_(lo, hi)
        d1 = (theta + h / 2).clamp(0, 1)
        d2 = (theta - h / 2).clamp(0, 1)
        w1, w2 = self._style(2, d=(d1, d2), n_out=2)  # builder.py:66-71
        losses = training.generator_losses(cfg, self.G, self.D, self.S, self.x[2], self.x[3],
                                           reconstruct_w, translation_w, w1, w2, h,
                                           latent_noise=latent_noise, before_discriminator=update_d,
                                           on_latent_grad=self._decoder_grads_final if overlap else None)
        training.backward_unit(losses[0])
        self.losses[3:].copy_(torch.cat([v.detach().reshape(1).float() for v in losses]))

    def _update_gms(self):
        # the buckets not launched yet (G encoder slice, M; everything without overlap) as one
        # coalesced collective, then wait for all of them
        opts = (self.oG, self.oM, self.oS)
        all_reduce_buckets([(o, lo, hi) for o in opts for lo, hi in o._missing()])
        for o in opts:
            o.wait_all_reduce()
        self.oG.step(reduced=True)
        self.oM.step(reduced=True)
        self.oS.step(reduced=True)
        ops.invalidate_packs()
        self.packs.phase(None)

    def _iteration(self):
        self._d_step_and_zero()
        self.oD.all_reduce_async()
        self._g_step(self._update_d)
        self._update_gms()

    # mode "segmented": NCCL between three captured segments, no overlap
    def _segment_a(self):
        self._d_step_and_zero()

    def _segment_b(self):
        ops.invalidate_packs()
        self.oD.step(reduced=True)
        self.packs.phase("segment_b")
        self._g_step(None, overlap=False)

    def _segment_c(self):
        self.oG.step(reduced=True)
        self.oM.step(reduced=True)
        self.oS.step(reduced=True)
        ops.invalidate_packs()
        self.packs.phase(None)

    def _exchange(self, opts):
        for o in opts:
            o.all_reduce_async()
        for o in opts:
            o.wait_all_reduce()

    def close(self):
        """Drop the captured graphs.  MUST run before `dist.destroy_process_group()`: tearing a NCCL
        communicator down while a CUDA graph that captured its collectives is alive hangs."""
        torch.cuda.synchronize(self.dev)
        self.graph = None
        self._warm_left = 0
        gc.collect()
        torch.cuda.synchronize(self.dev)

    # ---------------------------------------------------------------- image pool I/O
    def pool_images(self):
        """The pool as the reference's `ImageBuffer.images` list of [1,C,H,W] tensors."""
        return [self.pool[i : i + 1].clone() for i in range(self.pool_index.count)]

    def load_pool(self, images):
        """Restore the pool from a checkpoint (a list like pool_images() returns)."""
        if len(images) > self.pool_size:
            raise ValueError(f"checkpoint holds {len(images)} pool images, pool size is {self.pool_size}")
        for i, img in enumerate(images):
            self.pool[i].copy_(img.reshape(self.pool.shape[1:]))
        self.pool_index.count = len(images)

    # ---------------------------------------------------------------- driver
    def _capture(self, parts):
        torch.cuda.synchronize()
        graphs = []
        pool = torch.cuda.graph_pool_handle()
        # no cyclic GC while a stream is capturing: collecting some OTHER dead graph or tensor
        # would cudaFree inside the capture and invalidate it
        gc_was_on = gc.isenabled()
        gc.collect()
        gc.disable()
        try:
            for part in parts:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g, pool=pool):
                    part()
                graphs.append(g)
        finally:
            if gc_was_on:
                gc.enable()
        self.graph = graphs

    def load_inputs(self, d_prints, d_marks, g_prints, g_marks):
        """Copy the four batches of this iteration into the static input buffers (device
        tensors: device-to-device; pinned host tensors: asynchronous H2D)."""
        for i, t in enumerate((d_prints, d_marks, g_prints, g_marks)):
            self.x[i].copy_(t, non_blocking=True)

    def run(self, *, h: torch.Tensor | str | None = None, sync_losses: bool = True):
        """Run one iteration on the batches last given to load_inputs().  Returns the ten
        logged scalars as floats (one device->host copy) or None if sync_losses is False.
        h: the finite-difference steps of the path-length loss -- None: drawn on the device
        generator like the reference's GPU run; "host": drawn on the host generator at the
        reference's position in the draw order (its CPU run; makes a run reproducible from the
        host RNG state alone); a tensor: injected (parity tests)."""
        if (h is not None) != self.inject_h:
            if self.graph is not None:
                raise RuntimeError("cannot switch h injection after the graph was captured")
            self.inject_h = h is not None
        slot = self.iterations & 1
        if self._staged[slot] is not None:
            self._staged[slot].synchronize()  # the copies issued two iterations ago have run
        self.rng_host, self.idx_host = self._rng_hosts[slot], self._idx_hosts[slot]
        self._sample_host(h)
        self.rng_dev.copy_(self.rng_host, non_blocking=True)
        self.idx_dev.copy_(self.idx_host, non_blocking=True)
        if self._staged[slot] is None:
            self._staged[slot] = torch.cuda.Event()
        self._staged[slot].record()
        if self.mode == "full":
            if not self.use_graph or self._warm_left > 0:
                # eager (also the warm-up: lazy inits, allocator pools, NCCL communicators)
                self._warm_left = max(0, self._warm_left - 1)
                self._iteration()
            else:
                if self.graph is None:
                    self._capture([self._iteration])
                self.graph[0].replay()
        else:
            segs = (self._segment_a, self._segment_b, self._segment_c)
            exch = ((self.oD,), (self.oG, self.oM, self.oS), ())
            if not self.use_graph or self._warm_left > 0:
                self._warm_left = max(0, self._warm_left - 1)
                for seg, ex in zip(segs, exch):
                    seg()
                    self._exchange(ex)
            else:
                if self.graph is None:
                    self._capture(segs)
                for g, ex in zip(self.graph, exch):
                    g.replay()
                    self._exchange(ex)
        ops.invalidate_packs()  # replays update the weights behind Python's back
        self.iterations += 1
        if not sync_losses:
            return None
        self.losses_host.copy_(self.losses, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        v = self.losses_host.tolist()
        return dict(zip(self.LOSS_NAMES, v))
