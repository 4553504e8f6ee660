"""Evaluation-time forward passes of the reference on the training kernels.

The reference runs them every `checkpoint_interval` steps under `no_grad`
(src/core/evaluation.py): `val_checkpoint` pushes `n_evaluation_images` shoeprints through
`generator(shoeprints, w)` at `inference_batch_size` with un-mixed styles (:48-57);
`image_checkpoint` decodes every one of 8 latents under the SAME 8 sampled styles (the
one-input -> many-outputs grid, :141-177) and builds the reconstruction / style-transfer grid
(:187-211).  Here each of those is ONE replayable CUDA graph over static buffers: the host
draws z exactly as `MappingNetwork._get_style_vector(mix_styles=False)` does (one
`torch.randn(batch, w_dim)`, builder.py:129-130) and everything else is device work.  PNG
writing and FID/KID stay with the caller (host I/O, third-party clean-fid)."""

from __future__ import annotations

import torch

from . import ops


class Sampler:
    def __init__(self, generator, mapping_network, style_extractor=None, *, device, use_graph: bool = True):
        self.G, self.M, self.S = generator, mapping_network, style_extractor
        self.dev = torch.device(device)
        self.use_graph = use_graph
        self._plans: dict = {}

    # ------------------------------------------------------------------ plumbing
    def _plan(self, key, build):
        """build() -> (static input tensors, fn producing output tensors).  Runs fn once eagerly
        (lazy initialisation outside capture), then captures it."""
        if key in self._plans:
            return self._plans[key]
        inputs, fn = build()
        with torch.no_grad():
            ops.invalidate_packs()
            out = fn()
            graph = None
            if self.use_graph:
                torch.cuda.synchronize()
                graph = torch.cuda.CUDAGraph()
                ops.invalidate_packs()  # every weight pack the graph reads is built inside it
                with torch.cuda.graph(graph):
                    out = fn()
                ops.invalidate_packs()
        self._plans[key] = (inputs, fn, graph, out)
        return self._plans[key]

    def _run(self, plan):
        inputs, fn, graph, out = plan
        if graph is not None:
            graph.replay()
            return out
        with torch.no_grad():
            ops.invalidate_packs()
            return fn()

    def _draw_z(self, batch: int) -> torch.Tensor:
        return torch.randn(batch, self.M.d_latent)  # host generator, like builder.py:129

    def _w(self, z_dev, batch):
        (w,) = ops.mapping(self.M.linears(), z_dev, n_blocks=self.G.n_style_blocks)
        return w

    # ------------------------------------------------------------------ public passes
    def translate(self, shoeprints: torch.Tensor) -> torch.Tensor:
        """generator(shoeprints, get_single_w(mix_styles=False, domain_variable=1))
        (reference evaluation.py:48-57).  Returns fp32 [B,C,H,W] (a static buffer: copy it before
        the next call)."""
        B = shoeprints.shape[0]

        def build():
            x = torch.zeros(tuple(shoeprints.shape), dtype=torch.float32, device=self.dev)
            z = torch.zeros(B, self.M.d_latent, device=self.dev)
            return (x, z), lambda: self.G(x, self._w(z, B))

        plan = self._plan(("translate", tuple(shoeprints.shape)), build)
        x, z = plan[0]
        x.copy_(shoeprints, non_blocking=True)
        z.copy_(self._draw_z(B), non_blocking=True)
        return self._run(plan)

    def one_to_many(self, shoeprints: torch.Tensor, k: int = 8) -> torch.Tensor:
        """Every input decoded under the SAME k sampled styles (reference evaluation.py:141-177:
        `generator.decode(latents[col].expand(8, ...), w)` per column).  One encode of the n
        inputs, one decode at batch n*k.  Returns [n, k, C, H, W]."""
        n = shoeprints.shape[0]

        def build():
            x = torch.zeros(tuple(shoeprints.shape), dtype=torch.float32, device=self.dev)
            z = torch.zeros(k, self.M.d_latent, device=self.dev)

            def fn():
                lat = self.G.encode(x)
                w = self._w(z, k)                                   # [nb, k, wd]
                y = self.G.decode(lat.repeat_interleave(k, dim=0), w.repeat(1, n, 1))
                return y.view(n, k, *y.shape[1:])

            return (x, z), fn

        plan = self._plan(("one_to_many", tuple(shoeprints.shape), k), build)
        x, z = plan[0]
        x.copy_(shoeprints, non_blocking=True)
        z.copy_(self._draw_z(k), non_blocking=True)
        return self._run(plan)

    def decoding_grid(self, shoeprints: torch.Tensor, shoemarks: torch.Tensor):
        """(reconstructed shoeprints, shoeprints translated with the shoemarks' extracted styles,
        reconstructed shoemarks) -- reference evaluation.py:179-199; one 3n decode."""
        if self.S is None:
            raise ValueError("decoding_grid needs the style extractor")
        n = shoeprints.shape[0]

        def build():
            xp = torch.zeros(tuple(shoeprints.shape), dtype=torch.float32, device=self.dev)
            xm = torch.zeros(tuple(shoemarks.shape), dtype=torch.float32, device=self.dev)

            def fn():
                lat = self.G.encode(torch.cat([xp, xm], dim=0))
                lp, lm = lat[:n], lat[n:]
                ws = self.S(xm)
                ws = ws.expand(self.G.n_style_blocks, *ws.shape)
                w0 = torch.zeros_like(ws)
                y = self.G.decode(torch.cat([lp, lp, lm], dim=0), torch.cat([w0, ws, ws], dim=1))
                return y[:n], y[n : 2 * n], y[2 * n :]

            return (xp, xm), fn

        plan = self._plan(("grid", tuple(shoeprints.shape)), build)
        xp, xm = plan[0]
        xp.copy_(shoeprints, non_blocking=True)
        xm.copy_(shoemarks, non_blocking=True)
        return self._run(plan)
