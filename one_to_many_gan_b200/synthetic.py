"""On-device synthetic batches (replaces src/data/datasets.py + the DataLoaders of reference
train.py:120-169 on the benchmark path): U(-1,1) fp32 [B,C,H,W] — the value range
`Normalize(0.5, 0.5)` produces — from a counter-based Philox4x32-10 stream keyed by
(seed, rank, stream id), advanced by one batch per `next()`."""

from __future__ import annotations

import torch

from . import kernels as K


class SyntheticImages:
    def __init__(self, batch_size: int, channels: int, image_size, device, *, seed: int = 42,
                 rank: int = 0, stream_id: int = 0):
        self.shape = (batch_size, channels, int(image_size[0]), int(image_size[1]))
        self.device = torch.device(device)
        self.seed = seed
        self.key = rank * 16 + stream_id
        self.numel = batch_size * channels * self.shape[2] * self.shape[3]
        self.step = 0

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        out = torch.empty(self.shape, dtype=torch.float32, device=self.device)
        offset = self.step * ((self.numel + 3) // 4 * 4)
        K.synth_uniform(out, self.seed, self.key, offset)
        self.step += 1
        return out
