"""Real-image input for train.py when `[training] synthetic_data = false`.

The reference's `ShoeDataset` (src/data/datasets.py:13-50) decodes every image once, keeps the
transformed fp32 tensors in host RAM and feeds the step through three 8-worker DataLoaders
(train.py:120-169).  Here the resized folder is decoded once to uint8 and lives in HBM
(`DeviceImages`): a batch is one gather kernel that applies ToTensor + Normalize(0.5, 0.5) and
the random horizontal flip (`otm_gather_batch`), so the training loop never touches host image
memory again -- no worker processes, no per-step H2D image copies.  Shuffling, the flip coin and
drop_last are host decisions drawn from ONE generator shared by both loaders like the
reference's `dataloader_g` (train.py:56-58); under `torchrun` every rank draws the same
permutation and takes batches rank, rank + world, ..."""

from __future__ import annotations

from pathlib import Path

import torch

from . import kernels as K


def load_folder_uint8(path, image_size, channels: int) -> torch.Tensor:
    """uint8 [N, C, H, W] of every .jpg/.png under `path`, resized like
    transforms.Resize(image_size) on the PIL image (reference train.py:120-126)."""
    from PIL import Image  # optional dependency, only for real-image training

    size = (int(image_size[0]), int(image_size[1]))
    path = Path(path)
    files = sorted(list(path.rglob("*.jpg")) + list(path.rglob("*.png")))
    if not files:
        raise FileNotFoundError(path)
    out = torch.empty((len(files), channels, *size), dtype=torch.uint8)
    for i, f in enumerate(files):
        img = Image.open(f)
        img = img.convert("L" if channels == 1 else "RGB").resize((size[1], size[0]), Image.BILINEAR)
        t = torch.frombuffer(bytearray(img.tobytes()), dtype=torch.uint8)
        out[i] = t.reshape(size[0], size[1], channels).permute(2, 0, 1)
    return out


class DeviceImages:
    """An epoch-cycling batch iterator over a uint8 image set resident in HBM."""

    def __init__(self, data_uint8: torch.Tensor, batch_size: int, device, generator: torch.Generator,
                 rank: int = 0, world: int = 1, flip: bool = True):
        if data_uint8.dtype != torch.uint8 or data_uint8.dim() != 4:
            raise ValueError("expected a uint8 [N,C,H,W] tensor")
        if len(data_uint8) < batch_size * world:
            raise ValueError(f"{len(data_uint8)} images cannot fill {world} batches of {batch_size}")
        self.data = data_uint8.contiguous().to(device)
        self.batch, self.gen, self.rank, self.world, self.flip = batch_size, generator, rank, world, flip
        self._plan = iter(())
        # two pinned staging slots: the previous batch's asynchronous H2D copy may still be in flight
        self._idx_host = [torch.zeros(batch_size, dtype=torch.int64).pin_memory() for _ in range(2)]
        self._flip_host = [torch.zeros(batch_size, dtype=torch.uint8).pin_memory() for _ in range(2)]
        self._events = [None, None]
        self._count = 0

    def _epoch(self):
        n = len(self.data)
        perm = torch.randperm(n, generator=self.gen)
        flips = (torch.rand(n, generator=self.gen) < 0.5) if self.flip else torch.zeros(n, dtype=torch.bool)
        n_batches = n // self.batch // self.world * self.world  # drop_last, equal work per rank
        for b in range(self.rank, n_batches, self.world):
            idx = perm[b * self.batch : (b + 1) * self.batch]
            yield idx, flips[idx]

    def __iter__(self):
        return self

    def __next__(self) -> torch.Tensor:
        try:
            idx, fl = next(self._plan)
        except StopIteration:
            self._plan = self._epoch()
            idx, fl = next(self._plan)
        slot = self._count & 1
        self._count += 1
        if self._events[slot] is not None:
            self._events[slot].synchronize()
        self._idx_host[slot].copy_(idx)
        self._flip_host[slot].copy_(fl.to(torch.uint8))
        idx_dev = self._idx_host[slot].to(self.data.device, non_blocking=True)
        flip_dev = self._flip_host[slot].to(self.data.device, non_blocking=True)
        if self._events[slot] is None:
            self._events[slot] = torch.cuda.Event()
        self._events[slot].record()
        return K.gather_batch(self.data, idx_dev, flip_dev)


def image_folder_loader(config, key: str, generator: torch.Generator, rank: int = 0, world: int = 1,
                        device=None):
    """`config["data"][key]/train` as a DeviceImages iterator."""
    data = load_folder_uint8(Path(config["data"][key]).expanduser() / "train", config["data"]["image_size"],
                             config["data"]["image_channels"])
    dev = device if device is not None else torch.device("cuda", torch.cuda.current_device())
    return DeviceImages(data, config["training"]["batch_size"], dev, generator, rank, world)
