"""Image-folder input for train.py when `[training] synthetic_data = false`: the tensor
contract of the reference's ShoeDataset (src/data/datasets.py:13-50, train.py:120-169) —
fp32 [B,C,H,W] in [-1,1], random horizontal flip, shuffled, drop_last.  Host-side I/O only."""

from __future__ import annotations

from pathlib import Path

import torch


def image_folder_loader(config, key: str, generator: torch.Generator, rank: int = 0, world: int = 1):
    """`generator`: ONE torch.Generator shared by every loader of a run, like the reference's
    `dataloader_g` (train.py:56-58): successive epochs of the two folders then draw different
    permutations.  Every data-parallel rank must pass an identically seeded generator: each epoch
    all ranks draw the same permutation and flip mask and rank r takes batches r, r+world, ..."""
    from PIL import Image  # optional dependency, only for real-image training

    size = tuple(config["data"]["image_size"])
    path = Path(config["data"][key]).expanduser() / "train"
    files = sorted(list(path.rglob("*.jpg")) + list(path.rglob("*.png")))
    if not files:
        raise FileNotFoundError(path)
    images = []
    for f in files:
        img = Image.open(f).resize((size[1], size[0]), Image.BILINEAR)
        t = torch.frombuffer(bytearray(img.tobytes()), dtype=torch.uint8).float() / 255.0
        t = t.reshape(size[0], size[1], -1).permute(2, 0, 1)[: config["data"]["image_channels"]]
        images.append((t - 0.5) / 0.5)
    data = torch.stack(images)
    batch = config["training"]["batch_size"]
    gen = generator
    if len(data) < batch * world:
        raise ValueError(f"{path}: {len(data)} images cannot fill {world} batches of {batch}")

    class _Loader:
        def __iter__(self):
            perm = torch.randperm(len(data), generator=gen)
            flips = torch.rand(len(data), generator=gen) < 0.5
            n_batches = len(perm) // batch // world * world  # drop_last, equal work per rank
            for b in range(rank, n_batches, world):
                idx = perm[b * batch : (b + 1) * batch]
                x = data[idx]
                x = torch.where(flips[idx][:, None, None, None], x.flip(-1), x)
                yield x.pin_memory()

    return _Loader()
