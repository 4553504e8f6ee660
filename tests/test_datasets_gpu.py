"""The device-resident real-data path (datasets.DeviceImages + otm_gather_batch) against the
reference's transform chain ToTensor -> Normalize(0.5, 0.5) (+ horizontal flip), reference
train.py:120-126 / datasets.py:48-50: exact in fp32 (same two operations per pixel)."""

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("shape", [(37, 1, 64, 48), (10, 3, 17, 13), (9, 1, 32, 30)])
def test_gather_batch_matches_transform_chain(shape):
    from one_to_many_gan_b200 import kernels as K

    g = torch.Generator().manual_seed(1)
    data = torch.randint(0, 256, shape, generator=g, dtype=torch.uint8)
    idx = torch.randint(0, shape[0], (8,), generator=g)
    flip = (torch.rand(8, generator=g) < 0.5)
    out = K.gather_batch(data.cuda(), idx.cuda(), flip.to(torch.uint8).cuda())
    ref = (data[idx].float() / 255.0 - 0.5) / 0.5
    ref = torch.where(flip[:, None, None, None], ref.flip(-1), ref)
    assert out.shape == ref.shape
    assert (out.cpu() - ref).abs().max().item() < 1e-6
    out2 = K.gather_batch(data.cuda(), idx.cuda(), None)
    assert (out2.cpu() - (data[idx].float() / 255.0 - 0.5) / 0.5).abs().max().item() < 1e-6


def test_device_images_epochs_shard_by_rank():
    from one_to_many_gan_b200.datasets import DeviceImages

    n, B = 23, 4
    data = (torch.arange(n, dtype=torch.uint8)[:, None, None, None] * 10).expand(n, 1, 8, 8).contiguous()
    seen = []
    for rank in range(2):
        it = DeviceImages(data, B, "cuda", torch.Generator().manual_seed(5), rank=rank, world=2, flip=False)
        ids = []
        for _ in range(2):  # one epoch = 23 // 4 // 2 * 2 = 4 batches -> 2 per rank
            x = next(it)
            assert x.shape == (B, 1, 8, 8)
            ids += [round((v.item() + 1) * 255 / 2 / 10) for v in x[:, 0, 0, 0]]
        seen.append(ids)
        nxt = next(it)  # a new epoch starts transparently
        assert nxt.shape == (B, 1, 8, 8)
    assert len(set(seen[0]) | set(seen[1])) == 4 * B  # disjoint shards of one permutation
    assert not (set(seen[0]) & set(seen[1]))
