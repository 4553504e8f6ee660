"""Host logic of the engine's image pool against the reference ImageBuffer semantics
(reference src/core/training.py:22-65), with integers standing in for images."""

import random

import pytest


def _reference(pool_size, batch, iters, seed):
    random.seed(seed)
    images, out = [], []
    for it in range(iters):
        ret = []
        for j in range(batch):
            new = (it, j)
            if len(images) < pool_size:
                images.append(new)
                ret.append(new)
            else:
                if random.uniform(0, 1) > 0.5:
                    k = random.randint(0, pool_size - 1)
                    ret.append(images[k])
                    images[k] = new
                else:
                    ret.append(new)
        out.append(ret)
    return out, images


@pytest.mark.parametrize("pool_size,batch", [(5, 2), (3, 4), (100, 4), (1, 3), (7, 7)])
def test_pool_indexer_matches_image_buffer(pool_size, batch):
    from one_to_many_gan_b200.engine import PoolIndexer

    iters = 40
    want, want_pool = _reference(pool_size, batch, iters, 11)
    random.seed(11)
    idx = PoolIndexer(pool_size)
    pool = [None] * (pool_size + 1)
    for it in range(iters):
        new = [(it, j) for j in range(batch)]
        src, dst, sto = idx.decide(batch)
        assert len(src) == len(dst) == len(sto) == batch
        sources = pool + new
        assert [sources[s] for s in src] == want[it], it
        real = [d for d in dst if d != pool_size]
        assert len(real) == len(set(real))  # no duplicate scatter targets except the scratch slot
        for d, j in zip(dst, sto):
            pool[d] = new[j]
    assert pool[:pool_size][: len(want_pool)] == want_pool


def test_pool_indexer_rejects_empty_pool():
    from one_to_many_gan_b200.engine import PoolIndexer

    with pytest.raises(ValueError):
        PoolIndexer(0)
