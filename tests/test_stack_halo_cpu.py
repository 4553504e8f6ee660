"""ops.stack_halo (host logic, pure tensor plumbing): batch ranges of a halo-carrying tensor
stacked as copies of the padded buffer == torch.cat of the interior views, forward and backward
(the batched decode / path-extraction latents of generator_step, reference training.py:170-243)."""

import torch


def test_stack_halo_matches_cat_forward_and_backward():
    from one_to_many_gan_b200 import kernels as K
    from one_to_many_gan_b200 import ops

    torch.manual_seed(0)
    B = 2
    src = K.alloc(2 * B, 8, 5, 6, torch.float32, "cpu", 1)
    K.padded_view(src, 1).normal_()
    src.requires_grad_(True)
    ops.with_halo(src, 1)
    for blocks in ([(0, 2 * B), (0, B)], [(0, B), (0, B)], [(B, B)]):
        out = ops.stack_halo(src, blocks)
        ref = torch.cat([src[a : a + k] for a, k in blocks], 0)
        assert ops.halo_of(out) == 1
        assert torch.equal(out, ref)
        k0 = blocks[0][1]
        a0 = blocks[0][0]
        assert torch.equal(K.padded_view(out, 1)[:k0], K.padded_view(src, 1)[a0 : a0 + k0])  # halo too
        g = torch.randn_like(ref)
        (g1,) = torch.autograd.grad(out, src, g)
        (g2,) = torch.autograd.grad(ref, src, g)
        assert torch.allclose(g1, g2)
