"""Evaluation forward passes (one_to_many_gan_b200/inference.py) against the oracle with the
reference's host-RNG draw: `val_checkpoint`'s generator forward with un-mixed styles
(reference evaluation.py:48-57) and the one-input -> K-outputs decode of `image_checkpoint`
(:141-177), eager and as replayed CUDA graphs; fp32 parity mode 1e-4, bf16 5e-2 end to end."""

import pytest
import torch

from oracle import reference_port as rp
from tests.test_modules_gpu import build, images, relerr

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode,tol", [("fp32", 1e-4), ("bf16", 5e-2)])
@pytest.mark.parametrize("use_graph", [False, True])
def test_sampler_matches_oracle(mode, tol, use_graph):
    from one_to_many_gan_b200.inference import Sampler

    dt = torch.float32 if mode == "fp32" else torch.bfloat16
    arch, P, D, G, M, S = build("down1", dt)
    P64 = {n: {k: v.double() for k, v in p.items()} for n, p in P.items()}
    smp = Sampler(G, M, S, device="cuda", use_graph=use_graph)
    B, K = 4, 3
    for rep in range(2):  # the second call replays the captured graph on new inputs
        x = images((B, 1, *arch.image_size), 50 + rep)
        xm = images((B, 1, *arch.image_size), 60 + rep)
        torch.manual_seed(7 + rep)
        y = smp.translate(x.cuda()).clone()
        torch.manual_seed(7 + rep)
        with torch.no_grad():
            w = rp.get_single_w(P64["M"], B, arch.n_style_blocks, arch, 1, mix_styles=False)
            y_ref = rp.generator_forward(P64["G"], x.double(), w, arch)
        assert relerr(y, y_ref) < tol, ("translate", rep)

        torch.manual_seed(17 + rep)
        grid = smp.one_to_many(x.cuda(), K).clone()
        torch.manual_seed(17 + rep)
        with torch.no_grad():
            wk = rp.get_single_w(P64["M"], K, arch.n_style_blocks, arch, 1, mix_styles=False)
            lat = rp.generator_encode(P64["G"], x.double(), arch)
            ref = torch.stack([rp.generator_decode(P64["G"], lat[c : c + 1].expand(K, -1, -1, -1), wk, arch)
                               for c in range(B)])
        assert grid.shape == ref.shape and relerr(grid, ref) < tol, ("one_to_many", rep)

        rec_p, trans, rec_m = [t.clone() for t in smp.decoding_grid(x.cuda(), xm.cuda())]
        with torch.no_grad():
            latp = rp.generator_encode(P64["G"], x.double(), arch)
            latm = rp.generator_encode(P64["G"], xm.double(), arch)
            ws = rp.style_extractor_forward(P64["S"], xm.double())
            ws = ws.expand(arch.n_style_blocks, *ws.shape)
            want = (rp.generator_decode(P64["G"], latp, torch.zeros_like(ws), arch),
                    rp.generator_decode(P64["G"], latp, ws, arch),
                    rp.generator_decode(P64["G"], latm, ws, arch))
        for name, a, b in zip(("rec prints", "translated", "rec marks"), (rec_p, trans, rec_m), want):
            assert relerr(a, b) < tol, (name, rep)
