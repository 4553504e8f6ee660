"""Pin the CPU oracle (oracle/reference_port.py) to reference-generated vectors.

The fixtures were produced by tests/golden/make_golden.py running the unmodified
reference (src/model/*, src/core/training.py) in the build container."""

import pytest
import torch

from oracle import reference_port as rp
from tests.golden_util import CASES, fingerprint, fp_err, load


def _arch(meta):
    return rp.Arch(
        image_size=tuple(meta["image_size"]),
        min_latent_resolution=meta["min_latent"],
        n_resnet_blocks=meta["n_res"],
    )


@pytest.mark.parametrize("case", CASES)
def test_init_matches_reference_draw_order(case):
    g = load(case)
    params = rp.init_all(_arch(g["meta"]), g["meta"]["seed"])
    for net in "DGMS":
        assert set(params[net]) == set(g["init_fp"][net]), net
        for k, v in params[net].items():
            assert fp_err(fingerprint(v), g["init_fp"][net][k]) < 1e-12, (net, k)


@pytest.mark.parametrize("case", CASES)
def test_forward_known_answers(case):
    g = load(case)
    meta = g["meta"]
    arch = _arch(meta)
    P = rp.init_all(arch, meta["seed"])
    gx = torch.Generator().manual_seed(g["forward"]["x_seed"])
    x = torch.rand(meta["batch"], 1, *meta["image_size"], generator=gx) * 2 - 1
    w = torch.rand(meta["n_style_blocks"], meta["batch"], arch.w_dim, generator=gx)
    assert arch.n_style_blocks == meta["n_style_blocks"]
    with torch.no_grad():
        z = rp.generator_encode(P["G"], x, arch)
        y = rp.generator_decode(P["G"], z, w, arch)
        feats = rp.generator_extract(P["G"], z, w, arch)
        assert fp_err(fingerprint(z), g["forward"]["latent_fp"]) < 1e-5
        assert fp_err(fingerprint(y), g["forward"]["g_out_fp"]) < 1e-5
        assert len(feats) == len(g["forward"]["feat_fp"])
        for f, want in zip(feats, g["forward"]["feat_fp"]):
            assert fp_err(fingerprint(f), want) < 1e-5
        torch.testing.assert_close(
            rp.discriminator_forward(P["D"], x), g["forward"]["d_out"], rtol=1e-4, atol=1e-5
        )
        torch.testing.assert_close(
            rp.style_extractor_forward(P["S"], x), g["forward"]["s_out"], rtol=1e-4, atol=1e-5
        )
        torch.testing.assert_close(
            rp.mapping_forward(P["M"], w[0], arch), g["forward"]["m_out"], rtol=1e-5, atol=1e-6
        )


def _batches(shape, seed):
    gen = torch.Generator().manual_seed(seed)
    while True:
        yield torch.rand(*shape, generator=gen) * 2 - 1


@pytest.mark.parametrize("case", CASES)
def test_training_steps_match_reference(case):
    g = load(case)
    meta = g["meta"]
    arch = _arch(meta)
    params = rp.init_all(arch, meta["seed"])  # also seeds torch + python random
    tr = rp.Trainer(arch, rp.Hyper(batch_size=meta["batch"]), params)
    shape = (meta["batch"], 1, *meta["image_size"])
    prints, marks = _batches(shape, meta["print_seed"]), _batches(shape, meta["mark_seed"])
    for it in range(meta["iters"]):
        d = tr.discriminator_step(next(prints), next(marks))
        if it == 0:
            for k, want in g["d_grad_fp"].items():
                assert fp_err(fingerprint(tr.last_grads["D"][k]), want) < 2e-3, k
        gl = tr.generator_step(next(prints), next(marks))
        got = torch.tensor([d[0], d[1][0], d[1][1], gl[0], *gl[1]], dtype=torch.float64)
        torch.testing.assert_close(got, g["losses"][it], rtol=2e-4, atol=1e-6)
        if it == 0:
            dead = _dead_bias_names()
            for net in "GMS":
                for k, want in g["g_grad_fp"][net].items():
                    if (net, k) in dead:
                        continue  # exact gradient is 0, both sides hold rounding noise
                    assert fp_err(fingerprint(tr.last_grads[net][k]), want) < 5e-3, (net, k)


def _dead_bias_names():
    """Biases cancelled by a following InstanceNorm."""
    dead = {("S", f"model.{i}.bias") for i in (3, 7, 11)}
    dead |= {("D", f"model.{i}.bias") for i in (3, 7, 11)}
    dead |= {("G", f"encoder.{i}.bias") for i in (1, 4, 8, 12)}
    return dead


def test_config1_twenty_iterations_match_reference():
    """BASELINE configs[0] literally (64x64, batch 4, 20 iterations): the port's fp32 losses
    against the reference's, iteration by iteration, inside the measured chaos envelope."""
    from tests.golden_util import LOSS_COLS, chaos_envelope

    g = load("d_64x64_config1_20it")
    meta = g["meta"]
    arch = _arch(meta)
    tr = rp.Trainer(arch, rp.Hyper(batch_size=meta["batch"]), rp.init_all(arch, meta["seed"]))
    shape = (meta["batch"], 1, *meta["image_size"])
    prints, marks = _batches(shape, meta["print_seed"]), _batches(shape, meta["mark_seed"])
    tol = chaos_envelope(g)
    assert tol[0] == 2e-4 and tol[1] < 1e-3  # the first iterations are gated tightly
    for it in range(meta["iters"]):
        d = tr.discriminator_step(next(prints), next(marks))
        gl = tr.generator_step(next(prints), next(marks))
        got = torch.tensor([d[0], d[1][0], d[1][1], gl[0], *gl[1]], dtype=torch.float64)
        want = g["losses"][it]
        rel = ((got - want).abs() / want.abs().clamp_min(1e-6))[LOSS_COLS].max().item()
        assert rel <= tol[it].item(), (it, rel, tol[it].item())
