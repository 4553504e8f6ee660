"""Checkpoint interchange on the CPU: the facades' state_dict layout is the reference's
(SURVEY.md §5), the run log prints the reference's line, and -- in the build container, where
/root/reference exists -- a state_dict produced by the REFERENCE's modules loads into the
facades with strict=True (and back)."""

import sys
from pathlib import Path

import pytest
import torch

from oracle import reference_port as rp

REF = Path("/root/reference")


def test_facade_state_dict_keys_match_the_oracle_layout():
    from one_to_many_gan_b200 import builder

    arch = rp.Arch(image_size=(64, 32), min_latent_resolution=16, n_resnet_blocks=3)
    P = rp.init_all(arch, 42)
    torch.manual_seed(42)
    D = builder.Discriminator(1)
    G = builder.Generator(1, 6, (64, 32), 16, 3)
    M = builder.MappingNetwork(6, 2, 0.9)
    S = builder.StyleExtractor(1, 6)
    for name, mod in (("D", D), ("G", G), ("M", M), ("S", S)):
        sd = mod.state_dict()
        assert list(sd) == list(P[name]) or set(sd) == set(P[name]), name
        for k, v in sd.items():
            assert torch.equal(v, P[name][k]), (name, k)


def test_run_log_line_is_the_reference_format():
    from one_to_many_gan_b200.checkpoint import RunLog

    log = RunLog(100)
    for i in range(2):
        log.record(disc=1.0 + i, sign_real=0.5, sign_fake=0.25, total_gen=10.0, gan=1.0, rec=2.0, idt=3.0,
                   kl=4.0, path=5.0, style=6.0, ada_p=0.0)
    assert log.line(50) == ("Step: 50/100, D loss: 1.5, D real/fake acc: 0.5/0.25, Total G loss: 10, "
                            "Gan loss 1, Idt loss 3, Rec loss 2, KL loss 4, Path loss 5, Style loss: 6, "
                            "ADA: 0, ")


@pytest.mark.skipif(not REF.exists(), reason="the reference checkout only exists in the build container")
def test_reference_state_dicts_load_into_the_facades_and_back():
    sys.path.insert(0, str(REF))
    sys.dont_write_bytecode = True
    from src.model import builder as ref  # the reference's own modules (read-only import)

    from one_to_many_gan_b200 import builder

    torch.manual_seed(7)
    pairs = [
        (ref.Discriminator(input_nc=1), builder.Discriminator(1)),
        (ref.Generator(1, 6, (64, 32), 16, 3), builder.Generator(1, 6, (64, 32), 16, 3)),
        (ref.MappingNetwork(6, 2, 0.9), builder.MappingNetwork(6, 2, 0.9)),
        (ref.StyleExtractor(1, 6), builder.StyleExtractor(1, 6)),
    ]
    for r, m in pairs:
        m.load_state_dict(r.state_dict(), strict=True)
        for (k1, v1), (k2, v2) in zip(r.state_dict().items(), m.state_dict().items()):
            assert k1 == k2 and torch.equal(v1, v2)
        r.load_state_dict(m.state_dict(), strict=True)
        # a torch.optim.Adam state of the reference module indexes parameters in the same order
        assert [tuple(p.shape) for p in r.parameters()] == [tuple(p.shape) for p in m.parameters()]


@pytest.mark.skipif(not REF.exists(), reason="the reference checkout only exists in the build container")
def test_train_py_torch_backend_runs_the_reference_path(tmp_path, capsys, monkeypatch):
    """`[training] backend = "torch"`: the one config flag selects the reference's own modules and
    step functions (identity stand-in for the un-installed pytorch-ada, exact while p == 0)."""
    from tests.test_train_gpu import TOML

    shim = tmp_path / "shim"
    shim.mkdir()
    (shim / "ada.py").write_text(
        "import torch\n"
        "class AdaptiveDiscriminatorAugmentation(torch.nn.Module):\n"
        "    def __init__(self, **kw):\n        super().__init__()\n"
        "    def set_p(self, p):\n        self.p = p\n"
        "    def forward(self, x):\n        return x\n")
    monkeypatch.syspath_prepend(str(REF))
    monkeypatch.syspath_prepend(str(shim))
    monkeypatch.syspath_prepend(str(Path(__file__).resolve().parents[1]))
    sys.dont_write_bytecode = True
    cfg = tmp_path / "cfg.toml"
    cfg.write_text(TOML.format(steps=2, ckpt=str(tmp_path / "ckpt"), precision="fp32", resume="false",
                               ckpt_every=100).replace('backend = "b200"', 'backend = "torch"'))
    import train

    prec, tf32 = torch.get_float32_matmul_precision(), torch.backends.cudnn.allow_tf32
    try:
        train.main(str(cfg))
    finally:  # the reference's train.py:67-68 settings are process-global
        torch.set_float32_matmul_precision(prec)
        torch.backends.cudnn.allow_tf32 = tf32
    out = capsys.readouterr().out
    assert "Step: 2/2, D loss:" in out and "ADA: 0" in out
