"""train.py end to end on the GPU (reference train.py:28-319): the b200 backend through the
CUDA-graph engine on synthetic data, the reference's log line and checkpoint layout, and the
resume path the reference lacks: save -> load -> continue must reproduce an
uninterrupted run."""

import re
import sys
from pathlib import Path

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parents[1]

TOML = """
[training]
batch_size = 2
random_seed = 42
training_steps = {steps}
image_buffer_size = 5
style_mixing_prob = 0.9
deterministic_cuda_kernels = false
gpu_number = 0
checkpoint_directory = "{ckpt}"
training_run = "run"
backend = "b200"
precision = "{precision}"
synthetic_data = true
resume = {resume}

[optimisation]
style_cycle_loss_lambda = 5.0
identity_loss_lambda = 5.0
reconstruction_loss_lambda = 5.0
kl_loss_lambda = 0.01
path_loss_lambda = 0.1
path_loss_jacobian_granularity = [0.1, 0.2]
learning_rate = 2e-3
mapping_network_learning_rate = 2e-5
adam_betas = [0.5, 0.99]

[ada]
discriminator_real_acc_target = 0.6
ada_overfitting_measurement_n_images = 256
ada_adjustment_size = 5.12e-4

[evaluation]
log_interval = 2
checkpoint_interval = {ckpt_every}
n_evaluation_images = 16
inference_batch_size = 2

[architecture]
w_dim = 6
add_latent_noise = false
min_latent_resolution = 16
n_resnet_blocks = 3
mapping_network_layers = 2

[data]
image_size = [32, 32]
image_channels = 1
shoemark_data_dir = "/nonexistent"
shoeprint_data_dir = "/nonexistent"
"""


def _run(tmp, steps, precision="fp32", resume="false", ckpt_every=2):
    sys.path.insert(0, str(ROOT))
    import train

    cfg = tmp / f"cfg_{steps}_{resume}.toml"
    cfg.write_text(TOML.format(steps=steps, ckpt=str(tmp / "ckpt"), precision=precision, resume=resume,
                               ckpt_every=ckpt_every))
    return train.main(str(cfg))


def test_train_py_runs_logs_and_checkpoints(tmp_path, capsys):
    eng = _run(tmp_path, 5, precision="bf16")
    assert eng.iterations == 5 and eng.graph is not None  # 2 eager warm-ups, then replays
    log = (tmp_path / "ckpt" / "run" / "log").read_text().strip().splitlines()
    assert len(log) == 3  # steps 2, 4 and the last
    pat = (r"Step: (\d+)/5, D loss: \S+, D real/fake acc: \S+/\S+, Total G loss: \S+, Gan loss \S+, "
           r"Idt loss \S+, Rec loss \S+, KL loss \S+, Path loss \S+, Style loss: \S+, ADA: 0, ")
    assert [int(re.fullmatch(pat, l.rstrip() + " ").group(1)) for l in log] == [2, 4, 5]
    files = sorted(p.name for p in (tmp_path / "ckpt" / "run" / "models").iterdir())
    assert files == ["2.tar", "4.tar", "5.tar"]
    blob = torch.load(tmp_path / "ckpt" / "run" / "models" / "5.tar", weights_only=False)
    want = {f"{n}_{k}" for n in ("generator", "discriminator", "mapping_network", "style_extractor")
            for k in ("state_dict", "optim_state_dict")} | {"ada_p", "image_buffer_images",
                                                            "image_buffer_size"}
    assert want <= set(blob) and blob["image_buffer_size"] == 5
    assert len(blob["image_buffer_images"]) == 5 and blob["image_buffer_images"][0].shape == (1, 1, 32, 32)
    # torch.optim.Adam can load the optimiser state (checkpoints interchange with the reference)
    from one_to_many_gan_b200 import builder

    G = builder.Generator(1, 6, (32, 32), 16, 3)
    G.load_state_dict(blob["generator_state_dict"], strict=True)
    opt = torch.optim.Adam(G.parameters(), lr=1.0)
    opt.load_state_dict(blob["generator_optim_state_dict"])
    assert opt.param_groups[0]["lr"] == 2e-3 and opt.state_dict()["state"][0]["step"] == 5


def test_resume_reproduces_uninterrupted_run(tmp_path):
    """An uninterrupted run checkpoints at step 5 and finishes at step 6; a second process state
    loads THAT checkpoint and runs step 6.  Same weights, Adam moments, pool, host RNG states and
    data-stream position => the same step up to the summation order of the fp32 atomics.  (Longer
    horizons cannot be compared tightly: this tiny model amplifies that noise ~10x per iteration.)"""
    import shutil

    a = tmp_path / "a"
    b = tmp_path / "b"
    a.mkdir()
    b.mkdir()
    straight = _run(a, 6, ckpt_every=5)
    w_straight = [o.param_arena.clone() for o in (straight.oD, straight.oG, straight.oM, straight.oS)]
    pool_straight = straight.pool.clone()
    models = b / "ckpt" / "run" / "models"
    models.mkdir(parents=True)
    shutil.copy(a / "ckpt" / "run" / "models" / "5.tar", models / "5.tar")
    resumed = _run(b, 6, resume="true", ckpt_every=100)
    assert resumed.iterations == 1                    # continued at step 5, did not start over
    before = torch.load(models / "5.tar", weights_only=False)["generator_state_dict"]
    moved = (resumed.G.state_dict()["encoder.1.weight.weight"].cpu()
             - before["encoder.1.weight.weight"].cpu()).abs().mean().item()
    assert moved > 1e-4                               # the resumed step did update the weights
    for x, o in zip(w_straight, (resumed.oD, resumed.oG, resumed.oM, resumed.oS)):
        assert (x - o.param_arena).abs().mean().item() < 1e-5
        assert ((x - o.param_arena).double().norm() / x.double().norm()).item() < 1e-4
    assert (pool_straight - resumed.pool).abs().max().item() < 1e-3
    assert resumed.oG.steps == 6 and int(resumed.oG.step_dev.item()) == 6
    assert (b / "ckpt" / "run" / "models" / "6.tar").exists()


def test_ada_probability_is_never_silently_dropped(tmp_path):
    """Once the controller asks for p > 0 the b200 backend stops unless the run opted out."""
    sys.path.insert(0, str(ROOT))
    import train
    from one_to_many_gan_b200 import training

    orig = training.ADAp.__call__
    training.ADAp.__call__ = lambda self: 0.25
    try:
        with pytest.raises(SystemExit, match="allow_identity"):
            _run(tmp_path, 2)
    finally:
        training.ADAp.__call__ = orig
    del train
