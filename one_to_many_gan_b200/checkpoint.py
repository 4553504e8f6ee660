"""Training-state checkpoints and the run log.

The checkpoint file is the dict the reference writes (src/core/evaluation.py:248-263: the eight
`*_state_dict` / `*_optim_state_dict` entries, `ada_p`, `image_buffer_images`,
`image_buffer_size`, saved as `<ckpt_dir>/<run>/models/<step>.tar`), so files interchange with
the reference in both directions: the module facades keep the reference's `state_dict` names and
`FlatAdam.state_dict()` has `torch.optim.Adam`'s layout.  One extra key, `b200_resume`, carries
what the reference never saves but a restart needs to CONTINUE rather than approximately
continue (iteration count, ADA controller window, host RNG states); the reference ignores it.

The reference has no loader at all (`infinite_run.sh` restarts from step 0);
`load_checkpoint` is that missing half."""

from __future__ import annotations

import random
from pathlib import Path

import torch

NETS = ("generator", "discriminator", "mapping_network", "style_extractor")


def checkpoint_path(config, step: int) -> Path:
    t = config["training"]
    return Path(t["checkpoint_directory"]) / t["training_run"] / "models" / f"{step}.tar"


def latest_checkpoint(config) -> Path | None:
    d = checkpoint_path(config, 0).parent
    if not d.is_dir():
        return None
    found = [(int(p.stem), p) for p in d.glob("*.tar") if p.stem.isdigit()]
    return max(found)[1] if found else None


def save_checkpoint(path, *, nets: dict, optimisers: dict, ada_p, pool_images, pool_size: int,
                    step: int) -> Path:
    """nets / optimisers: dicts keyed by NETS.  pool_images: list of [1,C,H,W] tensors (the
    reference's ImageBuffer.images)."""
    path = Path(path)
    path.parent.mkdir(parents=True, exist_ok=True)
    blob = {}
    for name in NETS:
        blob[f"{name}_state_dict"] = nets[name].state_dict()
        blob[f"{name}_optim_state_dict"] = optimisers[name].state_dict()
    blob["ada_p"] = float(ada_p())
    blob["image_buffer_images"] = [t.detach().clone() for t in pool_images]
    blob["image_buffer_size"] = int(pool_size)
    blob["b200_resume"] = {
        "step": int(step),
        "ada_curr_batch": int(getattr(ada_p, "curr_batch", 0)),
        "ada_scores": [float(s) for s in getattr(ada_p, "mean_real_scores", [])],
        "torch_rng": torch.get_rng_state(),
        "python_rng": random.getstate(),
        "cuda_rng": torch.cuda.get_rng_state() if torch.cuda.is_available() else None,
    }
    torch.save(blob, path)
    return path


def load_checkpoint(path, *, nets: dict, optimisers: dict, ada_p=None, device=None,
                    restore_rng: bool = True) -> dict:
    """Restore the four networks and optimisers in place; returns
    {step, pool_images, pool_size, ada_p}.  Accepts the reference's own checkpoints (no
    `b200_resume` key: the step is then taken from the Adam state)."""
    blob = torch.load(Path(path), map_location=device or "cpu", weights_only=False)
    for name in NETS:
        nets[name].load_state_dict(blob[f"{name}_state_dict"])
        optimisers[name].load_state_dict(blob[f"{name}_optim_state_dict"])
    from . import ops

    ops.invalidate_packs()  # the weights changed behind the staging cache
    extra = blob.get("b200_resume", {})
    step = extra.get("step")
    if step is None:
        st = blob["generator_optim_state_dict"]["state"]
        step = int(next(iter(st.values()))["step"]) if st else 0
    if ada_p is not None:
        ada_p.p = torch.tensor(float(blob["ada_p"]))
        ada_p.curr_batch = int(extra.get("ada_curr_batch", 0))
        ada_p.mean_real_scores = [torch.tensor(s) for s in extra.get("ada_scores", [])]
    if restore_rng and "torch_rng" in extra:
        torch.set_rng_state(extra["torch_rng"].cpu())
        random.setstate(extra["python_rng"])
        if extra.get("cuda_rng") is not None and torch.cuda.is_available():
            torch.cuda.set_rng_state(extra["cuda_rng"].cpu())
    return {"step": int(step), "pool_images": blob["image_buffer_images"],
            "pool_size": int(blob["image_buffer_size"]), "ada_p": float(blob["ada_p"])}


class RunLog:
    """Means of the per-iteration scalars between two log lines, printed in the reference's
    format (src/core/evaluation.py:290-308) so existing log parsers keep working."""

    FIELDS = ("disc", "sign_real", "sign_fake", "total_gen", "gan", "idt", "rec", "kl", "path",
              "style", "ada_p")

    def __init__(self, training_steps: int):
        self.training_steps = training_steps
        self._sum = dict.fromkeys(self.FIELDS, 0.0)
        self._n = 0

    def record(self, **values: float) -> None:
        for k in self.FIELDS:
            self._sum[k] += float(values[k])
        self._n += 1

    def line(self, step: int) -> str:
        n = max(self._n, 1)
        m = {k: f"{v / n:.6g}" for k, v in self._sum.items()}
        self._sum = dict.fromkeys(self.FIELDS, 0.0)
        self._n = 0
        return (f"Step: {step}/{self.training_steps}, D loss: {m['disc']}, "
                f"D real/fake acc: {m['sign_real']}/{m['sign_fake']}, "
                f"Total G loss: {m['total_gen']}, Gan loss {m['gan']}, Idt loss {m['idt']}, "
                f"Rec loss {m['rec']}, KL loss {m['kl']}, Path loss {m['path']}, "
                f"Style loss: {m['style']}, ADA: {m['ada_p']}, ")
