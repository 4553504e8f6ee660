"""TOML hyper-parameters (same keys as reference config.toml / src/data/config.py:71-85) plus
the optional keys that select the B200 path.  Unknown keys are tolerated, exactly like the
reference's bare `tomllib.load`."""

from __future__ import annotations

import tomllib
from pathlib import Path

import torch

# new, optional keys (default = reference behaviour)
DEFAULTS = {
    "training": {"backend": "b200", "precision": "fp32", "synthetic_data": False,
                 "styles_per_input": 1},
    "architecture": {"start_filters": 64},
}


def load_config(path):
    path = Path(path)
    with path.open("rb") as f:
        config = tomllib.load(f)
    config["training"]["checkpoint_directory"] = Path(config["training"]["checkpoint_directory"])
    config["data"]["shoeprint_data_dir"] = Path(config["data"]["shoeprint_data_dir"])
    config["data"]["shoemark_data_dir"] = Path(config["data"]["shoemark_data_dir"])
    for section, values in DEFAULTS.items():
        for key, val in values.items():
            config.setdefault(section, {}).setdefault(key, val)
    return config


def act_dtype(config) -> torch.dtype:
    prec = config["training"].get("precision", "fp32")
    if prec not in ("fp32", "bf16"):
        raise ValueError(f"[training] precision must be 'fp32' or 'bf16', got {prec!r}")
    return torch.bfloat16 if prec == "bf16" else torch.float32
