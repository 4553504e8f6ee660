"""Helpers shared by the golden-vector tests."""

from pathlib import Path

import torch

GOLDEN = Path(__file__).resolve().parent / "golden"
CASES = ["a_32x32_down1", "b_64x64_default", "c_32x48_down2"]


def load(name):
    return torch.load(GOLDEN / f"{name}.pt", weights_only=False)


def fingerprint(t: torch.Tensor) -> torch.Tensor:
    """Must match tests/golden/make_golden.py::fingerprint."""
    f = t.detach().double().cpu().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, 16).long()
    return torch.cat([torch.stack([f.sum(), f.abs().sum(), f.norm()]), f[idx]])


def fp_err(got: torch.Tensor, want: torch.Tensor) -> float:
    """Error of a fingerprint relative to the tensor's own scale.

    Entries 1 (abs-sum) and 2 (l2) set the scale; the 16 samples and the plain
    sum are compared absolutely against rms-like magnitudes so near-zero
    samples do not blow up a relative measure."""
    want = want.double()
    got = got.double()
    l2 = want[2].abs().clamp_min(1e-30)
    e_norm = ((got[2] - want[2]).abs() / l2).item()
    e_abs = ((got[1] - want[1]).abs() / want[1].abs().clamp_min(1e-30)).item()
    smp_scale = want[3:].abs().max().clamp_min(1e-30)
    e_smp = ((got[3:] - want[3:]).abs().max() / smp_scale).item()
    return max(e_norm, e_abs, e_smp)


LOSS_COLS = [0, 3, 4, 5, 6, 7, 8, 9]  # disc, total_gen, gan, rec, idt, kl, path, style (not the signs)


def chaos_envelope(g, margin: float = 5.0) -> torch.Tensor:
    """Per-iteration relative tolerance for a long seeded run (golden case d): `margin` x the
    running maximum of the reference's fp32 losses against the fp64 replay of the same run
    (recorded by make_golden.py), floored at 2e-4 and capped at 0.5.  Training is chaotic: from
    iteration ~3 on two fp32 executions only agree statistically, and this is the measured size
    of that effect, not a guess."""
    ref, p64 = g["losses"].double(), g["losses_port_fp64"].double()
    rel = ((ref - p64).abs() / p64.abs().clamp_min(1e-6))[:, LOSS_COLS].max(dim=1).values
    env = torch.cummax(rel, 0).values
    return torch.clamp(margin * env, min=2e-4, max=0.5)
