"""CPU-side checks of the drop-in boundary: the C-ABI library builds for sm_100a, loads, and
exports every symbol include/otm_b200.h declares; calls without a GPU fail loudly."""

import ctypes
import re
from pathlib import Path

import pytest
import torch

ROOT = Path(__file__).resolve().parents[1]


def _declared():
    text = (ROOT / "include" / "otm_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(otm_[A-Za-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from one_to_many_gan_b200 import _lib

    names = _declared()
    assert len(names) >= 30
    assert set(names) == set(_lib.SYMBOLS), set(names) ^ set(_lib.SYMBOLS)
    raw = ctypes.CDLL(str(_lib.LIB_PATH))
    for n in names:
        assert hasattr(raw, n), n
    assert _lib.lib.otm_version() == 1


def test_sass_contains_blackwell_opcodes():
    import shutil
    import subprocess

    from one_to_many_gan_b200 import _lib

    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    if not Path(cuobjdump).exists():
        pytest.skip("cuobjdump not available")
    sass = subprocess.run([cuobjdump, "-sass", str(_lib.LIB_PATH)], capture_output=True, text=True).stdout
    for op in ("UTCHMMA", "UTMALDG", "LDTM", "UTCBAR"):
        assert op in sass, f"{op} missing: the conv kernels are not tcgen05/TMA code"
    # warp-level mma.sync is allowed ONLY in the four thin layers (Cin = 1 / Cout = 1: K = 16/49 or
    # N = 1 is below a tcgen05 tile) and in the halo-ring correction of the reflect-pad dgrad (M =
    # 66 ring positions per line, 0.6 % of the dgrad's MACs); every dense conv kernel must be tcgen05
    for chunk in sass.split("Function : ")[1:]:
        name = chunk.split("\n", 1)[0]
        if "HMMA." in chunk.replace("UTCHMMA", ""):
            assert any(k in name for k in ("cin1_mma", "cout1_mma", "conv_reflect_border")), \
                f"legacy mma.sync in {name}"


def test_cpu_tensor_is_rejected():
    from one_to_many_gan_b200 import _lib, kernels

    x = torch.zeros(1, 8, 4, 4).contiguous(memory_format=torch.channels_last)
    with pytest.raises(_lib.OtmError):
        kernels.instnorm_stats(x)


def test_bad_geometry_reports_error():
    from one_to_many_gan_b200 import _lib

    a = _lib.ConvFwdArgs()
    rc = _lib.lib.otm_conv_fwd(ctypes.byref(a), None)
    assert rc == -1
    assert b"null" in _lib.lib.otm_last_error()
