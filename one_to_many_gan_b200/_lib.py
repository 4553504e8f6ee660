"""ctypes binding of libotm_b200.so (C-ABI declared in include/otm_b200.h).

There is no CPU fallback: if the shared library is missing the import fails loudly, and
every op raises when handed a non-CUDA tensor."""

from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import torch

_PKG = Path(__file__).resolve().parent
# OTM_B200_LIB selects another build of the same library (kernel tuning experiments)
LIB_PATH = Path(os.environ["OTM_B200_LIB"]) if os.environ.get("OTM_B200_LIB") else _PKG / "libotm_b200.so"

F32, BF16 = 0, 1
ACT_NONE, ACT_RELU, ACT_LRELU, ACT_TANH = 0, 1, 2, 3
PATH_AUTO, PATH_SIMT, PATH_TC = 0, 1, 2


class Tensor(C.Structure):
    _fields_ = [
        ("ptr", C.c_void_p),
        ("dtype", C.c_int32),
        ("n", C.c_int32),
        ("h", C.c_int32),
        ("w", C.c_int32),
        ("c", C.c_int32),
        ("sn", C.c_int64),
        ("sh", C.c_int64),
        ("sw", C.c_int64),
    ]


class ConvFwdArgs(C.Structure):
    _fields_ = [
        ("x", Tensor),
        ("x_halo", C.c_int32),
        ("wpack", C.c_void_p),
        ("w_batch_stride", C.c_int64),
        ("kh", C.c_int32),
        ("kw", C.c_int32),
        ("pad", C.c_int32),
        ("y", Tensor),
        ("y_halo", C.c_int32),
        ("alpha", C.c_float),
        ("row_scale", C.c_void_p),
        ("bias", C.c_void_p),
        ("act", C.c_int32),
        ("residual", Tensor),
        ("path", C.c_int32),
        ("post_scale", C.c_void_p),
        ("stat_sums", C.c_void_p),
        ("residual_mode", C.c_int32),
        ("dot_sums", C.c_void_p),
    ]


class ConvReflectBorderArgs(C.Structure):
    _fields_ = [
        ("dy", Tensor),
        ("wpack", C.c_void_p),
        ("w_batch_stride", C.c_int64),
        ("y", Tensor),
        ("row_scale", C.c_void_p),
        ("post_scale", C.c_void_p),
        ("gate", Tensor),
        ("dot_sums", C.c_void_p),
    ]


class ConvWgradArgs(C.Structure):
    _fields_ = [
        ("x", Tensor),
        ("x_halo", C.c_int32),
        ("dy", Tensor),
        ("kh", C.c_int32),
        ("kw", C.c_int32),
        ("pad", C.c_int32),
        ("dw", C.c_void_p),
        ("alpha", C.c_float),
        ("rs", C.c_void_p),
        ("cs", C.c_void_p),
        ("path", C.c_int32),
        ("ws", C.c_void_p),
        ("wfwd", C.c_void_p),
        ("P", C.c_void_p),
        ("wfwd_batch_stride", C.c_int64),
        ("Q", C.c_void_p),
    ]


class WeightPackArgs(C.Structure):
    _fields_ = [
        ("w", C.c_void_p),
        ("cout", C.c_int32),
        ("cin", C.c_int32),
        ("kh", C.c_int32),
        ("kw", C.c_int32),
        ("alpha", C.c_float),
        ("cs", C.c_void_p),
        ("rs", C.c_void_p),
        ("nb", C.c_int32),
        ("transpose", C.c_int32),
        ("out", C.c_void_p),
        ("out_dtype", C.c_int32),
    ]


class ModBwdArgs(C.Structure):
    _fields_ = [
        ("w", C.c_void_p),
        ("cout", C.c_int32),
        ("cin", C.c_int32),
        ("taps", C.c_int32),
        ("alpha", C.c_float),
        ("s", C.c_void_p),
        ("sigma_inv", C.c_void_p),
        ("q", C.c_void_p),
        ("P", C.c_void_p),
        ("Q", C.c_void_p),
        ("nb", C.c_int32),
        ("ds", C.c_void_p),
        ("dw", C.c_void_p),
        ("q_scaled", C.c_int32),
    ]


class NormActArgs(C.Structure):
    _fields_ = [
        ("x", Tensor),
        ("stats", C.c_void_p),
        ("act", C.c_int32),
        ("residual", Tensor),
        ("y", Tensor),
        ("y_halo", C.c_int32),
    ]


class NormActBwdArgs(C.Structure):
    _fields_ = [
        ("g", Tensor),
        ("g_halo", C.c_int32),
        ("g2", Tensor),
        ("x", Tensor),
        ("stats", C.c_void_p),
        ("act", C.c_int32),
        ("gx", Tensor),
        ("gres", Tensor),
        ("sums", C.c_void_p),
        ("g_down", C.c_int32),
    ]


class DownArgs(C.Structure):
    _fields_ = [
        ("x", Tensor),
        ("stats", C.c_void_p),
        ("act", C.c_int32),
        ("y", Tensor),
        ("y_halo", C.c_int32),
    ]


class ModOutArgs(C.Structure):
    _fields_ = [
        ("g", Tensor),
        ("g_halo", C.c_int32),
        ("g2", Tensor),
        ("out", Tensor),
        ("res", Tensor),
        ("act", C.c_int32),
        ("gy", Tensor),
        ("P", C.c_void_p),
        ("gy_scale", C.c_void_p),
    ]


class ModInArgs(C.Structure):
    _fields_ = [
        ("g", Tensor),
        ("g_halo", C.c_int32),
        ("x", Tensor),
        ("s", C.c_void_p),
        ("gadd", Tensor),
        ("gx", Tensor),
        ("Q", C.c_void_p),
        ("relu_mask", C.c_int32),
        ("gx_scale", C.c_void_p),
    ]


class AdamArgs(C.Structure):
    _fields_ = [
        ("param", C.c_void_p),
        ("grad", C.c_void_p),
        ("m", C.c_void_p),
        ("v", C.c_void_p),
        ("n", C.c_int64),
        ("lr", C.c_float),
        ("beta1", C.c_float),
        ("beta2", C.c_float),
        ("eps", C.c_float),
        ("grad_scale", C.c_float),
        ("step", C.c_void_p),
    ]


MAX_LINEAR_JOBS, MAX_STYLE_DIM, MAX_MAPPING_LAYERS = 16, 32, 8


class LinearJob(C.Structure):
    _fields_ = [
        ("x", C.c_void_p),
        ("x_row_stride", C.c_int64),
        ("w", C.c_void_p),
        ("bias", C.c_void_p),
        ("y", C.c_void_p),
        ("dy", C.c_void_p),
        ("dw", C.c_void_p),
        ("dbias", C.c_void_p),
        ("dx", C.c_void_p),
        ("dx_row_stride", C.c_int64),
        ("n", C.c_int32),
        ("k", C.c_int32),
        ("o", C.c_int32),
    ]


class MappingArgs(C.Structure):
    _fields_ = [
        ("z1", C.c_void_p),
        ("z2", C.c_void_p),
        ("cross", C.c_void_p),
        ("w", C.c_void_p * MAX_MAPPING_LAYERS),
        ("b", C.c_void_p * MAX_MAPPING_LAYERS),
        ("dw", C.c_void_p * MAX_MAPPING_LAYERS),
        ("db", C.c_void_p * MAX_MAPPING_LAYERS),
        ("features", C.c_int32),
        ("n_layers", C.c_int32),
        ("batch", C.c_int32),
        ("n_blocks", C.c_int32),
        ("d", C.c_void_p * 2),
        ("d_const", C.c_float * 2),
        ("out", C.c_void_p * 2),
        ("dout", C.c_void_p * 2),
    ]


# every symbol include/otm_b200.h declares: (restype, argtypes)
_P = C.POINTER
SYMBOLS = {
    "otm_last_error": (C.c_char_p, []),
    "otm_version": (C.c_int, []),
    "otm_launch_count": (C.c_int64, []),
    "otm_conv_fwd": (C.c_int, [_P(ConvFwdArgs), C.c_void_p]),
    "otm_conv_fwd_uses_tcgen05": (C.c_int, [_P(ConvFwdArgs)]),
    "otm_conv_fwd_fuses_stats": (C.c_int, [_P(ConvFwdArgs)]),
    "otm_conv_fwd_fuses_gate": (C.c_int, [_P(ConvFwdArgs)]),
    "otm_conv_reflect_border": (C.c_int, [_P(ConvReflectBorderArgs), C.c_void_p]),
    "otm_conv_wgrad": (C.c_int, [_P(ConvWgradArgs), C.c_void_p]),
    "otm_conv_wgrad_uses_tcgen05": (C.c_int, [_P(ConvWgradArgs)]),
    "otm_conv_wgrad_fuses_P": (C.c_int, [_P(ConvWgradArgs)]),
    "otm_conv_wgrad_fuses_Q": (C.c_int, [_P(ConvWgradArgs)]),
    "otm_conv_wgrad_workspace_bytes": (C.c_int64, [_P(ConvWgradArgs)]),
    "otm_weight_pack": (C.c_int, [_P(WeightPackArgs), C.c_void_p]),
    "otm_weight_pack_multi": (C.c_int, [_P(WeightPackArgs), C.c_int32, C.c_void_p]),
    "otm_weight_sqsum": (
        C.c_int,
        [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p],
    ),
    "otm_demod": (
        C.c_int,
        [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_float, C.c_void_p, C.c_void_p],
    ),
    "otm_mod_bwd": (C.c_int, [_P(ModBwdArgs), C.c_void_p]),
    "otm_instnorm_finalize": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_float, C.c_void_p]),
    "otm_instnorm_stats": (C.c_int, [_P(Tensor), C.c_float, C.c_void_p, C.c_void_p, C.c_void_p]),
    "otm_norm_act": (C.c_int, [_P(NormActArgs), C.c_void_p]),
    "otm_norm_act_bwd": (C.c_int, [_P(NormActBwdArgs), C.c_void_p]),
    "otm_norm_act_bwd_bwd": (
        C.c_int,
        [_P(Tensor), _P(Tensor), _P(Tensor), C.c_void_p, C.c_int32, _P(Tensor), _P(Tensor), C.c_void_p,
         C.c_void_p],
    ),
    "otm_down": (C.c_int, [_P(DownArgs), C.c_void_p]),
    "otm_down_bwd": (C.c_int, [_P(Tensor), C.c_int32, _P(Tensor), C.c_void_p]),
    "otm_up": (C.c_int, [_P(Tensor), _P(Tensor), C.c_int32, C.c_void_p, C.c_void_p]),
    "otm_up_bwd": (C.c_int, [_P(Tensor), C.c_int32, _P(Tensor), C.c_void_p, C.c_void_p]),
    "otm_mod_out": (C.c_int, [_P(ModOutArgs), C.c_void_p]),
    "otm_mod_in": (C.c_int, [_P(ModInArgs), C.c_void_p]),
    "otm_channel_sum": (C.c_int, [_P(Tensor), C.c_void_p, C.c_int32, C.c_void_p]),
    "otm_loss_style_cycle": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_float, C.c_float,
         C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p],
    ),
    "otm_linear_fwd": (C.c_int, [_P(LinearJob), C.c_int32, C.c_void_p]),
    "otm_linear_bwd": (C.c_int, [_P(LinearJob), C.c_int32, C.c_void_p]),
    "otm_mapping_fwd": (C.c_int, [_P(MappingArgs), C.c_void_p]),
    "otm_mapping_bwd": (C.c_int, [_P(MappingArgs), C.c_void_p]),
    "otm_avgpool": (C.c_int, [_P(Tensor), C.c_void_p, C.c_void_p]),
    "otm_avgpool_bwd": (C.c_int, [C.c_void_p, _P(Tensor), C.c_void_p]),
    "otm_loss_lsgan": (
        C.c_int,
        [_P(Tensor), C.c_float, C.c_float, C.c_void_p, _P(Tensor), C.c_void_p],
    ),
    "otm_loss_l1": (C.c_int, [_P(Tensor), _P(Tensor), C.c_float, C.c_void_p, _P(Tensor), C.c_void_p]),
    "otm_moments": (C.c_int, [_P(Tensor), C.c_void_p, C.c_void_p]),
    "otm_affine_grad": (C.c_int, [_P(Tensor), C.c_void_p, _P(Tensor), C.c_int32, C.c_void_p]),
    "otm_loss_path": (
        C.c_int,
        [_P(Tensor), _P(Tensor), C.c_void_p, C.c_float, C.c_float, C.c_void_p, _P(Tensor),
         _P(Tensor), C.c_void_p],
    ),
    "otm_adam": (C.c_int, [_P(AdamArgs), C.c_void_p]),
    "otm_synth_uniform": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p],
    ),
    "otm_gather_batch": (
        C.c_int,
        [C.c_void_p, C.c_int64, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32,
         C.c_void_p, C.c_void_p],
    ),
    "otm_cast": (C.c_int, [_P(Tensor), _P(Tensor), C.c_void_p]),
    "otm_add_inplace": (C.c_int, [_P(Tensor), _P(Tensor), C.c_void_p]),
}


def _load() -> C.CDLL:
    if not LIB_PATH.exists():
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -m one_to_many_gan_b200.build` "
            "(there is no CPU/PyTorch fallback for the B200 path)"
        )
    lib = C.CDLL(str(LIB_PATH))
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # raises AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    return lib


lib = _load()


class OtmError(RuntimeError):
    pass


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        raise OtmError(f"{what}: {lib.otm_last_error().decode()} (code {rc})")


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def dtype_code(t: torch.Tensor) -> int:
    if t.dtype == torch.float32:
        return F32
    if t.dtype == torch.bfloat16:
        return BF16
    raise TypeError(f"unsupported dtype {t.dtype}")


NULL_T = Tensor()


def tdesc(t: torch.Tensor | None) -> Tensor:
    """Describe a logically-NCHW tensor whose channel stride is 1 (NHWC memory)."""
    if t is None:
        return Tensor()
    if not t.is_cuda:
        raise OtmError("otm_b200 ops need CUDA tensors (no CPU fallback)")
    if t.dim() != 4:
        raise ValueError(f"expected [N,C,H,W], got {tuple(t.shape)}")
    n, c, h, w = t.shape
    sn, sc, sh, sw = t.stride()
    if c != 1 and sc != 1:
        raise ValueError(f"tensor must be channels-last (channel stride 1), strides={t.stride()}")
    return Tensor(t.data_ptr(), dtype_code(t), n, h, w, c, sn, sh, sw)


def ptr(t: torch.Tensor | None) -> int | None:
    if t is None:
        return None
    if not t.is_cuda:
        raise OtmError("otm_b200 ops need CUDA tensors (no CPU fallback)")
    return t.data_ptr()
