"""ctypes binding of libskeldiff_sm100a.so (the C ABI declared in include/skeldiff_b200.h).

There is deliberately no fallback: if the shared object is missing, or a CUDA device is not
available when an operator is called, the call raises.  PyTorch is used only for device memory,
streams and (elsewhere) torch.distributed.
"""
from __future__ import annotations

import ctypes as C
import os
import threading
from pathlib import Path
from typing import Optional

import torch

_LIB_PATH = Path(__file__).resolve().parent / "csrc" / "libskeldiff_sm100a.so"
_lib: Optional[C.CDLL] = None

PREC_FP32, PREC_BF16, PREC_BF16X3 = 0, 1, 2
ACT_NONE, ACT_TANH, ACT_TANH_TANH = 0, 1, 2
PREC_F16X2 = 3
PRECISIONS = {"fp32": PREC_FP32, "bf16": PREC_BF16, "bf16x3": PREC_BF16X3, "fp16x2": PREC_F16X2}
FP32_GRADE_TC = ("bf16x3", "fp16x2")       # tensor-core precisions that meet the fp32 parity gate


class NativeError(RuntimeError):
    pass


class SdView(C.Structure):
    _fields_ = [("ptr", C.c_void_p), ("sample_stride", C.c_int64), ("node_stride", C.c_int64),
                ("rep", C.c_int32), ("width", C.c_int32)]


class SdGlinArgs(C.Structure):
    _fields_ = [("a0", SdView), ("a1", SdView), ("row_scale_dev", C.c_void_p),
                ("scale_shift_dev", C.c_void_p), ("ss_row_dev", C.c_void_p), ("ss_row", C.c_int32),
                ("ss_row_stride", C.c_int64), ("act", C.c_int32), ("residual", SdView), ("out", SdView),
                ("scratch_dev", C.c_void_p), ("batch", C.c_int32), ("precision", C.c_int32)]


# name -> (restype, argtypes); must list every symbol include/skeldiff_b200.h declares
_P, _I, _I64, _U64, _F, _SZ = C.c_void_p, C.c_int, C.c_int64, C.c_uint64, C.c_float, C.c_size_t
SIGNATURES = {
    "sd_last_error": (C.c_char_p, []),
    "sd_version": (_I, []),
    "sd_launch_count": (_U64, []),
    "sd_device_supported": (_I, [_I]),
    "sd_glin_create": (_I, [_I, _P, _I, _I, _I, _P, _P, _P, C.POINTER(_P)]),
    "sd_glin_set_bf16": (_I, [_P, _P, _I]),
    "sd_glin_set_kmajor": (_I, [_P, _P]),
    "sd_glin_destroy": (None, [_P]),
    "sd_glin_forward": (_I, [_P, C.POINTER(SdGlinArgs), _P]),
    "sd_glin_forward_bf16": (_I, [_P, _P, _P, _P, _I, _P, _P, _I, _P, _I, _P]),
    "sd_node_attention": (_I, [_P, _P, _I, _I, _I, _I, _P]),
    "sd_row_inv_norm": (_I, [_P, _P, _I64, _I, _P]),
    "sd_time_table": (_I, [_P, _I, _I, _F, _I, _P, _P, _P, _P, _P, _P, _I, _P, _P, _P]),
    "sd_denoiser_create": (_I, [_I, _I, _I, _I, _I, _I, _I, C.POINTER(_P)]),
    "sd_denoiser_set_layer": (_I, [_P, _I, _P]),
    "sd_denoiser_set_time_table": (_I, [_P, _P, _I]),
    "sd_denoiser_destroy": (None, [_P]),
    "sd_denoiser_workspace_bytes": (_SZ, [_P, _I, _I]),
    "sd_denoiser_forward": (_I, [_P, C.POINTER(SdView), C.POINTER(SdView), _P, _I, _P, _I, _P, _I, _P]),
    "sd_diffusion_create": (_I, [_I, _I, _I, _P, _P, _P, _P, _P, _P, C.POINTER(_P)]),
    "sd_diffusion_destroy": (None, [_P]),
    "sd_reverse_step": (_I, [_P, _P, _P, C.POINTER(SdView), _P, _P, _I, _I, _I, _P]),
    "sd_q_sample": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "sd_mahalanobis_loss": (_I, [_P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "sd_sample_workspace_bytes": (_SZ, [_P, _P, _I, _I]),
    "sd_sample_loop": (_I, [_P, _P, _P, C.POINTER(SdView), _P, _P, _I, _I, _P, _I, _P]),
    "sd_fill_normal": (_I, [_P, _I64, _U64, _U64, _P]),
    "sd_motion_metrics": (_I, [_P, _P, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "sd_best_sample": (_I, [_P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "sd_multimodal_metrics": (_I, [_P, _P, _P, _P, _I, _I, _I, _I, _I, _F, _P, _P, _P, _P]),
    "sd_node_mix_transposed": (_I, [_P, _P, _P, _I, _I, _I, _P]),
    "sd_glin_backward_scratch_bytes": (C.c_size_t, [_P, _I]),
    "sd_glin_backward_params": (_I, [_P, _P, _P, _P, _P, _P, _P, _P, _P, _P, _I, _I, _P]),
    "sd_ss_tanh_forward": (_I, [_P, _P, _P, _P, _I, _I, _I, _P]),
    "sd_ss_tanh_backward": (_I, [_P, _P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "sd_rmsnorm_forward": (_I, [_P, _P, _P, _P, _I64, _I, _P]),
    "sd_rmsnorm_backward_blocks": (_I, [_I64]),
    "sd_rmsnorm_backward": (_I, [_P, _P, _P, _P, _P, _P, _I64, _I, _P]),
    "sd_node_attention_backward": (_I, [_P, _P, _P, _I, _I, _I, _I, _P]),
    "sd_mahalanobis_loss_backward": (_I, [_P, _P, _P, _P, _P, _P, _I, _I, _I, _P]),
    "sd_gru_create": (_I, [_I, _P, _I, _I, _I, _P, _P, _P, _P, _P, _I, C.POINTER(_P)]),
    "sd_gru_set_fused": (_I, [_P, _P, _P, _P, _P]),
    "sd_gru_set_fused_f16x2": (_I, [_P, _P]),
    "sd_gru_set_bf16x3": (_I, [_P, _P]),
    "sd_gru_set_f16x2": (_I, [_P, _P]),
    "sd_glin_set_f16x2": (_I, [_P, _P]),
    "sd_gru_destroy": (None, [_P]),
    "sd_encode_workspace_bytes": (_SZ, [_I, _I, _I, _I, _I]),
    "sd_encode": (_I, [_P, _P, _I, _P, _P, _I, _I, _I, _P, _I, _P, _I, _P]),
    "sd_decode_workspace_bytes": (_SZ, [_I, _I, _I]),
    "sd_decode": (_I, [_P, _P, _P, C.POINTER(SdView), C.POINTER(SdView), _P, _I, _I, _I, _P, _P, _I, _P]),
}


def library_path() -> Path:
    return _LIB_PATH


def load() -> C.CDLL:
    """Load the shared object and bind every declared symbol (raises if anything is missing)."""
    global _lib
    if _lib is not None:
        return _lib
    if not _LIB_PATH.exists():
        raise NativeError(f"{_LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(nvcc, sm_100a). There is no CPU/PyTorch fallback for this path.")
    lib = C.CDLL(os.fspath(_LIB_PATH))
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def last_error() -> str:
    return load().sd_last_error().decode("utf-8", "replace")


# Device guard.  The library launches on the CURRENT CUDA device (and sizes its grids from that device's SM count), while the
# caller's tensors may live on another one (a model on cuda:1 with cuda:0 current).  Every native call in this package has the
# shape  check(lib.fn(..., stream_ptr(device)), "fn"):  stream_ptr switches to the tensors' device while the arguments are
# being evaluated, check() switches back once the call has returned (also when it failed).
_guard = threading.local()


def _restore_device() -> None:
    prev = getattr(_guard, "prev", None)
    if prev is not None:
        _guard.prev = None
        torch.cuda.set_device(prev)


def check(rc: int, what: str) -> None:
    _restore_device()
    if rc != 0:
        raise NativeError(f"{what} failed (status {rc}): {last_error()}")


def require_cuda(t: torch.Tensor, name: str = "tensor") -> None:
    if not t.is_cuda:
        raise NativeError(f"{name} must be a CUDA tensor: skeletondiffusion_b200 has no CPU path "
                          f"(got device {t.device})")


def stream_ptr(device: torch.device) -> int:
    """Stream handle of `device`'s current stream; makes `device` the current CUDA device until the matching check()."""
    device = torch.device(device)
    idx = device.index if device.index is not None else torch.cuda.current_device()
    cur = torch.cuda.current_device()
    if idx != cur:
        if getattr(_guard, "prev", None) is None:
            _guard.prev = cur
        torch.cuda.set_device(idx)
    return torch.cuda.current_stream(device).cuda_stream


def dptr(t: Optional[torch.Tensor]) -> Optional[int]:
    return None if t is None else t.data_ptr()


def view_of(t: Optional[torch.Tensor], rep: int = 1) -> SdView:
    """sd_view of a [B, N, W] fp32 tensor whose last dim is contiguous (arbitrary B/N strides)."""
    if t is None:
        return SdView(None, 0, 0, 1, 0)
    assert t.dim() == 3 and t.dtype == torch.float32 and (t.shape[-1] == 1 or t.stride(-1) == 1), \
        f"view_of expects fp32 [B,N,W] with unit inner stride, got {tuple(t.shape)} {t.stride()} {t.dtype}"
    return SdView(t.data_ptr(), t.stride(0), t.stride(1), int(rep), t.shape[-1])
