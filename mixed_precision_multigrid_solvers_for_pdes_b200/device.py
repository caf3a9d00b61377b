"""Device-memory plumbing: pitched (nx, ny) fields held in torch CUDA tensors.

PyTorch is used for allocation, streams and torch.distributed only; all arithmetic on the hot
path happens in libmgb200 kernels.  A field is a 2-D tensor view with strides (ld, 1) whose row
pitch ld is a multiple of 32 elements, so every row starts 128-byte aligned (the vector/TMA
kernels need 16-byte aligned rows; n = 2^k + 1 is odd, hence the padding).
"""
from __future__ import annotations

from typing import Tuple, Union

import numpy as np
import torch

from . import _lib

ArrayLike = Union[np.ndarray, torch.Tensor]

_TORCH_DT = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64}
_NP_DT = {torch.float32: np.float32, torch.float64: np.float64}
PITCH_ALIGN = 32  # elements


def torch_dtype(dt) -> torch.dtype:
    if isinstance(dt, torch.dtype):
        if dt not in _NP_DT:
            raise TypeError(f"unsupported dtype {dt}")
        return dt
    try:
        return _TORCH_DT[np.dtype(dt)]
    except (KeyError, TypeError):
        raise TypeError(f"unsupported dtype {dt!r}: only float32/float64 fields exist on this path")


def np_dtype(dt):
    return _NP_DT[torch_dtype(dt)]


def code(dt) -> int:
    return _lib.F64 if torch_dtype(dt) == torch.float64 else _lib.F32


def require_cuda(device=None) -> torch.device:
    if not torch.cuda.is_available():
        raise _lib.MGLibraryError(
            "no CUDA device: this package runs its hot path only on the GPU (sm_100a); there is no CPU fallback")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    if dev.type != "cuda":
        raise _lib.MGLibraryError(f"device {dev} is not a CUDA device; there is no CPU fallback")
    return dev


def pitch_for(ny: int) -> int:
    return (ny + PITCH_ALIGN - 1) // PITCH_ALIGN * PITCH_ALIGN


def empty_field(nx: int, ny: int, dtype, device=None, zero: bool = True, pad_rows: int = 0) -> torch.Tensor:
    """(nx, ny) view with strides (ld, 1) into a fresh pitched buffer.  `pad_rows` extra zero rows
    are allocated after the field (scratch some kernels may read past the end of)."""
    dev = require_cuda(device)
    ld = pitch_for(ny)
    alloc = torch.zeros if zero else torch.empty
    buf = alloc((nx + pad_rows) * ld, dtype=torch_dtype(dtype), device=dev)
    return buf.view(nx + pad_rows, ld)[:nx, :ny]


def ld(t: torch.Tensor) -> int:
    if t.dim() != 2 or (t.shape[1] > 1 and t.stride(1) != 1):
        raise ValueError("field tensors must be 2-D with unit stride along the last axis")
    return t.stride(0) if t.shape[0] > 1 else max(t.stride(0), t.shape[1])


def stream_ptr() -> int:
    return torch.cuda.current_stream().cuda_stream


def to_device(x: ArrayLike, device=None, dtype=None) -> Tuple[torch.Tensor, bool]:
    """Return (pitched CUDA field, came_from_numpy).  NumPy input is copied into a pitched buffer;
    CUDA tensors with unit inner stride are used as they are (no copy)."""
    if isinstance(x, torch.Tensor):
        if x.dim() != 2:
            raise ValueError("fields must be 2-D")
        if not x.is_cuda:
            t = empty_field(x.shape[0], x.shape[1], dtype or x.dtype, device, zero=False)
            t.copy_(x)
            return t, False
        if dtype is not None and torch_dtype(dtype) != x.dtype:
            t = empty_field(x.shape[0], x.shape[1], dtype, x.device, zero=False)
            t.copy_(x)
            return t, False
        if x.shape[1] > 1 and x.stride(1) != 1:
            t = empty_field(x.shape[0], x.shape[1], x.dtype, x.device, zero=False)
            t.copy_(x)
            return t, False
        return x, False
    a = np.asarray(x)
    if a.ndim != 2:
        raise ValueError("fields must be 2-D")
    tdt = torch_dtype(dtype if dtype is not None else a.dtype)
    t = empty_field(a.shape[0], a.shape[1], tdt, device, zero=False)
    t.copy_(torch.from_numpy(np.ascontiguousarray(a)))
    return t, True


def to_host(t: torch.Tensor) -> np.ndarray:
    return t.detach().cpu().numpy().copy()


def like_input(result: torch.Tensor, was_numpy: bool):
    return to_host(result) if was_numpy else result
