"""Glue shared by the drop-in modules: puts the `hop` package on sys.path, converts between the
reference's list-of-ndarray conventions and batched CUDA tensors, and maps per-instance status words
to the reference's exception types.  No numerical work happens here."""
from __future__ import annotations

import ctypes as C
import os
import sys

import numpy as np

_PKG = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _PKG not in sys.path:
    sys.path.insert(0, _PKG)

import torch  # noqa: E402

from hop import _cabi, api, cases  # noqa: E402


def device():
    _cabi.require_device()
    return torch.device("cuda", torch.cuda.current_device())


def dev(a):
    return torch.as_tensor(np.ascontiguousarray(np.asarray(a, dtype=np.float64)), device=device())


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def require_dynamics(F):
    """The B200 path runs dynamics as device functions: F must come from systems.make_* (hop.cases.Dynamics)."""
    if not hasattr(F, "hop_sys"):
        raise TypeError("F must be a device-registered dynamics object returned by systems.make_* "
                        "(an arbitrary Python closure cannot run on the GPU; see INTEGRATION.md)")
    return F


def raise_status(st: int, what: str):
    code = int(st) & 0xFF
    if code == 1:
        raise FloatingPointError(f"Non-finite values in {what}")
    if code == 2:
        raise np.linalg.LinAlgError(f"{what} failed: matrix not PD / singular after the jitter ladder")


def stack(lst):
    return np.stack([np.asarray(a, dtype=np.float64) for a in lst])
