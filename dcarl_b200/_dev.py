"""Device-buffer plumbing.  PyTorch is used only to hold device memory and streams."""
import ctypes as C

import numpy as np
import torch

from . import _abi

_NP2T = {np.dtype(np.float32): torch.float32, np.dtype(np.float64): torch.float64,
         np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64,
         np.dtype(np.uint8): torch.uint8}


def require_cuda():
    if not torch.cuda.is_available():
        raise RuntimeError("dcarl_b200: no CUDA device visible; the package has no CPU fallback")
    _abi.lib()


def stream_handle(stream=None):
    s = stream if stream is not None else torch.cuda.current_stream()
    return C.c_void_p(s.cuda_stream)


def ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def is_device(x):
    return isinstance(x, torch.Tensor) and x.is_cuda


def as_device(x, dtype, device=None):
    """Array-like / tensor -> contiguous CUDA tensor of `dtype` (torch dtype)."""
    if isinstance(x, torch.Tensor):
        return x.to(device=device or "cuda", dtype=dtype).contiguous()
    npdt = {v: k for k, v in _NP2T.items()}[dtype]
    a = np.ascontiguousarray(x, dtype=npdt)
    return torch.from_numpy(a).to(device or "cuda")


def float_dtype_of(x, default=torch.float64):
    """float32 stays float32, everything else computes in float64 (the reference is float64)."""
    if isinstance(x, torch.Tensor):
        return torch.float32 if x.dtype == torch.float32 else default
    if isinstance(x, np.ndarray) and x.dtype == np.float32:
        return torch.float32
    return default


def empty(shape, dtype, device=None):
    return torch.empty(shape, dtype=dtype, device=device or "cuda")
