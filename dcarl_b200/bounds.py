"""The four bound functions of the reference, on the GPU.

Drop-in for ``upper_bound / lower_bound / CI_lower_bound / mean_value`` of
Simulation_testing/Simulation_1/test_DCARL.py:10-28 (same copy in Simulation_2/test_DCARL.py:9-27):
same names, same positional/keyword arguments (``loc`` is accepted and unused, as in the
reference), a Python float back.  ``bounds_batch`` evaluates many sample arrays in one launch.
"""
import numpy as np
import torch

from . import _abi, _dev


def bounds_batch(arrays=None, *, flat=None, offsets=None, alpha=0.05, scale=150, cap=100.0):
    """Evaluate all four bounds for many sample arrays in one kernel launch.

    Either ``arrays`` (a list of 1-D array-likes) or ``flat`` + ``offsets`` (CSR layout;
    ``flat`` may be a CUDA tensor).  Returns float64 [n, 4] = (upper, lower, CI lower, mean_value),
    a CUDA tensor if ``flat`` was one, else a numpy array."""
    _dev.require_cuda()
    on_device = _dev.is_device(flat)
    if arrays is not None:
        arrays = [np.asarray(a, dtype=np.float64 if np.asarray(a).dtype != np.float32 else np.float32).ravel()
                  for a in arrays]
        lens = np.array([a.size for a in arrays], np.int64)
        offsets = np.concatenate([[0], np.cumsum(lens)]).astype(np.int64)
        dt = np.float32 if arrays and all(a.dtype == np.float32 for a in arrays) else np.float64
        flat = np.concatenate([a.astype(dt) for a in arrays]) if arrays else np.zeros(0, dt)
    fdt = _dev.float_dtype_of(flat)
    x = _dev.as_device(flat, fdt)
    off = _dev.as_device(offsets, torch.int64)
    n_seg = off.numel() - 1
    out = _dev.empty((max(n_seg, 0), 4), torch.float64)
    p = _abi.Params(alpha=alpha, scale=scale, cap=cap)
    _abi.check(_abi.lib().dcarl_bounds(_dev.ptr(x), int(fdt == torch.float64), _dev.ptr(off), n_seg, p,
                                       _dev.ptr(out), _dev.stream_handle()))
    return out if on_device else out.cpu().numpy()


def _one(data_array, alpha, scale, col):
    a = data_array if _dev.is_device(data_array) else np.asarray(data_array)
    if _dev.is_device(a):
        off = torch.tensor([0, a.numel()], dtype=torch.int64, device=a.device)
        return float(bounds_batch(flat=a.reshape(-1), offsets=off, alpha=alpha, scale=scale)[0, col])
    return float(bounds_batch([a], alpha=alpha, scale=scale)[0, col])


def upper_bound(data_array, alpha=0.05, loc=-50, scale=150):
    """min(100, mean + scale*sqrt(ln(1/alpha)/2/n)) — test_DCARL.py:10-12."""
    return _one(data_array, alpha, scale, 0)


def lower_bound(data_array, alpha=0.05, loc=-50, scale=150):
    """mean - scale*sqrt(ln(1/alpha)/2/n) — test_DCARL.py:14-16."""
    return _one(data_array, alpha, scale, 1)


def CI_lower_bound(data_array, alpha=0.05, loc=-50, scale=150):
    """sum/n/(n+1) - 4*std/(n+1) + sum/(n+1) - scale*sqrt(ln(1/alpha)/2/(n+1)) — test_DCARL.py:18-24."""
    return _one(data_array, alpha, scale, 2)


def mean_value(data_array, alpha=0.05, loc=-50, scale=150):
    """min(100, mean) — test_DCARL.py:26-28."""
    return _one(data_array, alpha, scale, 3)


def cell_values(n, total, total_sq, act, *, alpha=0.05, scale=150, cap=100.0, rule_act=0):
    """The bound of cells given their moments (count, sum, sum of squares) and action, evaluated by
    the same device code the stream kernel uses: upper bound for ``act == rule_act``, else
    min(lower, CI lower).  numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    _dev.require_cuda()
    on_device = _dev.is_device(total)
    dn = _dev.as_device(n, torch.int32).reshape(-1)
    d1 = _dev.as_device(total, torch.float64).reshape(-1)
    d2 = _dev.as_device(total_sq, torch.float64).reshape(-1)
    da = _dev.as_device(act, torch.int32).reshape(-1)
    cnt = dn.numel()
    if not (d1.numel() == d2.numel() == da.numel() == cnt):
        raise ValueError("n, total, total_sq and act must have the same length")
    out = _dev.empty((cnt,), torch.float64)
    p = _abi.Params(alpha=alpha, scale=scale, cap=cap, rule_act=rule_act)
    _abi.check(_abi.lib().dcarl_cell_values(_dev.ptr(dn), _dev.ptr(d1), _dev.ptr(d2), _dev.ptr(da), cnt, p,
                                            _dev.ptr(out), _dev.stream_handle()))
    return out if on_device else out.cpu().numpy()
