"""The {state, action, return} sampler of the reference, on the GPU.

Drop-in for Simulation_testing/Simulation_Data_Collection/Data_Sampling/data_sampling.py:
``Data_Generation()`` (:30-67), ``add_an_act_data`` (:5-9), ``random_state_norm`` (:12-17) and the
unused ``random_state_manual`` (:19-28).  The reference draws its deviates one call at a time from
numpy's / Python's global generators; here the deviates are drawn in bulk (from the same global
generators, in the same order, so a seeded run reproduces the reference's files bit for bit) and
the state binning, rejection and return construction run in one kernel.
"""
import os
import random

import numpy as np
import torch

from . import _abi, _dev


def sample_returns(q_star, z_state, act, z_ret, *, loc=3.0, sd=1.0, span=6.0, sigma=50.0, out_dtype=None, out=None):
    """One draw per element: idx = floor((loc+sd*z_state)/span*S) (-1 when rejected),
    R = q_star[idx, act] + sigma*z_ret.  numpy in -> numpy out, CUDA tensors in -> CUDA tensors out.
    out=(idx int32 [n], R [n]) reuses caller-held CUDA tensors (the call is 30 us of kernel for 1.3e7 draws:
    two allocations per call are a visible share of it)."""
    _dev.require_cuda()
    on_device = _dev.is_device(z_state)
    q_dt = _dev.float_dtype_of(q_star)
    z_dt = _dev.float_dtype_of(z_state)
    o_dt = out_dtype or z_dt
    q = _dev.as_device(q_star, q_dt)
    S, A = q.shape
    zs, zr = _dev.as_device(z_state, z_dt).reshape(-1), _dev.as_device(z_ret, z_dt).reshape(-1)
    a = _dev.as_device(act, torch.int32).reshape(-1)
    n = zs.numel()
    if zr.numel() != n or a.numel() != n:
        raise ValueError("z_state, z_ret and act must have the same length")
    if out is not None:
        idx, R = out
        o_dt = R.dtype
        if not (idx.is_cuda and R.is_cuda and idx.dtype == torch.int32 and idx.numel() == n and R.numel() == n
                and R.dtype in (torch.float32, torch.float64)):
            raise ValueError("out must be (int32 [n], float32/float64 [n]) CUDA tensors")
    else:
        idx = _dev.empty((n,), torch.int32)
        R = _dev.empty((n,), o_dt)
    _abi.check(_abi.lib().dcarl_sample_returns(
        _dev.ptr(q), int(q_dt == torch.float64), _dev.ptr(zs), _dev.ptr(zr), int(z_dt == torch.float64),
        _dev.ptr(a), n, S, A, loc, sd, span, sigma, _dev.ptr(idx), _dev.ptr(R), int(o_dt == torch.float64),
        _dev.stream_handle()))
    return (idx, R) if on_device else (idx.cpu().numpy(), R.cpu().numpy())


def sample_returns_philox(q_star, n, seed, *, loc=3.0, sd=1.0, span=6.0, sigma=50.0, out_dtype=torch.float32, out=None):
    """Throughput mode: deviates and actions from on-device Philox4x32-10.  Returns CUDA tensors
    (idx int32 with -1 = rejected, act int32, R); out=(idx, act, R) reuses caller-held ones."""
    _dev.require_cuda()
    q_dt = _dev.float_dtype_of(q_star, default=torch.float32 if _dev.is_device(q_star) else torch.float64)
    q = _dev.as_device(q_star, q_dt)
    S, A = q.shape
    if out is not None:
        idx, act, R = out
        out_dtype = R.dtype
    else:
        idx = _dev.empty((n,), torch.int32)
        act = _dev.empty((n,), torch.int32)
        R = _dev.empty((n,), out_dtype)
    _abi.check(_abi.lib().dcarl_sample_returns_philox(
        _dev.ptr(q), int(q_dt == torch.float64), int(seed) & (2 ** 64 - 1), n, S, A, loc, sd, span, sigma,
        _dev.ptr(idx), _dev.ptr(act), _dev.ptr(R), int(out_dtype == torch.float64), _dev.stream_handle()))
    return idx, act, R


def add_an_act_data(act, action_value):
    """One return draw N(action_value[act], 50^2) — data_sampling.py:5-9."""
    z = np.random.standard_normal(1)
    _, R = sample_returns(np.asarray(action_value, np.float64)[None, :], np.array([0.0]), [act], z,
                          loc=0.0, sd=1.0, span=1.0)   # one state: every draw lands in bin 0
    return float(R[0])


def random_state_norm(state_num, size):
    """floor(N(3,1)/6*state_num) — data_sampling.py:12-17 (out-of-range indices are kept, as there)."""
    z = np.random.standard_normal(size)
    # binning only: a one-action table of zeros; rejected draws are re-derived on the host below
    idx, _ = sample_returns(np.zeros((state_num, 1)), z, np.zeros(size, np.int32), np.zeros(size))
    out = idx.astype(int)
    rej = out < 0
    if rej.any():   # the reference returns the raw out-of-range bin; the kernel flags it as -1
        out[rej] = np.floor((3 + 1 * z[rej]) / 6 * state_num).astype(int)
    return out


def random_state_manual(state_num, size):
    """data_sampling.py:19-28 (unused by the reference's own Data_Generation; host RNG bookkeeping only)."""
    return [random.randint(1, state_num - 1) if random.random() > 0.1 else 0 for _ in range(size)]


def generate(state_num=20, data_size=50000, action_num=11, min_value=-50, max_value=100):
    """Data_Generation() returning arrays (data [n_valid, 4], action_values [S, A], states [S])."""
    states = np.random.uniform(0.0, 1.0, state_num)                                   # DS:39
    action_values = np.stack([min_value + (max_value - min_value) * np.random.uniform(0.0, 1.0, action_num)
                              for _ in range(state_num)])                             # DS:42-43
    z_state = np.random.standard_normal(data_size)                                    # DS:14
    # the reference consumes one randint and one normal per ACCEPTED draw, in order (DS:50-55)
    bins = np.floor((3 + 1 * z_state) / 6 * state_num)
    n_valid = int(((bins >= 0) & (bins < state_num)).sum())
    acts_v = np.fromiter((random.randint(0, action_num - 1) for _ in range(n_valid)), np.int32, n_valid)
    z_ret_v = np.random.standard_normal(n_valid) if n_valid else np.zeros(0)
    valid = (bins >= 0) & (bins < state_num)
    acts = np.zeros(data_size, np.int32)
    z_ret = np.zeros(data_size)
    acts[valid] = acts_v
    z_ret[valid] = z_ret_v
    idx, R = sample_returns(action_values, z_state, acts, z_ret)
    keep = idx >= 0
    if not np.array_equal(keep, valid):
        raise RuntimeError("device binning disagrees with the host acceptance mask")
    data = np.stack([idx[keep].astype(np.float64), states[idx[keep]], acts[keep].astype(np.float64), R[keep]], axis=1)
    return data, action_values, states


def Data_Generation(out_dir="Simulation_testing/Simulation_Data_Collection"):
    """data_sampling.py:30-67: draws the Sim-2 dataset and saves data.npy / action_value.npy /
    states.npy under ``out_dir`` (the reference's relative path by default)."""
    data, action_values, states = generate()
    os.makedirs(out_dir, exist_ok=True)
    np.save(os.path.join(out_dir, "data.npy"), data)                                  # DS:65-67
    np.save(os.path.join(out_dir, "action_value.npy"), action_values)
    np.save(os.path.join(out_dir, "states.npy"), states)
