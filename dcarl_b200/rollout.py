"""Constant-velocity rollout surrogate — see csrc/rollout.cu for the cited formulas."""
import numpy as np
import torch

from . import _abi, _dev

# ego-speed multipliers standing in for the 11 CARLA candidate trajectories (action 0 = brake / rule)
DEFAULT_SPEED_SCALE = np.array([0.0, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1, 1.2, 1.3, 1.4], np.float32)


def rollout_cv(state0, action, steps=200, dt=0.05, radius=1.0, pass_x=30.0, speed_scale=DEFAULT_SPEED_SCALE):
    """state0 float32 [n, 20], action int32 [n] -> (return float32, steps int32, term int32)."""
    _dev.require_cuda()
    on_device = _dev.is_device(state0)
    s0 = _dev.as_device(state0, torch.float32).reshape(-1, 20).contiguous()
    a = _dev.as_device(action, torch.int32).reshape(-1)
    ss = _dev.as_device(speed_scale, torch.float32).reshape(-1)
    n = s0.shape[0]
    if a.numel() != n:
        raise ValueError("one action per scenario expected")
    ret, st, tm = _dev.empty((n,), torch.float32), _dev.empty((n,), torch.int32), _dev.empty((n,), torch.int32)
    _abi.check(_abi.lib().dcarl_rollout_cv(_dev.ptr(s0), _dev.ptr(a), n, int(steps), float(dt), float(radius),
                                           float(pass_x), _dev.ptr(ss), ss.numel(), _dev.ptr(ret), _dev.ptr(st),
                                           _dev.ptr(tm), _dev.stream_handle()))
    if on_device:
        return ret, st, tm
    tmh = tm.cpu().numpy()
    if (tmh < 0).any():
        raise IndexError("action outside [0, len(speed_scale))")
    return ret.cpu().numpy(), st.cpu().numpy(), tmh
