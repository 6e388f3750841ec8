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


# ---- candidate-trajectory rollout and the closed collect -> buffer -> confidence loop ------------------------
def rollout_params(**kw):
    """dcarl_rollout_params with the reference's constants (JTP:19-31, TestScenario_Town03.py:84, 182-198,
    403-410); keyword overrides by field name."""
    p = _abi.RolloutParams()
    _abi.lib().dcarl_default_rollout_params(p)
    for k, v in kw.items():
        if not hasattr(p, k):
            raise TypeError(f"unknown rollout parameter {k!r}")
        setattr(p, k, v)
    return p


def rollout_candidates(plan, ego_xy, obs, n_obs, params=None, stream=None):
    """The episode return of every action of every scenario (dcarl_rollout_candidates).

    plan: a FrenetResult of ``frenet_plan(..., want_paths=True)`` on CUDA tensors (its paths / n_valid / cost
    are consumed in place, nothing returns to the host); ego_xy [n, 2], obs [n, max_obs, 5], n_obs [n]: the
    planner's own inputs.  Returns CUDA tensors (ret float64 [n, C+1], steps int32 [n, C+1], term int32
    [n, C+1]); action 0 is the brake trajectory, action c+1 candidate c (dqn_value_collect.py:92-101, 144-145)."""
    _dev.require_cuda()
    if plan.paths is None:
        raise ValueError("rollout_candidates needs the candidate paths: call frenet_plan(..., want_paths=True)")
    paths = plan.paths
    dv = paths.device
    n, C, Tm, _ = paths.shape
    p = params or rollout_params()
    d_ego = _dev.as_device(ego_xy, torch.float64, dv).reshape(n, 2)
    d_nobs = _dev.as_device(n_obs, torch.int32, dv).reshape(n)
    d_obs = _dev.as_device(obs, torch.float64, dv)
    max_obs = int(d_obs.shape[1]) if d_obs.dim() == 3 else 0
    d_obs = d_obs.reshape(n, max_obs, 5).contiguous() if max_obs else None
    n_valid = plan.n_valid.to(torch.int32).contiguous()
    cost = plan.cost.to(torch.float64).contiguous()
    ret = torch.empty((n, C + 1), dtype=torch.float64, device=dv)
    steps = torch.empty((n, C + 1), dtype=torch.int32, device=dv)
    term = torch.empty((n, C + 1), dtype=torch.int32, device=dv)
    with torch.cuda.device(dv):
        _abi.check(_abi.lib().dcarl_rollout_candidates(
            _dev.ptr(paths.contiguous()), _dev.ptr(n_valid), _dev.ptr(cost), _dev.ptr(d_ego), _dev.ptr(d_obs),
            _dev.ptr(d_nobs), max_obs, n, C, Tm, p, _dev.ptr(ret), _dev.ptr(steps), _dev.ptr(term),
            _dev.stream_handle(stream)))
    return ret, steps, term


def collect_and_select(routes, path_id, ego, obs, n_obs, buf, *, frenet=None, rollout=None, rls_params=None,
                       state20=None, index=False):
    """The collector's loop for a batch of scenarios, on the device from end to end
    (dqn_value_collect.py:86-145 with TestScenario_Town03.py:403-422 as the reward / termination rules):

      planner tick        frenet_start + frenet_plan                 candidates of every scenario
      candidate rollout   rollout_candidates                          one episode return per action
      record              buf.append(state (+) action, return)        (state, used_action, episode_reward) rows
      confidence + pick   buf.query(state)                            box-kernel moments, act_test

    ego [n, 5] = {x, y, vx, vy, v}; obs [n, max_obs, 5] = {x, y, vx, vy, yaw}; buf a BoxBuffer whose D is the
    length of the recorded state (20: ego x, y, vx, vy, yaw + three vehicles, TestScenario_Town03.py:231-253;
    built from `ego` / `obs` unless `state20` is given).  Returns a dict of CUDA tensors: plan (FrenetResult),
    ret / steps / term [n, C+1], rows_appended, and the act_test answer for every scenario's state after the
    new rows are in the buffer: n / mean / var [n, C+1] and action [n].  index=True sorts the buffer by its most
    selective observation dim after the append (BoxBuffer.build_index: the rows change order) and answers the n
    queries from their index ranges in one launch instead of n / 64 passes over the whole buffer."""
    from .frenet import frenet_plan, frenet_start
    dv = routes.device
    d_ego = _dev.as_device(ego, torch.float64, dv).reshape(-1, 5)
    n = d_ego.shape[0]
    d_obs = _dev.as_device(obs, torch.float64, dv)
    d_nobs = _dev.as_device(n_obs, torch.int32, dv).reshape(n)
    d_pid = _dev.as_device(path_id, torch.int32, dv).reshape(n)
    start = frenet_start(routes, d_pid, d_ego)
    ego_xy = d_ego[:, :2].contiguous()
    plan = frenet_plan(routes, d_pid, start, ego_xy, d_obs, d_nobs, frenet, want_paths=True)
    ret, steps, term = rollout_candidates(plan, ego_xy, d_obs, d_nobs, rollout)
    A = ret.shape[1]
    if state20 is None:
        if buf.D != 20:
            raise ValueError("the default recorded state has 20 entries: pass state20 for another buffer width")
        yaw = torch.atan2(d_ego[:, 3], d_ego[:, 2])
        st = torch.zeros((n, 20), dtype=torch.float64, device=dv)
        st[:, 0:4] = d_ego[:, 0:4]
        st[:, 4] = yaw
        k = min(3, d_obs.shape[1]) if d_obs.dim() == 3 else 0
        if k:
            present = (torch.arange(k, device=dv)[None, :] < d_nobs[:, None]).to(torch.float64)
            st[:, 5:5 + 5 * k] = (d_obs[:, :k, :] * present[:, :, None]).reshape(n, 5 * k)
    else:
        st = _dev.as_device(state20, torch.float64, dv).reshape(n, buf.D)
    # one recorded row per (scenario, action): state (+) action -> episode return   (DVC:128-137)
    rows = torch.cat([st[:, None, :].expand(n, A, buf.D),
                      torch.arange(A, dtype=torch.float64, device=dv)[None, :, None].expand(n, A, 1)], 2)
    buf.append(rows.reshape(n * A, buf.D + 1), ret.reshape(n * A))
    if index:
        buf.build_index()
        mom = buf.moments_indexed(st, A)
    else:
        mom = buf.moments(st, A)
    mv, pick = buf.select(mom, rls_params)
    return dict(plan=plan, ret=ret, steps=steps, term=term, rows_appended=n * A, state=st,
                n=mom[:, :, 0], mean=mv[:, :, 0], var=mv[:, :, 1], action=pick)
