"""Frenet candidate trajectories + collision-checked selection on the GPU.

Mirror of the reference's junction planner, Simulation_testing/Simulation_Data_Collection/
Data_From_Carla/Agent/zzz/JunctionTrajectoryPlanner.py (JTP), with predict.py (PRD),
cubic_spline_planner.py (SPL) and the two helpers of tools.py it needs:

    frenet_params(**constants)            the module constants JTP:13-39 -> the planner's sample grids
    Routes(polylines)                     build_frenet_path / generate_target_course (JTP:79-90, 279-290)
    frenet_start(routes, path_id, ego)                           calculate_start_state (JTP:248-276) for n scenarios
    frenet_plan(routes, path_id, start, ego_xy, obs, n_obs)      one planning tick for n scenarios
    JunctionTrajectoryPlanner             drop-in class: trajectory_update(dynamic_map) -> (TrajectoryAction, index)

The batched call is the product path: one warp per scenario (csrc/frenet.cu).  numpy in -> numpy out,
CUDA tensors in -> CUDA tensors out.  No CPU fallback: without the built library or a device it raises.
"""
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch

from . import _abi, _dev
from ._abi import FrenetParams


def frenet_params(MAX_SPEED=50.0 / 3.6, MAX_ACCEL=10.0, MAX_CURVATURE=500.0, MAX_LEFT_WIDTH=-4, MAX_RIGHT_WIDTH=4,
                  D_ROAD_W=2, DT=0.3, MAXT=4.2, MINT=4.0, TARGET_SPEED=30.0 / 3.6, D_T_S=15 / 3.6, N_S_SAMPLE=1,
                  OBSTACLES_CONSIDERED=4, ROBOT_RADIUS=1, MOVE_GAP=1, KJ=0.1, KT=0.1, KD=1.0, KLAT=1.0, KLON=1.0):
    """Keyword names and defaults are the module constants of JTP:13-39.  The loops of
    calc_frenet_paths (JTP:300, 303, 308, 315) and predict (PRD:91) run over np.arange grids; the
    same calls produce the grids here, so their lengths (e.g. len(np.arange(0, 4.2, 0.3)) == 15)
    are the reference's."""
    p = FrenetParams()
    d = np.arange(MAX_LEFT_WIDTH, MAX_RIGHT_WIDTH + 1, D_ROAD_W).astype(np.float64)
    T = np.arange(MINT, MAXT, DT).astype(np.float64)
    tv = np.arange(TARGET_SPEED - D_T_S * N_S_SAMPLE, TARGET_SPEED + D_T_S * N_S_SAMPLE, D_T_S).astype(np.float64)
    if not (1 <= len(d) <= _abi.FRENET_MAX_D and 1 <= len(T) <= _abi.FRENET_MAX_T and 1 <= len(tv) <= _abi.FRENET_MAX_V):
        raise ValueError("frenet_params: sample grids outside the supported sizes (16 offsets, 4 horizons, 8 speeds)")
    p.d_samples[:len(d)] = d.tolist()
    p.T_samples[:len(T)] = T.tolist()
    p.tv_samples[:len(tv)] = tv.tolist()
    p.n_t[:len(T)] = [len(np.arange(0.0, Ti, DT)) for Ti in T]
    p.n_d, p.n_T, p.n_v = len(d), len(T), len(tv)
    p.n_pred = len(np.arange(0.0, MAXT, DT))
    p.obstacles_considered = int(OBSTACLES_CONSIDERED)
    p.dt, p.target_speed = float(DT), float(TARGET_SPEED)
    p.max_speed, p.max_accel, p.max_curvature = float(MAX_SPEED), float(MAX_ACCEL), float(MAX_CURVATURE)
    p.robot_radius, p.move_gap = float(ROBOT_RADIUS), float(MOVE_GAP)
    p.kj, p.kt, p.kd, p.klat, p.klon = float(KJ), float(KT), float(KD), float(KLAT), float(KLON)
    return p


def dense_polyline2d(line, resolution):
    """tools.py:69-92 — route ingest: resample a lane centre line to ~`resolution` metres (host, once per route)."""
    line = np.asarray(line, np.float64)
    if line is None or len(line) == 0:
        raise ValueError("Line input is null")
    s = np.concatenate([[0], np.cumsum(np.linalg.norm(np.diff(line, axis=0), axis=1))])
    num = int(round(s[-1] / resolution))
    s_space = np.linspace(0, s[-1], num=num)
    return np.array([np.interp(s_space, s, line[:, 0]), np.interp(s_space, s, line[:, 1])]).T


class Routes:
    """P reference paths as cubic splines on the device (SPL:123-178 via dcarl_spline_build).

    polylines: list of [n_i, 2] knot arrays (what the reference feeds Spline2D: the densified lane,
    JTP:82-89), or with `resolution` raw centre lines that are densified first."""

    def __init__(self, polylines, resolution=None, device=None):
        _dev.require_cuda()
        lines = [np.ascontiguousarray(np.asarray(l, np.float64).reshape(-1, 2)) for l in polylines]
        if resolution is not None:
            lines = [dense_polyline2d(l, resolution) for l in lines]
        if not lines or any(len(l) < 2 for l in lines):
            raise ValueError("Routes: every route needs at least two knots")
        self.n_paths = len(lines)
        self.knot_off_host = np.r_[0, np.cumsum([len(l) for l in lines])].astype(np.int64)
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.knots = torch.from_numpy(np.ascontiguousarray(np.concatenate(lines))).to(self.device).contiguous()
        self.knot_off = torch.from_numpy(self.knot_off_host).to(self.device)
        self.coef = torch.empty((self.knots.shape[0], 9), dtype=torch.float64, device=self.device)
        _abi.check(_abi.lib().dcarl_spline_build(_dev.ptr(self.knots), _dev.ptr(self.knot_off), self.n_paths,
                                                 _dev.ptr(self.coef), _dev.stream_handle()))

    def coefficients(self, r):
        """Host copy of route r's rows {s, ax, bx, cx, dx, ay, by, cy, dy}."""
        return self.coef[self.knot_off_host[r]:self.knot_off_host[r + 1]].cpu().numpy()


@dataclass
class FrenetResult:
    index: object           # [n] int32: 0 = brake, c+1 = candidate c                      (JTP:123-130)
    cost: object            # [n, C] cf                                                     (JTP:333)
    feasible: object        # [n, C] bool, check_paths                                      (JTP:381-394)
    collision_free: object  # [n, C] bool, predict.check_collision                          (PRD:21-60)
    n_valid: object         # [n, C] points before the path leaves the route                (JTP:347-349)
    traj: object            # [n, Tm, 3] x, y, desired speed of the selected trajectory     (JTP:154-156)
    traj_len: object        # [n]
    paths: Optional[object] = None   # [n, C, Tm, 2] global paths, NaN past n_valid


def frenet_plan(routes, path_id, start, ego_xy, obs, n_obs, params=None, *, want_paths=False, stream=None):
    """One planning tick for n scenarios (dcarl_frenet_plan; layouts in include/dcarl_b200.h).

    path_id [n]; start [n, 5] = {s0, c_d, c_d_d, c_d_dd, c_speed}; ego_xy [n, 2]; obs [n, max_obs, 5] =
    {x, y, vx, vy, yaw}; n_obs [n].  Raises IndexError if a path_id is outside the routes."""
    _dev.require_cuda()
    p = params or frenet_params()
    on_device = _dev.is_device(start)
    dv = routes.device
    d_pid = _dev.as_device(path_id, torch.int32, dv).reshape(-1)
    n = int(d_pid.shape[0])
    d_start = _dev.as_device(start, torch.float64, dv).reshape(n, 5)
    d_ego = _dev.as_device(ego_xy, torch.float64, dv).reshape(n, 2)
    d_nobs = _dev.as_device(n_obs, torch.int32, dv).reshape(n)
    d_obs = _dev.as_device(obs, torch.float64, dv)
    max_obs = int(d_obs.numel() // (5 * n)) if n and d_obs.numel() else 0
    d_obs = d_obs.reshape(n, max_obs, 5) if max_obs else None
    C = p.n_d * p.n_T * p.n_v
    Tm = max(p.n_t[:p.n_T])
    e = lambda shape, dt: torch.empty(shape, dtype=dt, device=dv)
    index, cost, flags = e((n,), torch.int32), e((n, C), torch.float64), e((n, C), torch.uint8)
    n_valid, traj, traj_len = e((n, C), torch.int32), e((n, Tm, 3), torch.float64), e((n,), torch.int32)
    paths = e((n, C, Tm, 2), torch.float64) if want_paths else None
    _abi.check(_abi.lib().dcarl_frenet_plan(
        _dev.ptr(routes.coef), _dev.ptr(routes.knot_off), routes.n_paths, _dev.ptr(d_pid), _dev.ptr(d_start),
        _dev.ptr(d_ego), _dev.ptr(d_obs), _dev.ptr(d_nobs), max_obs, n, p, _dev.ptr(index), _dev.ptr(cost),
        _dev.ptr(flags), _dev.ptr(n_valid), _dev.ptr(paths), _dev.ptr(traj), _dev.ptr(traj_len),
        _dev.stream_handle(stream)))
    res = FrenetResult(index, cost, (flags & 1) != 0, (flags & 2) != 0, n_valid, traj, traj_len, paths)
    if not on_device:
        for k, v in vars(res).items():
            if v is not None:
                setattr(res, k, v.cpu().numpy())
        if n and (res.index < 0).any():
            raise IndexError("frenet_plan: path_id outside [0, n_paths)")
    return res


# ---- drop-in single-scenario class -----------------------------------------------------------------
class TrajectoryAction(object):
    """actions.py:7-11."""

    def __init__(self, trajectory=None, desired_speed=None):
        self.trajectory = trajectory
        self.desired_speed = desired_speed


def frenet_start(routes, path_id, ego, *, stream=None):
    """calculate_start_state (JTP:248-276) for n scenarios: ego [n, 5] = {x, y, vx, vy, v} projected on
    route path_id[i] (tools.get_frenet_state, tools.py:138-270) -> start [n, 5] = {s0, c_d, c_d_d, 0, c_speed},
    the `start` input of frenet_plan.  numpy in -> numpy out, CUDA tensors in -> CUDA tensor out."""
    _dev.require_cuda()
    on_device = _dev.is_device(ego)
    dv = routes.device
    d_pid = _dev.as_device(path_id, torch.int32, dv).reshape(-1)
    n = int(d_pid.shape[0])
    d_ego = _dev.as_device(ego, torch.float64, dv).reshape(n, 5)
    out = torch.empty((n, 5), dtype=torch.float64, device=dv)
    _abi.check(_abi.lib().dcarl_frenet_start(_dev.ptr(routes.coef), _dev.ptr(routes.knot_off), routes.n_paths,
                                             _dev.ptr(d_pid), _dev.ptr(d_ego), n, _dev.ptr(out), _dev.stream_handle(stream)))
    return out if on_device else out.cpu().numpy()


class JunctionTrajectoryPlanner(object):
    """Drop-in for JTP:44-290 as the scripts use it (`trajectory_update(dynamic_map)`), one scenario per
    call through the batched kernel.  dynamic_map: .ego_vehicle (x, y, v, vx, vy), .vehicles (x, y, vx,
    vy, yaw each), .lanes[0].central_path (points with .position.x/.y), .lanes_updated.

    Stateful like the reference: the trajectory returned by a tick is kept (update_last_trajectory,
    JTP:158-161) and a later brake tick (index 0) hands it out again with zeroed speeds whenever it has more
    than five speed samples and the ego is faster than 1 m/s (get_backup_trajectory, JTP:142-152); only the
    first tick, a slow ego, or a tick after clear_buff() brake on the cheapest candidate of the tick itself."""

    def __init__(self, target_line=0, params=None):
        self.params = params or frenet_params()
        self.target_line = target_line
        self.routes = None
        self.ref_path = None
        self.all_trajectory = None
        self.last_result = None
        self.c_speed = 0.0
        self.last_trajectory_rule = None          # (xy [n, 2], s_d list) of the previous tick's trajectory

    def clear_buff(self, clean_csp=True):
        if self.routes is None:                   # JTP:66-67
            return
        self.last_trajectory_rule = None          # JTP:69-72
        if clean_csp:
            self.routes = None
            self.ref_path = None

    def build_frenet_path(self, dynamic_map, clean_current_csp=False):
        if self.routes is None or clean_current_csp:
            pts = np.array([(pt.position.x, pt.position.y) for pt in dynamic_map.lanes[0].central_path])
            self.ref_path = dense_polyline2d(pts, 2)
            self.routes = Routes([self.ref_path])

    def get_all_candidate_trajectories(self, dynamic_map):
        if self.routes is None or getattr(dynamic_map, "lanes_updated", False):
            self.build_frenet_path(dynamic_map, True)
        ego = dynamic_map.ego_vehicle
        start = frenet_start(self.routes, [0], [[ego.x, ego.y, ego.vx, ego.vy, ego.v]])   # JTP:248-276
        self.c_speed = float(start[0][4])
        veh = list(dynamic_map.vehicles)
        obs = np.array([[v.x, v.y, v.vx, v.vy, v.yaw] for v in veh], np.float64).reshape(1, -1, 5)
        if obs.shape[1] > 32:                                                   # the kernel ranks at most 32
            keep = np.argsort(np.hypot(obs[0, :, 0] - ego.x, obs[0, :, 1] - ego.y), kind="stable")[:32]
            obs = obs[:, np.sort(keep)]
        res = frenet_plan(self.routes, [0], start, [[ego.x, ego.y]], obs if obs.shape[1] else np.zeros((1, 1, 5)),
                          [obs.shape[1]], self.params, want_paths=True)
        self.last_result = res
        self.all_trajectory = [(res.paths[0, c, :res.n_valid[0, c]], float(res.cost[0, c]), c)
                               for c in range(res.cost.shape[1])]
        return self.all_trajectory

    def get_optimal_trajectory(self, path_tuples=None):
        return int(self.last_result.index[0])

    def get_trajectory_by_index(self, act_idx, all_trajectory=None):
        """JTP:132-152 for the tick's selected index: (xy [n, 2], desired speeds); 0 = the backup trajectory."""
        res = self.last_result
        if int(act_idx) != int(res.index[0]):
            raise ValueError("the drop-in class hands out the selected trajectory only; the batched frenet_plan / "
                             "rollout_candidates cover the other candidates")
        if int(act_idx) == 0:
            last = self.last_trajectory_rule
            if last is not None and len(last[1]) > 5 and self.c_speed > 1:          # JTP:144-146
                return last[0], [0] * len(last[1])
        # the kernel's out_traj: the selected candidate, or for index 0 the cheapest one with zeroed speeds (JTP:148-150)
        n = int(res.traj_len[0])
        k = int(np.sum(~np.isnan(res.traj[0, :, 2])))
        return res.traj[0, :n, :2].copy(), res.traj[0, :k, 2].tolist()

    def update_last_trajectory(self, trajectory):                                     # JTP:158-161
        self.last_trajectory_rule = (np.asarray(trajectory[0], np.float64).copy(), list(trajectory[1]))

    def trajectory_update(self, dynamic_map):                                         # JTP:88-101
        if self.get_all_candidate_trajectories(dynamic_map) is None:
            return None
        index = int(self.last_result.index[0])
        final = self.get_trajectory_by_index(index)
        self.update_last_trajectory(final)
        return TrajectoryAction(final[0], final[1]), index
