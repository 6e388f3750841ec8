#!/usr/bin/env python
"""Generate tests/golden/frenet.npz by RUNNING THE REFERENCE planner in this container.

    python tools/make_golden_frenet.py            # needs /root/reference

Executed, unmodified (paths relative to /root/reference/Simulation_testing/
Simulation_Data_Collection/Data_From_Carla): Agent/zzz/JunctionTrajectoryPlanner.py
(get_all_candidate_trajectories :103-121, get_optimal_trajectory :123-130,
get_trajectory_by_index :132-140, calc_frenet_paths :292-340, calc_global_paths :342-379,
check_paths :381-394), Agent/zzz/predict.py (:6-110), Agent/zzz/cubic_spline_planner.py,
Agent/zzz/tools.py (dense_polyline2d :69-92, get_frenet_state :235-270).

The planner is driven through its public entry points with a stub `dynamic_map` (plain
attribute objects: ego_vehicle, vehicles, lanes[0].central_path) — the only thing the CARLA side
supplies.  Routes: a straight lane, an arc, an S-bend and a short lane (so that candidates run
off its end and are truncated, :346-349).  Frozen per scenario: the planner's inputs (spline knots,
Frenet start state, ego speed, obstacle states) and everything it derives (candidate costs,
feasibility, collision flags, global paths, the selected index and trajectory).
"""
import os
import sys
import types

import numpy as np

REF = "/root/reference/Simulation_testing/Simulation_Data_Collection/Data_From_Carla"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
sys.path.insert(0, REF)

from Agent.zzz import JunctionTrajectoryPlanner as jtp   # noqa: E402
from Agent.zzz.tools import dense_polyline2d               # noqa: E402

NS = types.SimpleNamespace


def route(kind):
    if kind == 0:    # straight, 120 m, heading 0.3 rad
        t = np.linspace(0, 120, 25)
        return np.c_[10 + t * np.cos(0.3), -5 + t * np.sin(0.3)]
    if kind == 1:    # left arc, radius 40 m, quarter circle + run-in
        a = np.linspace(0, np.pi / 2, 30)
        arc = np.c_[40 * np.sin(a), 40 * (1 - np.cos(a))]
        run = np.c_[np.linspace(-30, -2, 8), np.zeros(8)]
        return np.vstack([run, arc])
    if kind == 2:    # S-bend
        x = np.linspace(0, 150, 40)
        return np.c_[x, 6 * np.sin(x / 25.0)]
    x = np.linspace(0, 34, 10)   # short lane: candidates leave it before 3.9 s
    return np.c_[x, 0.02 * x * x]


def make_map(rng, line, along, speed, lat, n_obs, near):
    """Ego at arc-length fraction `along` of the route, lateral offset `lat`; obstacles scattered
    ahead (near=True puts them on the candidate corridor so that some candidates collide)."""
    seg = np.diff(line, axis=0)
    L = np.r_[0, np.cumsum(np.hypot(seg[:, 0], seg[:, 1]))]
    s = along * L[-1]
    i = min(np.searchsorted(L, s, side="right") - 1, len(seg) - 1)
    u = seg[i] / np.hypot(*seg[i])
    nrm = np.array([-u[1], u[0]])
    p = line[i] + u * (s - L[i]) + nrm * lat
    yaw = np.arctan2(u[1], u[0]) + rng.normal(0, 0.05)
    ego = NS(x=float(p[0]), y=float(p[1]), v=float(speed), vx=float(speed * np.cos(yaw)),
             vy=float(speed * np.sin(yaw)), yaw=float(yaw))
    vehicles = []
    for _ in range(n_obs):
        if near:
            ahead = rng.uniform(3, 28)
            side = rng.uniform(-4.5, 4.5) - lat
            vv = rng.uniform(0, 1.5)
        else:
            ahead = rng.uniform(-20, 80)
            side = rng.uniform(-25, 25)
            vv = rng.uniform(0, 6)
        q = p + u * ahead + nrm * side
        yy = rng.uniform(-np.pi, np.pi)
        vehicles.append(NS(x=float(q[0]), y=float(q[1]), vx=float(vv * np.cos(yy)), vy=float(vv * np.sin(yy)),
                           yaw=float(yy)))
    path = [NS(position=NS(x=float(a), y=float(b))) for a, b in line]
    return NS(ego_vehicle=ego, vehicles=vehicles, lanes=[NS(central_path=path)], lanes_updated=False)


def run_one(dm):
    pl = jtp.JunctionTrajectoryPlanner()
    cands = pl.get_all_candidate_trajectories(dm)          # calls initialize -> build_frenet_path, predict
    assert cands is not None
    # start state as the planner derived it (calculate_start_state :248-276)
    st = pl.calculate_start_state(dm)
    nc = len(cands)
    nt = max(len(c[0].t) for c in cands)

    def ragged(key):
        out = np.full((nc, nt), np.nan)
        for i, c in enumerate(cands):
            v = getattr(c[0], key)
            out[i, :len(v)] = v
        return out
    out = dict(
        knots_x=np.asarray(pl.ref_path[:, 0]), knots_y=np.asarray(pl.ref_path[:, 1]),
        spline_s=np.asarray(pl.csp.s, np.float64),
        spline_cx=np.stack([pl.csp.sx.a, np.r_[pl.csp.sx.b, np.nan], pl.csp.sx.c, np.r_[pl.csp.sx.d, np.nan]]),
        spline_cy=np.stack([pl.csp.sy.a, np.r_[pl.csp.sy.b, np.nan], pl.csp.sy.c, np.r_[pl.csp.sy.d, np.nan]]),
        start=np.array([st.s0, st.c_d, st.c_d_d, st.c_d_dd, pl.c_speed]),
        ego_xy=np.array([dm.ego_vehicle.x, dm.ego_vehicle.y]),
        ego_v=np.array([dm.ego_vehicle.vx, dm.ego_vehicle.vy, dm.ego_vehicle.v]),
        lane_xy=np.array([[q.position.x, q.position.y] for q in dm.lanes[0].central_path]),
        obs=np.array([[v.x, v.y, v.vx, v.vy, v.yaw] for v in dm.vehicles]).reshape(-1, 5),
        cost=np.array([c[1] for c in cands]),
        n_valid=np.array([len(c[0].x) for c in cands]),
        x=np.full((nc, nt), np.nan), y=np.full((nc, nt), np.nan),
        s=ragged("s"), d=ragged("d"), s_d=ragged("s_d"), s_dd=ragged("s_dd"),
        n_t=np.array([len(c[0].t) for c in cands]),
        n_pred=len(pl.obs_prediction.predict_paths[0].t) if pl.obs_prediction.predict_paths else 0,
    )
    for i, c in enumerate(cands):
        out["x"][i, :len(c[0].x)] = c[0].x
        out["y"][i, :len(c[0].y)] = c[0].y
    ok = np.zeros(nc, bool)
    for fp, _, idx in pl.check_paths(list(cands)):
        ok[idx] = True
    out["feasible"] = ok
    out["collision_free"] = np.array([bool(pl.obs_prediction.check_collision(c[0])) for c in cands])
    index = pl.get_optimal_trajectory(cands)
    out["index"] = np.int64(index)
    traj = pl.get_trajectory_by_index(index, cands)         # index 0: lowest-cost candidate with zeroed speeds
    out["traj_xy"] = np.full((nt, 2), np.nan)
    out["traj_xy"][:len(traj.x), 0] = traj.x
    out["traj_xy"][:len(traj.x), 1] = traj.y
    out["traj_speed"] = np.full(nt, np.nan)
    out["traj_speed"][:len(traj.s_d)] = traj.s_d
    return out


ALT = dict(MAX_LEFT_WIDTH=-3, MAX_RIGHT_WIDTH=3, D_ROAD_W=1, MAXT=4.5, MAX_SPEED=11.0, MAX_ACCEL=2.0, MAX_CURVATURE=0.5,
           OBSTACLES_CONSIDERED=2, ROBOT_RADIUS=2, KD=3.0, KJ=0.3)


def main(alt=False):
    """alt=True: the same scenarios with other module constants (7 offsets x 2 horizons x 2 speeds,
    tighter limits, 2 obstacles considered, wider circles) -> frenet_alt.npz."""
    if alt:
        for k, v in ALT.items():
            setattr(jtp, k, v)
    rng = np.random.default_rng(20240611)
    cases = []
    for kind in range(4):
        line = route(kind)
        for j in range(16 if not alt else 10):
            near = j % 4 != 3
            n_obs = [3, 3, 0, 6, 1, 3, 4, 3, 2, 5, 3, 3, 6, 6, 4, 3][j]
            along = rng.uniform(0.02, 0.55 if kind != 3 else 0.75)
            speed = rng.uniform(0.5, 11.0) if j % 8 != 5 else rng.uniform(12.5, 17.0)   # some beyond MAX_SPEED
            dm = make_map(rng, line, along, speed, rng.uniform(-2.5, 2.5), n_obs, near)
            o = run_one(dm)
            o["route"] = np.int64(kind)
            cases.append(o)
    # a symmetric start (c_d = 0 on the straight lane, no lateral speed): exactly tied candidate costs
    line = route(0)
    dm = make_map(rng, line, 0.25, 6.0, 0.0, 0, False)
    dm.ego_vehicle.yaw = 0.3
    dm.ego_vehicle.vx, dm.ego_vehicle.vy = 6.0 * np.cos(0.3), 6.0 * np.sin(0.3)
    o = run_one(dm)
    o["route"] = np.int64(0)
    cases.append(o)

    n = len(cases)
    nk = max(len(c["knots_x"]) for c in cases)
    no = max(len(c["obs"]) for c in cases)
    pack = dict(n_cases=np.int64(n))

    def pad(key, width, fill=np.nan):
        arr = np.full((n, width) + cases[0][key].shape[1:], fill)
        for i, c in enumerate(cases):
            arr[i, :len(c[key])] = c[key]
        return arr
    pack["knots_x"], pack["knots_y"], pack["spline_s"] = pad("knots_x", nk), pad("knots_y", nk), pad("spline_s", nk)
    pack["n_knots"] = np.array([len(c["knots_x"]) for c in cases])
    pack["spline_cx"] = np.stack([np.pad(c["spline_cx"], ((0, 0), (0, nk - c["spline_cx"].shape[1])), constant_values=np.nan) for c in cases])
    pack["spline_cy"] = np.stack([np.pad(c["spline_cy"], ((0, 0), (0, nk - c["spline_cy"].shape[1])), constant_values=np.nan) for c in cases])
    pack["obs"] = pad("obs", no)
    pack["lane_xy"] = pad("lane_xy", max(len(c["lane_xy"]) for c in cases))
    pack["n_lane"] = np.array([len(c["lane_xy"]) for c in cases])
    pack["n_obs"] = np.array([len(c["obs"]) for c in cases])
    for k in ("start", "ego_xy", "ego_v", "cost", "n_valid", "n_t", "x", "y", "s", "d", "s_d", "s_dd", "feasible", "collision_free",
              "index", "traj_xy", "traj_speed", "route", "n_pred"):
        pack[k] = np.stack([np.asarray(c[k]) for c in cases])
    # the module constants the run used (JunctionTrajectoryPlanner.py:13-41)
    for k in ("MAX_SPEED", "MAX_ACCEL", "MAX_CURVATURE", "MAX_LEFT_WIDTH", "MAX_RIGHT_WIDTH", "D_ROAD_W", "DT", "MAXT",
              "MINT", "TARGET_SPEED", "D_T_S", "N_S_SAMPLE", "OBSTACLES_CONSIDERED", "ROBOT_RADIUS", "MOVE_GAP",
              "KJ", "KT", "KD", "KLAT", "KLON"):
        pack["const_" + k] = np.float64(getattr(jtp, k))
    np.savez_compressed(os.path.join(OUT, "frenet_alt.npz" if alt else "frenet.npz"), **pack)
    idx = pack["index"]
    print(f"{n} scenarios; selected index histogram {np.bincount(idx, minlength=11).tolist()}; "
          f"truncated candidates {(pack['n_valid'] < pack['n_t']).sum()}; "
          f"infeasible {(~pack['feasible']).sum()}; colliding {(~pack['collision_free']).sum()}; n_pred {set(pack['n_pred'].tolist())}")


if __name__ == "__main__":
    main(alt="--alt" in sys.argv)
