#!/usr/bin/env python
"""Generate tests/golden/frenet_ticks.npz: the reference planner driven over CONSECUTIVE ticks through
`trajectory_update(dynamic_map)` (JunctionTrajectoryPlanner.py:88-101), so that the hand-over of the previous
tick's trajectory on a brake tick (get_backup_trajectory, JTP:142-152; update_last_trajectory, JTP:158-161) is
pinned: episodes where every candidate is blocked on some tick after a free one, at speeds above and below
1 m/s, and one that brakes on its very first tick.

    python tools/make_golden_frenet_ticks.py            # needs /root/reference
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_frenet as mg          # noqa: E402

jtp = mg.jtp
NS = types.SimpleNamespace


def wall(ego):
    """Stopped vehicles on the lane where every candidate still is at the first sampled steps t = 2, 4, 6 of
    predict.check_collision (the candidates only fan out later; 4 vehicles are considered, JTP:28)."""
    u = np.array([np.cos(ego.yaw), np.sin(ego.yaw)])
    out = []
    for t in (2, 4, 6):
        q = np.array([ego.x, ego.y]) + u * max(ego.v, 1.0) * 0.3 * t
        out.append(NS(x=float(q[0]), y=float(q[1]), vx=0.0, vy=0.0, yaw=float(ego.yaw)))
    return out


def episode(line, speeds, blocked):
    """One planner instance, len(speeds) ticks: the ego advances along the lane at the given speed, the wall
    is present on the ticks listed in `blocked`."""
    pl = jtp.JunctionTrajectoryPlanner()
    seg = np.diff(line, axis=0)
    L = np.r_[0, np.cumsum(np.hypot(seg[:, 0], seg[:, 1]))]
    s = 0.15 * L[-1]
    ticks = []
    for k, v in enumerate(speeds):
        i = min(np.searchsorted(L, s, side="right") - 1, len(seg) - 1)
        u = seg[i] / np.hypot(*seg[i])
        p = line[i] + u * (s - L[i])
        yaw = float(np.arctan2(u[1], u[0]))
        ego = NS(x=float(p[0]), y=float(p[1]), v=float(v), vx=float(v * np.cos(yaw)), vy=float(v * np.sin(yaw)), yaw=yaw)
        veh = wall(ego) if k in blocked else []
        path = [NS(position=NS(x=float(a), y=float(b))) for a, b in line]
        dm = NS(ego_vehicle=ego, vehicles=veh, lanes=[NS(central_path=path)], lanes_updated=False)
        act, index = pl.trajectory_update(dm)
        ticks.append(dict(ego=[ego.x, ego.y, ego.vx, ego.vy, ego.v, ego.yaw],
                          obs=np.array([[q.x, q.y, q.vx, q.vy, q.yaw] for q in veh]).reshape(-1, 5),
                          index=index, traj=np.asarray(act.trajectory, float), speed=np.asarray(act.desired_speed, float)))
        s += v * 0.3
    return ticks


def main():
    eps = []
    for kind, speeds, blocked in ((0, [6.0, 6.0, 6.0, 5.0], {1, 2}),      # brake on ticks 1, 2 after a free tick: last trajectory reused twice
                                  (2, [4.0, 0.8, 3.0], {1}),              # c_speed <= 1 on the brake tick: cheapest candidate instead
                                  (1, [7.0, 7.0], {0}),                   # brake on the very first tick: no previous trajectory
                                  (0, [5.0, 5.0, 5.0], {1})):             # free, brake, free again
        eps.append((kind, episode(mg.route(kind), speeds, blocked)))
    pack = dict(n_episodes=np.int64(len(eps)))
    for e, (kind, ticks) in enumerate(eps):
        pack[f"e{e}_route"] = np.int64(kind)
        pack[f"e{e}_lane"] = mg.route(kind)
        pack[f"e{e}_n"] = np.int64(len(ticks))
        for k, t in enumerate(ticks):
            for key in ("ego", "obs", "index", "traj", "speed"):
                pack[f"e{e}_t{k}_{key}"] = np.asarray(t[key])
    np.savez_compressed(os.path.join(mg.OUT, "frenet_ticks.npz"), **pack)
    print([[int(t["index"]) for t in ticks] for _, ticks in eps])


if __name__ == "__main__":
    main()
