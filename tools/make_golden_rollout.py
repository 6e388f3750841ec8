#!/usr/bin/env python
"""Generate tests/golden/rollout.npz by driving THE REFERENCE's planner and collision predicate here.

    python tools/make_golden_rollout.py            # needs /root/reference

The scenarios are those of tests/golden/frenet.npz (same generator, same seed: tools/make_golden_frenet.py).
For every scenario the unmodified reference planner (Agent/zzz/JunctionTrajectoryPlanner.py) produces the
candidate trajectories, and for every action a = 0 .. C (dqn_value_collect.py:144-145; 0 = the brake
trajectory of get_trajectory_by_index, JTP:132-153) one episode is rolled along that trajectory with the
rules of Test_Scenarios/TestScenario_Town03.py:403-422 and :182-201.  The collision time comes from the
reference's own `predict.check_collision` (predict.py:21-60), called on growing prefixes of the trajectory:
the first prefix it rejects ends the episode at the sampled step it added.  CARLA's vehicle dynamics are
replaced by "the ego is at point k at time k*DT" (speed = point spacing / DT).
"""
import copy
import math
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_frenet as mg          # noqa: E402  (also puts the reference on sys.path)

jtp = mg.jtp
PASS_Y = {0: 8.0, 1: 2.0, 2: -3.0, 3: 1.5}      # one threshold per route so that some episodes "pass"
TICKS, STUCK_V, STAY, R_COLL = 6.0, 0.1, 2.0, -100.0


def first_hit(pred, fp):
    """Index of the sampled step at which predict.check_collision first rejects a prefix of fp, or None."""
    n = len(fp.x)
    if pred.check_collision(fp):
        return None
    for m in range(2, n + 1):
        pre = copy.copy(fp)
        pre.x, pre.y, pre.t = fp.x[:m], fp.y[:m], fp.t[:m]
        if not pred.check_collision(pre):
            # the step the last extension added to range(2, min(m - 1, len(path.t) - 1), 2)
            lim = min(m - 1, len(pred.predict_paths[0].t) - 1)
            return max(range(2, lim, 2))
    raise AssertionError("whole trajectory rejected but no prefix is")


def episode(fp, brake, pred, pass_y):
    n = len(fp.x)
    if brake:
        fp = copy.copy(fp)
        fp.x, fp.y = [fp.x[0]] * n, [fp.y[0]] * n          # desired speeds are all zero (JTP:147-150): no motion
    hit = first_hit(pred, fp)
    total, clock = 0.0, 0.0
    for k in range(n):
        k0 = k if k + 1 < n else k - 1
        v = math.hypot(fp.x[k0 + 1] - fp.x[k0], fp.y[k0 + 1] - fp.y[k0]) / jtp.DT if n >= 2 else 0.0
        reward, done, term = TICKS * math.sqrt(v) * 0.1, False, 0
        if hit is not None and k == hit:
            reward, done, term = R_COLL, True, 1
        if fp.y[k] < pass_y:
            done, term = True, term or 2
        elif k > 0 and v < STUCK_V:
            if k * jtp.DT - clock > STAY:
                reward, done, term = 0.0, True, 3
        else:
            clock = k * jtp.DT
        total += reward
        if done:
            return total, k + 1, term
    return total, n, 0


def main():
    rng = np.random.default_rng(20240611)
    rows = []
    for kind in range(4):
        line = mg.route(kind)
        for j in range(16):
            near = j % 4 != 3
            n_obs = [3, 3, 0, 6, 1, 3, 4, 3, 2, 5, 3, 3, 6, 6, 4, 3][j]
            along = rng.uniform(0.02, 0.55 if kind != 3 else 0.75)
            speed = rng.uniform(0.5, 11.0) if j % 8 != 5 else rng.uniform(12.5, 17.0)
            dm = mg.make_map(rng, line, along, speed, rng.uniform(-2.5, 2.5), n_obs, near)
            pl = jtp.JunctionTrajectoryPlanner()
            cands = pl.get_all_candidate_trajectories(dm)
            index = pl.get_optimal_trajectory(cands)
            free = [bool(pl.obs_prediction.check_collision(c[0])) for c in cands]
            C = len(cands)
            out = np.zeros((C + 1, 3))
            for act in range(C + 1):
                fp = copy.deepcopy(pl.get_trajectory_by_index(act, copy.deepcopy(cands)))
                out[act] = episode(fp, act == 0, pl.obs_prediction, PASS_Y[kind])
            # a candidate the planner calls collision-free never collides in its episode, and vice versa
            for c in range(C):
                assert (out[c + 1, 2] == 1) <= (not free[c])
            ego = dm.ego_vehicle
            state20 = [ego.x, ego.y, ego.vx, ego.vy, ego.yaw]
            for v in dm.vehicles[:3]:
                state20 += [v.x, v.y, v.vx, v.vy, v.yaw]
            state20 += [0.0] * (20 - len(state20))
            rows.append(dict(ret=out[:, 0], steps=out[:, 1].astype(np.int32), term=out[:, 2].astype(np.int32),
                             index=index, pass_y=PASS_Y[kind], state20=np.array(state20)))
    g = np.load(os.path.join(mg.OUT, "frenet.npz"))
    n = len(rows)
    assert np.array_equal(g["index"][:n], [r["index"] for r in rows]), "scenario order differs from frenet.npz"
    np.savez_compressed(os.path.join(mg.OUT, "rollout.npz"),
                        ret=np.stack([r["ret"] for r in rows]), steps=np.stack([r["steps"] for r in rows]),
                        term=np.stack([r["term"] for r in rows]), pass_y=np.array([r["pass_y"] for r in rows]),
                        state20=np.stack([r["state20"] for r in rows]), n_cases=np.int64(n),
                        ticks_per_point=TICKS, stuck_speed=STUCK_V, stay_thres=STAY, collision_reward=R_COLL)
    term = np.stack([r["term"] for r in rows])
    print(f"{n} scenarios x {term.shape[1]} actions; term histogram (ran out, collision, pass, stuck) = "
          f"{np.bincount(term.ravel(), minlength=4).tolist()}")


if __name__ == "__main__":
    main()
