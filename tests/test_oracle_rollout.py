"""The candidate-rollout oracle (oracle/pyref.py, written from the reference's lines) against episodes rolled
with the reference's own planner objects and predict.check_collision (tests/golden/rollout.npz, made by
tools/make_golden_rollout.py on the scenarios of tests/golden/frenet.npz).  CPU only."""
import numpy as np

from oracle import pyref


def _scenario(g, i):
    C = g["cost"].shape[1]
    nv = g["n_valid"][i].astype(int)
    paths = np.stack([g["x"][i], g["y"][i]], axis=2)                   # [C, Tm, 2], NaN past n_valid
    no = int(g["n_obs"][i])
    return C, nv, paths, g["cost"][i], g["ego_xy"][i], [tuple(v) for v in g["obs"][i, :no]]


def test_rollout_oracle_matches_episodes_rolled_with_the_reference_objects(golden):
    g, r = golden("frenet"), golden("rollout")
    n = int(r["n_cases"])
    assert r["ret"].shape == (n, g["cost"].shape[1] + 1)
    seen = set()
    for i in range(n):
        C, nv, paths, cost, ego_xy, veh = _scenario(g, i)
        out = pyref.rollout_candidates(paths, nv, cost, ego_xy, veh, considered=int(g["const_OBSTACLES_CONSIDERED"]),
                                       move_gap=float(g["const_MOVE_GAP"]), n_pred=int(g["n_pred"][i]) or 15,
                                       dt=float(g["const_DT"]), radius=float(g["const_ROBOT_RADIUS"]),
                                       ticks_per_point=float(r["ticks_per_point"]), collision_reward=float(r["collision_reward"]),
                                       pass_y=float(r["pass_y"][i]), stuck_speed=float(r["stuck_speed"]),
                                       stay_thres=float(r["stay_thres"]))
        ret = np.array([o[0] for o in out])
        assert np.array_equal([o[1] for o in out], r["steps"][i]), i
        assert np.array_equal([o[2] for o in out], r["term"][i]), i
        assert np.abs(ret - r["ret"][i]).max() < 1e-12, i
        seen |= set(r["term"][i].tolist())
        # the planner's collision flags and the episodes agree: a collision-free candidate never collides
        assert not ((r["term"][i, 1:] == 1) & g["collision_free"][i]).any()
    assert seen == {0, 1, 2, 3}                                        # ran out, collision, pass and stuck all occur


def test_rollout_rules_in_isolation():
    # straight trajectory at 5 m/s, no vehicles: 14 points, the trajectory runs out
    xs, ys = [1.5 * k for k in range(14)], [100.0] * 14
    ret, steps, term = pyref.rollout_episode(xs, ys, False, [], 15)
    assert (steps, term) == (14, 0) and abs(ret - 14 * 6 * 0.1 * np.sqrt(5.0)) < 1e-12
    # braking: speed 0, the stuck clock is armed at the first step and fires once 2 s have passed (k = 7: 2.1 s)
    ret, steps, term = pyref.rollout_episode(xs, ys, True, [], 15)
    assert (ret, steps, term) == (0.0, 8, 3)
    # a vehicle parked on point 4: only even sampled steps 2, 4, ... are tested (predict.py:52-55)
    hit4 = [(xs[4], ys[4], 0.0, 0.0)]
    ret, steps, term = pyref.rollout_episode(xs, ys, False, hit4, 15)
    assert (steps, term) == (5, 1) and abs(ret - (4 * 6 * 0.1 * np.sqrt(5.0) - 100.0)) < 1e-12
    hit3 = [(xs[3], ys[3], 0.0, 0.0)]                                  # an odd step is never looked at
    assert pyref.rollout_episode(xs, ys, False, hit3, 15)[2] == 0
    hit12 = [(xs[12], ys[12], 0.0, 0.0)]                               # nor the last sampled steps: t < min(14 - 1, 15 - 1)
    assert pyref.rollout_episode(xs, ys, False, hit12, 15)[2] == 1
    assert pyref.rollout_episode(xs[:13], ys[:13], False, hit12, 15)[2] == 0
    # pass: the step reward stands, the episode ends (TestScenario_Town03.py:414-415)
    yy = [100.0 - 10 * k for k in range(14)]
    ret, steps, term = pyref.rollout_episode(xs, yy, False, [], 15, pass_y=73.7)
    assert (steps, term) == (4, 2)
    v = np.hypot(1.5, 10.0) / 0.3
    assert abs(ret - 4 * 6 * 0.1 * np.sqrt(v)) < 1e-12
