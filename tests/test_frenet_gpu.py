"""Parity of the CUDA Frenet planner (through the C ABI) with the reference's golden outputs and
with the oracle.  Selected index, feasibility / collision flags and truncation lengths are exact;
costs, paths and spline coefficients within 1e-9 (float64)."""
import types

import numpy as np
import pytest
import torch

import dcarl_b200 as dc
from dcarl_b200 import frenet as fr
from oracle import cport
from helpers_frenet import frenet_state, golden_batch, params_from_golden, synth_routes, synth_scenarios

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def g(golden):
    return golden("frenet")


def _routes_from(batch):
    off = batch["knot_off"]
    return fr.Routes([batch["knots"][off[i]:off[i + 1]] for i in range(len(off) - 1)])


def test_spline_build_matches_the_reference_coefficients(g):
    b = golden_batch(g)
    routes = _routes_from(b)
    for i in range(int(g["n_cases"])):
        coef = routes.coefficients(i)
        nk = int(g["n_knots"][i])
        assert np.abs(coef[:, 0] - g["spline_s"][i, :nk]).max() < 1e-12
        for col, ref in ((1, g["spline_cx"][i]), (5, g["spline_cy"][i])):
            for k in range(4):
                want = ref[k, :nk]
                m = ~np.isnan(want)
                assert np.abs(coef[m, col + k] - want[m]).max() < 1e-10


def test_plan_matches_the_reference_run(g):
    p = params_from_golden(g, fr.frenet_params)
    b = golden_batch(g)
    res = fr.frenet_plan(_routes_from(b), b["path_id"], b["start"], b["ego_xy"], b["obs"], b["n_obs"], p, want_paths=True)
    assert np.array_equal(res.index, g["index"])
    assert np.array_equal(res.n_valid, g["n_valid"])
    assert np.array_equal(res.feasible, g["feasible"])
    assert np.array_equal(res.collision_free, g["collision_free"])
    assert np.abs(res.cost - g["cost"]).max() / np.abs(g["cost"]).max() < 1e-12
    m = ~np.isnan(g["x"])
    assert np.array_equal(np.isnan(res.paths[..., 0]), ~m)
    assert np.abs(res.paths[..., 0][m] - g["x"][m]).max() < 1e-9
    assert np.abs(res.paths[..., 1][m] - g["y"][m]).max() < 1e-9
    tm = ~np.isnan(g["traj_xy"][..., 0])
    assert np.array_equal(res.traj_len, tm.sum(axis=1))
    assert np.abs(res.traj[..., 0][tm] - g["traj_xy"][..., 0][tm]).max() < 1e-9
    assert np.abs(res.traj[..., 1][tm] - g["traj_xy"][..., 1][tm]).max() < 1e-9
    assert np.abs(res.traj[..., 2] - g["traj_speed"]).max() < 1e-9


def test_plan_matches_the_reference_run_with_other_constants(golden):
    """tests/golden/frenet_alt.npz: the reference planner run with 28 candidates of 14 / 15 points, tighter
    limits, two obstacles considered (tools/make_golden_frenet.py --alt)."""
    a = golden("frenet_alt")
    p = params_from_golden(a, fr.frenet_params)
    b = golden_batch(a)
    res = fr.frenet_plan(_routes_from(b), b["path_id"], b["start"], b["ego_xy"], b["obs"], b["n_obs"], p, want_paths=True)
    assert np.array_equal(res.index, a["index"])
    assert np.array_equal(res.n_valid, a["n_valid"])
    assert np.array_equal(res.feasible, a["feasible"]) and np.array_equal(res.collision_free, a["collision_free"])
    assert np.abs(res.cost - a["cost"]).max() / np.abs(a["cost"]).max() < 1e-11
    m = ~np.isnan(a["x"])
    assert np.array_equal(np.isnan(res.paths[..., 0]), ~m)
    assert np.abs(res.paths[..., 0][m] - a["x"][m]).max() < 1e-8 and np.abs(res.paths[..., 1][m] - a["y"][m]).max() < 1e-8
    sm = ~np.isnan(a["traj_speed"])
    assert np.array_equal(np.isnan(res.traj[..., 2]), ~sm)
    assert np.abs(res.traj[..., 2][sm] - a["traj_speed"][sm]).max() < 1e-8


def _against_oracle(n, n_paths, seed, params_kw=None, max_obs=6):
    rng = np.random.default_rng(seed)
    knots, off = synth_routes(rng, n_paths)
    pid, start, ego, obs, n_obs = synth_scenarios(rng, n, knots, off, max_obs=max_obs)
    kw = params_kw or {}
    po, pg = cport.frenet_params(**kw), fr.frenet_params(**kw)
    coef = np.concatenate([cport.spline_build(knots[off[i]:off[i + 1]]) for i in range(n_paths)])
    ref = cport.frenet_plan(coef, off, pid, start, ego, obs, n_obs, po)
    routes = fr.Routes([knots[off[i]:off[i + 1]] for i in range(n_paths)])
    assert np.abs(routes.coef.cpu().numpy() - coef).max() < 1e-9
    res = fr.frenet_plan(routes, pid, start, ego, obs, n_obs, pg, want_paths=True)
    flags = res.feasible.astype(np.uint8) | (res.collision_free.astype(np.uint8) << 1)
    # a flag can only differ where a limit is met to within rounding; none is expected
    assert np.array_equal(flags, ref["flags"]), f"{(flags != ref['flags']).sum()} flag mismatches"
    assert np.array_equal(res.index, ref["index"])
    assert np.array_equal(res.n_valid, ref["n_valid"]) and np.array_equal(res.traj_len, ref["traj_len"])
    assert np.abs(res.cost - ref["cost"]).max() / np.abs(ref["cost"]).max() < 1e-11
    m = ~np.isnan(ref["paths"])
    assert np.array_equal(np.isnan(res.paths), ~m)
    assert np.abs(res.paths[m] - ref["paths"][m]).max() < 1e-8
    tm = ~np.isnan(ref["traj"])
    assert np.array_equal(np.isnan(res.traj), ~tm) and np.abs(res.traj[tm] - ref["traj"][tm]).max() < 1e-8
    return res


def test_random_scenarios_vs_oracle():
    res = _against_oracle(6000, 37, seed=11)
    hist = np.bincount(res.index, minlength=11)
    assert hist[0] > 0 and (hist[1:] > 0).sum() >= 4          # brake and several different candidates occur
    assert (~res.collision_free).any() and (~res.feasible).any() and (res.n_valid < 14).any()


def test_other_grids_and_limits_vs_oracle():
    # 7 offsets x 2 horizons x 2 speeds = 28 candidates, 14- and 15-point paths, tighter limits, 2 obstacles considered
    _against_oracle(1500, 9, seed=12, params_kw=dict(MAX_LEFT_WIDTH=-3, MAX_RIGHT_WIDTH=3, D_ROAD_W=1, MINT=4.0, MAXT=4.5,
                                                      MAX_SPEED=11.0, MAX_ACCEL=2.0, MAX_CURVATURE=0.5,
                                                      OBSTACLES_CONSIDERED=2, ROBOT_RADIUS=2))
    _against_oracle(700, 5, seed=13, params_kw=dict(N_S_SAMPLE=2, D_T_S=5 / 3.6, KD=3.0, KJ=0.3), max_obs=12)


def test_device_tensors_no_obstacles_and_bad_route_id():
    rng = np.random.default_rng(5)
    knots, off = synth_routes(rng, 3)
    routes = fr.Routes([knots[off[i]:off[i + 1]] for i in range(3)])
    pid, start, ego, obs, n_obs = synth_scenarios(rng, 64, knots, off, max_obs=4)
    dv = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
    res = fr.frenet_plan(routes, dv(pid, torch.int32), dv(start, torch.float64), dv(ego, torch.float64),
                         dv(obs, torch.float64), dv(np.zeros_like(n_obs), torch.int32))
    assert res.index.is_cuda and bool(res.collision_free.all())             # no obstacles: nothing collides (PRD:23-24)
    host = fr.frenet_plan(routes, pid, start, ego, obs, np.zeros_like(n_obs))
    assert np.array_equal(host.index, res.index.cpu().numpy())
    none = fr.frenet_plan(routes, pid, start, ego, np.zeros((64, 0, 5)), np.zeros(64, np.int32))   # no obstacle array at all
    assert np.array_equal(none.index, host.index) and none.collision_free.all()
    bad = pid.copy()
    bad[7] = 3
    with pytest.raises(IndexError):
        fr.frenet_plan(routes, bad, start, ego, obs, n_obs)
    empty = fr.frenet_plan(routes, np.zeros(0, np.int32), np.zeros((0, 5)), np.zeros((0, 2)), np.zeros((0, 1, 5)), np.zeros(0, np.int32))
    assert empty.index.shape == (0,)
    with pytest.raises(dc.DcarlError):
        fr.frenet_plan(routes, pid, start, ego, obs, n_obs, fr.frenet_params(D_ROAD_W=1, MAX_LEFT_WIDTH=-7, MAX_RIGHT_WIDTH=7, N_S_SAMPLE=2))   # 60 candidates


def test_start_state_matches_the_reference_and_the_numpy_restatement(g):
    b = golden_batch(g)
    routes = _routes_from(b)
    ego = np.concatenate([g["ego_xy"], g["ego_v"]], axis=1)                 # x, y, vx, vy, v
    start = fr.frenet_start(routes, b["path_id"], ego)
    assert np.abs(start - g["start"]).max() < 1e-9                          # calculate_start_state of the reference run
    # random egos around random routes, incl. beyond both ends of the route
    rng = np.random.default_rng(21)
    knots, off = synth_routes(rng, 12)
    routes = fr.Routes([knots[off[i]:off[i + 1]] for i in range(12)])
    n = 4000
    pid = rng.integers(0, 12, n).astype(np.int32)
    j = off[:-1][pid] + (rng.random(n) * (off[1:] - off[:-1])[pid]).astype(np.int64)
    ego = np.concatenate([knots[j] + rng.normal(0, 6, (n, 2)), rng.normal(0, 5, (n, 2)), rng.uniform(0, 15, (n, 1))], axis=1)
    got = fr.frenet_start(routes, pid, ego)
    for i in range(n):
        s, d, vd = frenet_state(ego[i, 0], ego[i, 1], ego[i, 2], ego[i, 3], knots[off[pid[i]]:off[pid[i] + 1]])
        assert np.allclose(got[i], [s, -d, vd, 0.0, ego[i, 4]], rtol=0, atol=1e-9), i
    assert np.isnan(fr.frenet_start(routes, [12], ego[:1])).all()          # route id out of range


def test_drop_in_planner_class_reproduces_the_reference_tick(g):
    NS = types.SimpleNamespace
    for i in (0, 5, 17, 21, 40, 52, 63):
        nl = int(g["n_lane"][i])
        lane = [NS(position=NS(x=float(x), y=float(y))) for x, y in g["lane_xy"][i, :nl]]
        no = int(g["n_obs"][i])
        veh = [NS(x=o[0], y=o[1], vx=o[2], vy=o[3], yaw=o[4]) for o in g["obs"][i, :no]]
        ego = NS(x=float(g["ego_xy"][i, 0]), y=float(g["ego_xy"][i, 1]), vx=float(g["ego_v"][i, 0]),
                 vy=float(g["ego_v"][i, 1]), v=float(g["ego_v"][i, 2]))
        dm = NS(ego_vehicle=ego, vehicles=veh, lanes=[NS(central_path=lane)], lanes_updated=False)
        planner = dc.JunctionTrajectoryPlanner()
        action, index = planner.trajectory_update(dm)
        assert index == int(g["index"][i])
        n = int((~np.isnan(g["traj_xy"][i, :, 0])).sum())
        assert action.trajectory.shape == (n, 2)
        assert np.abs(action.trajectory - g["traj_xy"][i, :n]).max() < 1e-8
        assert np.abs(np.array(action.desired_speed) - g["traj_speed"][i]).max() < 1e-8
        assert len(planner.all_trajectory) == g["cost"].shape[1]


# ---- candidate rollout + the closed collect -> buffer -> confidence loop ---------------------------------
def _rollout_kw(g, r, i=None):
    return dict(dt=float(g["const_DT"]), robot_radius=float(g["const_ROBOT_RADIUS"]), move_gap=float(g["const_MOVE_GAP"]),
                ticks_per_point=float(r["ticks_per_point"]), collision_reward=float(r["collision_reward"]),
                stuck_speed=float(r["stuck_speed"]), stay_thres=float(r["stay_thres"]), n_pred=15,
                obstacles_considered=int(g["const_OBSTACLES_CONSIDERED"]))


def test_candidate_rollout_matches_episodes_rolled_with_the_reference_objects(g, golden):
    """dcarl_rollout_candidates on the planner's own device outputs against tests/golden/rollout.npz (the
    reference's planner + predict.check_collision drove those episodes, tools/make_golden_rollout.py)."""
    import torch
    r = golden("rollout")
    n = int(r["n_cases"])
    p = params_from_golden(g, fr.frenet_params)
    b = golden_batch(g)
    routes = _routes_from(b)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
    for pass_y in np.unique(r["pass_y"]):                               # one launch per pass threshold (per route)
        m = np.nonzero(r["pass_y"] == pass_y)[0]
        d = (t(b["path_id"][m], torch.int32), t(b["start"][m], torch.float64), t(b["ego_xy"][m], torch.float64),
             t(b["obs"][m], torch.float64), t(b["n_obs"][m], torch.int32))
        plan = fr.frenet_plan(routes, *d, p, want_paths=True)
        ret, steps, term = dc.rollout_candidates(plan, d[2], d[3], d[4], dc.rollout_params(pass_y=float(pass_y), **_rollout_kw(g, r)))
        assert np.array_equal(steps.cpu().numpy(), r["steps"][m]) and np.array_equal(term.cpu().numpy(), r["term"][m])
        assert np.abs(ret.cpu().numpy() - r["ret"][m]).max() < 1e-9
    assert n == 64


def test_candidate_rollout_random_scenarios_vs_oracle():
    import torch
    from oracle import pyref
    rng = np.random.default_rng(31)
    knots, off = synth_routes(rng, 7)
    routes = fr.Routes([knots[off[i]:off[i + 1]] for i in range(7)])
    pid, start, ego, obs, n_obs = synth_scenarios(rng, 400, knots, off, max_obs=6)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", dt)
    d = (t(pid, torch.int32), t(start, torch.float64), t(ego, torch.float64), t(obs, torch.float64), t(n_obs, torch.int32))
    plan = fr.frenet_plan(routes, *d, want_paths=True)
    pass_y = float(np.median(ego[:, 1]))
    rp = dc.rollout_params(pass_y=pass_y, stay_thres=0.5)
    ret, steps, term = (x.cpu().numpy() for x in dc.rollout_candidates(plan, d[2], d[3], d[4], rp))
    paths, nv, cost = plan.paths.cpu().numpy(), plan.n_valid.cpu().numpy(), plan.cost.cpu().numpy()
    for i in range(400):
        want = pyref.rollout_candidates(paths[i], nv[i], cost[i], ego[i], [tuple(v) for v in obs[i, :n_obs[i]]], considered=4,
                                        move_gap=1.0, n_pred=15, dt=0.3, radius=1.0, pass_y=pass_y, stay_thres=0.5)
        assert [w[1] for w in want] == steps[i].tolist() and [w[2] for w in want] == term[i].tolist(), i
        assert np.abs(np.array([w[0] for w in want]) - ret[i]).max() < 1e-9
    assert set(np.unique(term)) >= {0, 1, 2}


def test_collect_and_select_closes_the_loop_on_the_device():
    """planner -> candidate rollout -> buffer rows -> box-kernel confidence -> act_test, one Python call."""
    import torch
    from oracle import cport
    rng = np.random.default_rng(41)
    knots, off = synth_routes(rng, 3)
    routes = fr.Routes([knots[off[i]:off[i + 1]] for i in range(3)])
    # 12 distinct scenarios, each repeated 30 times with small jitter so that the boxes of a scenario overlap
    pid, start, ego_xy, obs, n_obs = synth_scenarios(rng, 12, knots, off, max_obs=3)
    rep = 30
    pid, n_obs = np.repeat(pid, rep), np.repeat(n_obs, rep)
    obs = np.repeat(obs, rep, axis=0)
    obs[:, :, :2] += rng.normal(0, 0.05, (len(pid), obs.shape[1], 2))
    u = rng.uniform(-np.pi, np.pi, 12)
    v = rng.uniform(3, 9, 12)
    ego = np.repeat(np.c_[ego_xy, v * np.cos(u), v * np.sin(u), v], rep, axis=0)
    ego[:, :2] += rng.normal(0, 0.05, (len(pid), 2))
    buf = dc.BoxBuffer(20)
    rls = dc.RlsParams(visited_times_thres=10)
    out = dc.collect_and_select(routes, pid, ego, obs, n_obs, buf, rollout=dc.rollout_params(pass_y=float("nan")), rls_params=rls)
    n, A = len(pid), out["ret"].shape[1]
    assert A == 11 and len(buf) == out["rows_appended"] == n * A
    st, ret = buf.state.cpu().numpy(), buf.ret.cpu().numpy()
    # the rows are (state (+) action, episode return), scenario-major (dqn_value_collect.py:128-137)
    assert np.array_equal(st[:, 20].reshape(n, A), np.tile(np.arange(A, dtype=np.float32), (n, 1)))
    assert np.array_equal(ret.reshape(n, A), out["ret"].cpu().numpy().astype(np.float32))
    assert np.array_equal(st[::A, :20], out["state"].cpu().numpy().astype(np.float32))
    # the confidence query of every scenario's state against the oracle on the same rows
    ref = cport.box_raw(st, ret, dc.VISITED_STATE_DIST[0], out["state"].cpu().numpy(), A)
    assert np.array_equal(out["n"].cpu().numpy(), ref[:, :, 0]) and ref[:, 0, 0].min() >= 10
    nn = np.maximum(ref[:, :, 0], 1)
    nmv = np.stack([ref[:, :, 0], np.where(ref[:, :, 0] > 0, ref[:, :, 1] / nn, -1.0),
                    np.where(ref[:, :, 0] > 0, np.maximum(ref[:, :, 2] / nn - (ref[:, :, 1] / nn) ** 2, 0), -1.0)], 2)
    assert np.array_equal(out["action"].cpu().numpy(), cport.box_select(nmv, visited_times_thres=10))
    assert np.abs(out["mean"].cpu().numpy() - nmv[:, :, 1]).max() < 1e-9
    # the same call answering the queries from a sorted index (one launch): same counts and picks
    buf2 = dc.BoxBuffer(20)
    out2 = dc.collect_and_select(routes, pid, ego, obs, n_obs, buf2, rollout=dc.rollout_params(pass_y=float("nan")),
                                 rls_params=rls, index=True)
    assert np.array_equal(out2["n"].cpu().numpy(), out["n"].cpu().numpy())
    assert np.array_equal(out2["action"].cpu().numpy(), out["action"].cpu().numpy())
    assert np.abs(out2["mean"].cpu().numpy() - out["mean"].cpu().numpy()).max() < 1e-12


def test_drop_in_planner_hands_over_the_previous_trajectory_on_brake_ticks(golden):
    """Consecutive ticks against the reference driven the same way (tools/make_golden_frenet_ticks.py): a brake
    tick returns the PREVIOUS tick's trajectory with zeroed speeds (JTP:142-152) unless the ego is slower than
    1 m/s or there is no previous trajectory."""
    t = golden("frenet_ticks")
    NS = types.SimpleNamespace
    brake_ticks = 0
    for e in range(int(t["n_episodes"])):
        lane = [NS(position=NS(x=float(x), y=float(y))) for x, y in t[f"e{e}_lane"]]
        planner = dc.JunctionTrajectoryPlanner()
        for k in range(int(t[f"e{e}_n"])):
            eg = t[f"e{e}_t{k}_ego"]
            ego = NS(x=float(eg[0]), y=float(eg[1]), vx=float(eg[2]), vy=float(eg[3]), v=float(eg[4]), yaw=float(eg[5]))
            veh = [NS(x=o[0], y=o[1], vx=o[2], vy=o[3], yaw=o[4]) for o in t[f"e{e}_t{k}_obs"]]
            dm = NS(ego_vehicle=ego, vehicles=veh, lanes=[NS(central_path=lane)], lanes_updated=False)
            action, index = planner.trajectory_update(dm)
            want_xy, want_v = t[f"e{e}_t{k}_traj"], t[f"e{e}_t{k}_speed"]
            assert index == int(t[f"e{e}_t{k}_index"]), (e, k)
            assert action.trajectory.shape == want_xy.shape, (e, k)
            assert np.abs(action.trajectory - want_xy).max() < 1e-8, (e, k)
            assert len(action.desired_speed) == len(want_v) and np.abs(np.array(action.desired_speed) - want_v).max() < 1e-8
            brake_ticks += index == 0
    assert brake_ticks >= 5
