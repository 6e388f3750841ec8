"""Shared helpers of the Frenet-planner tests: golden fixture -> batched planner inputs, and a
synthetic scenario generator for sizes the golden file does not cover."""
import numpy as np


def params_from_golden(g, factory):
    """Planner parameters from the module constants the reference run used (frozen in the fixture)."""
    kw = {k[len("const_"):]: float(g[k]) for k in g.files if k.startswith("const_")}
    for k in ("MAX_LEFT_WIDTH", "MAX_RIGHT_WIDTH", "D_ROAD_W", "N_S_SAMPLE", "OBSTACLES_CONSIDERED", "ROBOT_RADIUS", "MOVE_GAP"):
        kw[k] = int(kw[k])
    return factory(**kw)


def golden_batch(g, spline_build=None):
    """All golden scenarios as one batch; every scenario carries its own route (one spline each)."""
    n = int(g["n_cases"])
    nk = g["n_knots"].astype(np.int64)
    knot_off = np.r_[0, np.cumsum(nk)]
    knots = np.concatenate([np.stack([g["knots_x"][i, :nk[i]], g["knots_y"][i, :nk[i]]], axis=1) for i in range(n)])
    out = dict(knots=knots, knot_off=knot_off, path_id=np.arange(n, dtype=np.int32), start=g["start"].copy(),
               ego_xy=g["ego_xy"].copy(), obs=np.nan_to_num(g["obs"]), n_obs=g["n_obs"].astype(np.int32))
    if spline_build is not None:
        out["coef"] = np.concatenate([spline_build(knots[knot_off[i]:knot_off[i + 1]]) for i in range(n)])
    return out


def synth_routes(rng, n_paths):
    """Smooth random lanes, 2 m knot spacing like dense_polyline2d(., 2)."""
    knots, off = [], [0]
    for _ in range(n_paths):
        L = rng.uniform(60, 220)
        s = np.linspace(0, L, int(round(L / 2)))
        a, w, ph = rng.uniform(0, 8), rng.uniform(15, 60), rng.uniform(0, 6.28)
        th = rng.uniform(-3.14, 3.14)
        x, y = s, a * np.sin(s / w + ph)
        xy = np.stack([x * np.cos(th) - y * np.sin(th), x * np.sin(th) + y * np.cos(th)], axis=1) + rng.uniform(-100, 100, 2)
        knots.append(xy)
        off.append(off[-1] + len(xy))
    return np.concatenate(knots), np.array(off, np.int64)


def synth_scenarios(rng, n, knots, knot_off, max_obs=6):
    """Random planner inputs: Frenet start state along a random route, ego speed 0..16 m/s (some
    beyond MAX_SPEED), up to max_obs slow obstacles on the candidate corridor ahead."""
    n_paths = len(knot_off) - 1
    path_id = rng.integers(0, n_paths, n).astype(np.int32)
    seg = np.diff(knots, axis=0)
    cum = np.r_[0, np.cumsum(np.hypot(seg[:, 0], seg[:, 1]))]
    start = np.zeros((n, 5))
    ego = np.zeros((n, 2))
    obs = np.zeros((n, max_obs, 5))
    n_obs = rng.integers(0, max_obs + 1, n).astype(np.int32)
    for i in range(n):
        k0, k1 = knot_off[path_id[i]], knot_off[path_id[i] + 1]
        length = cum[k1 - 1] - cum[k0]
        s0 = rng.uniform(0.5, 0.9 * length)
        c_d = rng.uniform(-3, 3)
        start[i] = [s0, c_d, rng.normal(0, 0.4), 0.0, rng.uniform(0.2, 16.0) if rng.random() < 0.15 else rng.uniform(0.2, 11.0)]
        j = k0 + np.searchsorted(cum[k0:k1] - cum[k0], s0, side="right") - 1
        j = min(j, k1 - 2)
        u = knots[j + 1] - knots[j]
        u /= np.hypot(*u)
        nrm = np.array([-u[1], u[0]])
        p = knots[j] + u * (s0 - (cum[j] - cum[k0]))
        ego[i] = p + nrm * c_d
        for q in range(n_obs[i]):
            pos = p + u * rng.uniform(2, 30) + nrm * rng.uniform(-5, 5)
            v, yaw = rng.uniform(0, 2.0), rng.uniform(-3.14, 3.14)
            obs[i, q] = [pos[0], pos[1], v * np.cos(yaw), v * np.sin(yaw), yaw]
    return path_id, start, ego, obs, n_obs
