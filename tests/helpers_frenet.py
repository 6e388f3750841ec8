"""Shared helpers of the Frenet-planner tests: golden fixture -> batched planner inputs, and a
synthetic scenario generator for sizes the golden file does not cover."""
import math

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


def _point_to_segment(x0, y0, x1, y1, x2, y2):
    """tools.py:122-135 dist_from_point_to_line2d."""
    l = math.sqrt((y2 - y1) * (y2 - y1) + (x2 - x1) * (x2 - x1))
    if l == 0:
        return math.sqrt((y0 - y1) * (y0 - y1) + (x0 - x1) * (x0 - x1)), 0, 0
    dl = ((y2 - y1) * x0 - (x2 - x1) * y0 + x2 * y1 - x1 * y2) / l
    d1 = (x1 * x1 + x0 * (x2 - x1) - x1 * x2 + y1 * y1 + y0 * (y2 - y1) - y1 * y2) / l
    d2 = (x2 * x2 - x0 * (x2 - x1) - x1 * x2 + y2 * y2 - y0 * (y2 - y1) - y1 * y2) / l
    return dl, d1, d2


def frenet_state(x0, y0, vx, vy, line):
    """numpy restatement (test infrastructure) of calculate_start_state (JTP:248-276): tools.get_frenet_state
    (tools.py:235-270) over dist_from_point_to_polyline2d (tools.py:138-232), with the all-zero tangent table
    the planner passes (JTP:84).  Returns (s, d, vd).  Pinned against the reference's run in test_frenet_cpu.py;
    the product computes this on the device (dcarl_frenet_start)."""
    line = np.asarray(line, np.float64)
    dist_line = np.linalg.norm(line - [x0, y0], axis=1)
    ci = int(np.argmin(dist_line))
    n = len(line)
    ctype = 0
    seg = lambda i, j: _point_to_segment(x0, y0, line[i, 0], line[i, 1], line[j, 0], line[j, 1])
    nh = nt = ph = pt = 0.0
    if ci == 0:
        dn, nh, nt = seg(0, 1)
        if nh < 0:
            dc = dist_line[ci] * (-1 if dn < 0 else 1)
        else:
            dc, ctype = dn, 1
    elif ci == n - 1:
        dp, ph, pt = seg(n - 2, n - 1)
        if pt < 0:
            dc = dist_line[ci] * (-1 if dp < 0 else 1)
        else:
            dc, ctype = dp, -1
    else:
        dp, ph, pt = seg(ci - 1, ci)
        dn, nh, nt = seg(ci, ci + 1)
        if pt < 0 and nh < 0:
            dc = dist_line[ci]
            if _point_to_segment(line[ci + 1, 0], line[ci + 1, 1], line[ci - 1, 0], line[ci - 1, 1],
                                 line[ci, 0], line[ci, 1])[0] > 0:
                dc *= -1
        elif pt < 0:
            dc, ctype = dn, 1
        elif nh < 0:
            dc, ctype = dp, -1
        elif abs(dn) > abs(dp):
            dc, ctype = dp, -1
        else:
            dc, ctype = dn, 1
    seglen = lambda pts: np.sum(np.linalg.norm(np.diff(pts, axis=0), axis=1))
    if ctype == 1:
        s = nh + seglen(line[:ci + 1])
        psi = math.atan2(line[ci + 1, 1] - line[ci, 1], line[ci + 1, 0] - line[ci, 0])
    elif ctype == -1:
        s = ph + seglen(line[:ci])
        psi = math.atan2(line[ci, 1] - line[ci - 1, 1], line[ci, 0] - line[ci - 1, 0])
    else:
        s = seglen(line[:ci + 1])
        psi = 0.0
    vd = -math.sin(psi) * vx + math.cos(psi) * vy
    return float(s), float(dc), float(vd)


