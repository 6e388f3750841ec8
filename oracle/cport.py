"""ctypes bindings to oracle/dcarl_oracle.c (see that file's header for provenance)."""
import ctypes as C
import os
import subprocess

import numpy as np

_DIR = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_DIR, "libdcarl_oracle.so")
_lib = None


class OrcParams(C.Structure):
    _fields_ = [("alpha", C.c_double), ("scale", C.c_double), ("cap", C.c_double),
                ("prior_rule", C.c_double), ("prior_other", C.c_double),
                ("activation_bonus", C.c_double), ("n_thres", C.c_int32),
                ("rule_act", C.c_int32)]


def default_params(**kw):
    p = dict(alpha=0.05, scale=150.0, cap=100.0, prior_rule=100.0, prior_other=-50.0,
             activation_bonus=0.9, n_thres=10, rule_act=0)
    p.update(kw)
    return OrcParams(**p)


def build(force=False):
    srcs = [os.path.join(_DIR, f) for f in ("dcarl_oracle.c", "frenet_oracle.c")]
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < max(os.path.getmtime(f) for f in srcs):
        subprocess.check_call(["make", "-C", _DIR, "-B", "libdcarl_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_SO)
        d = C.c_double
        _lib.orc_np_sum.restype = d
        _lib.orc_upper_bound.restype = d
        _lib.orc_lower_bound.restype = d
        _lib.orc_ci_lower_bound.restype = d
        _lib.orc_mean_value.restype = d
        _lib.orc_sample_returns.restype = C.c_long
    return _lib


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


def _f64(x):
    return np.ascontiguousarray(x, np.float64)


def _i32(x):
    return np.ascontiguousarray(x, np.int32)


def np_sum(x):
    x = _f64(x)
    return lib().orc_np_sum(_p(x, C.c_double), C.c_long(x.size))


def bounds(x, alpha=0.05, scale=150.0, cap=100.0):
    """(upper_bound, lower_bound, CI_lower_bound, mean_value) of one sample array."""
    x = _f64(x)
    n = C.c_long(x.size)
    tmp = np.empty(max(x.size, 1))
    px, a, s, c = _p(x, C.c_double), C.c_double(alpha), C.c_double(scale), C.c_double(cap)
    L = lib()
    return (L.orc_upper_bound(px, n, a, s, c), L.orc_lower_bound(px, n, a, s),
            L.orc_ci_lower_bound(px, n, a, s, _p(tmp, C.c_double)), L.orc_mean_value(px, n, c))


def stream(idx, act, R, S, A, params=None, incremental=False, want_overall=True):
    """idx/act/R: [n_steps] or [n_inst, n_steps].  Returns dict of row-ordered outputs."""
    idx, act, R = _i32(idx), _i32(act), _f64(R)
    single = idx.ndim == 1
    if single:
        idx, act, R = idx[None], act[None], R[None]
    n_inst, n_steps = idx.shape
    p = params or default_params()
    val = np.empty((n_inst, n_steps))
    oact = np.empty((n_inst, n_steps), np.int32)
    astep = np.empty((n_inst, S), np.int64)
    ov = np.empty((n_inst, n_steps)) if want_overall else None
    rc = lib().orc_stream_batch(_p(idx, C.c_int32), _p(act, C.c_int32), _p(R, C.c_double),
                                C.c_long(n_inst), C.c_long(n_steps), C.c_int(S), C.c_int(A),
                                C.byref(p), C.c_int(int(incremental)), _p(val, C.c_double),
                                _p(oact, C.c_int32), _p(astep, C.c_int64),
                                _p(ov, C.c_double) if want_overall else None)
    if rc:
        raise ValueError("oracle: idx/act out of range")
    out = dict(value=val, act=oact, activation_step=astep, overall=ov)
    if single:
        out = {k: (v[0] if v is not None else None) for k, v in out.items()}
    return out


def box_moments(buf_state, buf_ret, dist, queries, n_actions):
    bs, br, dist, q = _f64(buf_state), _f64(buf_ret), _f64(dist).ravel(), _f64(queries)
    q = q.reshape(-1, bs.shape[1] - 1)
    out = np.empty((q.shape[0], n_actions, 3))
    lib().orc_box_moments(_p(bs, C.c_double), _p(br, C.c_double), C.c_long(bs.shape[0]),
                          C.c_int(bs.shape[1] - 1), _p(dist, C.c_double), _p(q, C.c_double),
                          C.c_int(q.shape[0]), C.c_int(n_actions), _p(out, C.c_double))
    return out


def box_raw(buf_state, buf_ret, dist, queries, n_actions, threads=1):
    """Raw moments [Q, n_actions, 3] = (n, sum R, sum R^2) in one pass per query (orc_box_raw).  The
    buffer may be float32 (the device copy, widened per element) or float64.  `threads` > 1 splits the
    rows into contiguous chunks (ctypes releases the GIL) and adds the partial moments in chunk order."""
    bs = np.ascontiguousarray(buf_state)
    f32 = bs.dtype == np.float32
    if not f32:
        bs = _f64(bs)
    br = np.ascontiguousarray(buf_ret, np.float32 if f32 else np.float64)
    dist, q = _f64(dist).ravel(), _f64(queries)
    D = bs.shape[1] - 1
    q = q.reshape(-1, D)
    Q, M = q.shape[0], bs.shape[0]
    t = C.c_float if f32 else C.c_double

    def part(lo, hi):
        out = np.empty((Q, n_actions, 3))
        lib().orc_box_raw(_p(bs[lo:hi], t), _p(br[lo:hi], t), C.c_int(int(f32)), C.c_long(hi - lo), C.c_int(D),
                          _p(dist, C.c_double), _p(q, C.c_double), C.c_int(Q), C.c_int(n_actions), _p(out, C.c_double))
        return out

    threads = max(1, min(int(threads), max(1, M // 4096)))
    if threads == 1:
        return part(0, M)
    from concurrent.futures import ThreadPoolExecutor
    edges = [M * i // threads for i in range(threads + 1)]
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda e: part(*e), zip(edges[:-1], edges[1:])))
    tot = outs[0]
    for o in outs[1:]:
        tot = tot + o
    return tot


def box_select(nmv, visited_times_thres=10, min_candidate_visits=5, rule_mean_gate=-0.1,
               confidence_thres=0.5):
    nmv = _f64(nmv)
    Q, na, _ = nmv.shape
    out = np.empty(Q, np.int32)
    lib().orc_box_select(_p(nmv, C.c_double), C.c_int(Q), C.c_int(na),
                         C.c_int(visited_times_thres), C.c_int(min_candidate_visits),
                         C.c_double(rule_mean_gate), C.c_double(confidence_thres),
                         _p(out, C.c_int32))
    return out


def sample_returns(q_star, z_state, act, z_ret, loc=3.0, sd=1.0, span=6.0, sigma=50.0):
    q_star, z_state, z_ret, act = _f64(q_star), _f64(z_state), _f64(z_ret), _i32(act)
    S, A = q_star.shape
    n = z_state.size
    idx = np.empty(n, np.int32)
    R = np.empty(n)
    nv = lib().orc_sample_returns(_p(q_star, C.c_double), _p(z_state, C.c_double),
                                  _p(act, C.c_int32), _p(z_ret, C.c_double), C.c_long(n),
                                  C.c_int(S), C.c_int(A), C.c_double(loc), C.c_double(sd),
                                  C.c_double(span), C.c_double(sigma), _p(idx, C.c_int32),
                                  _p(R, C.c_double))
    return idx, R, nv


def rollout_cv(state0, action, steps, dt, radius, pass_x, speed_scale):
    s0 = np.ascontiguousarray(state0, np.float32).reshape(-1, 20)
    action = _i32(action)
    ss = np.ascontiguousarray(speed_scale, np.float32)
    n = s0.shape[0]
    ret = np.empty(n, np.float32)
    st = np.empty(n, np.int32)
    tm = np.empty(n, np.int32)
    rc = lib().orc_rollout_cv(_p(s0, C.c_float), _p(action, C.c_int32), C.c_long(n),
                              C.c_int(steps), C.c_float(dt), C.c_float(radius),
                              C.c_float(pass_x), _p(ss, C.c_float), C.c_int(ss.size),
                              _p(ret, C.c_float), _p(st, C.c_int32), _p(tm, C.c_int32))
    if rc:
        raise ValueError("oracle: action out of range")
    return ret, st, tm


# ---- Frenet candidate planner (oracle/frenet_oracle.c) ---------------------------------------
FR_MAX_D, FR_MAX_T, FR_MAX_V = 16, 4, 8


class FrenetParams(C.Structure):
    """Field-for-field mirror of fr_params (frenet_oracle.c) / dcarl_frenet_params (the C ABI)."""
    _fields_ = [("d_samples", C.c_double * FR_MAX_D), ("T_samples", C.c_double * FR_MAX_T),
                ("tv_samples", C.c_double * FR_MAX_V), ("n_t", C.c_int32 * FR_MAX_T),
                ("n_d", C.c_int32), ("n_T", C.c_int32), ("n_v", C.c_int32), ("n_pred", C.c_int32),
                ("obstacles_considered", C.c_int32), ("pad_", C.c_int32),
                ("dt", C.c_double), ("target_speed", C.c_double),
                ("max_speed", C.c_double), ("max_accel", C.c_double), ("max_curvature", C.c_double),
                ("robot_radius", C.c_double), ("move_gap", C.c_double),
                ("kj", C.c_double), ("kt", C.c_double), ("kd", C.c_double), ("klat", C.c_double), ("klon", C.c_double)]


def frenet_params(MAX_SPEED=50.0 / 3.6, MAX_ACCEL=10.0, MAX_CURVATURE=500.0, MAX_LEFT_WIDTH=-4, MAX_RIGHT_WIDTH=4,
                  D_ROAD_W=2, DT=0.3, MAXT=4.2, MINT=4.0, TARGET_SPEED=30.0 / 3.6, D_T_S=15 / 3.6, N_S_SAMPLE=1,
                  OBSTACLES_CONSIDERED=4, ROBOT_RADIUS=1, MOVE_GAP=1, KJ=0.1, KT=0.1, KD=1.0, KLAT=1.0, KLON=1.0):
    """The module constants of JunctionTrajectoryPlanner.py:13-39 turned into the sample grids the
    planner loops over (np.arange, like the reference)."""
    p = FrenetParams()
    d = np.arange(MAX_LEFT_WIDTH, MAX_RIGHT_WIDTH + 1, D_ROAD_W).astype(np.float64)
    T = np.arange(MINT, MAXT, DT).astype(np.float64)
    tv = np.arange(TARGET_SPEED - D_T_S * N_S_SAMPLE, TARGET_SPEED + D_T_S * N_S_SAMPLE, D_T_S).astype(np.float64)
    assert len(d) <= FR_MAX_D and len(T) <= FR_MAX_T and len(tv) <= FR_MAX_V
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


def spline_build(xy):
    """xy [n, 2] knots -> coef [n, 9] rows {s, ax, bx, cx, dx, ay, by, cy, dy} (cubic_spline_planner.py:18-45)."""
    xy = _f64(xy)
    coef = np.empty((xy.shape[0], 9))
    rc = lib().orc_spline_build(_p(xy, C.c_double), C.c_int(xy.shape[0]), _p(coef, C.c_double))
    if rc:
        raise ValueError("oracle: knot count out of range")
    return coef


def frenet_plan(coef, knot_off, path_id, start, ego_xy, obs, n_obs, params, want_paths=True):
    """n scenarios: see orc_frenet_plan.  Returns a dict of numpy arrays."""
    coef, start, ego_xy = _f64(coef), _f64(start).reshape(-1, 5), _f64(ego_xy).reshape(-1, 2)
    knot_off = np.ascontiguousarray(knot_off, np.int64)
    path_id, n_obs = _i32(path_id), _i32(n_obs)
    n = start.shape[0]
    obs = _f64(obs).reshape(n, -1, 5) if np.size(obs) else np.zeros((n, 1, 5))
    max_obs = obs.shape[1]
    Cn = params.n_d * params.n_T * params.n_v
    Tm = max(params.n_t[:params.n_T])
    if Cn > 32 or Tm > 32 or max_obs > 16:
        raise ValueError("oracle: at most 32 candidates, 32 points per path, 16 obstacles per scenario")
    out = dict(index=np.empty(n, np.int32), cost=np.empty((n, Cn)), flags=np.empty((n, Cn), np.uint8),
               n_valid=np.empty((n, Cn), np.int32), paths=np.empty((n, Cn, Tm, 2)) if want_paths else None,
               traj=np.empty((n, Tm, 3)), traj_len=np.empty(n, np.int32))
    lib().orc_frenet_plan(_p(coef, C.c_double), _p(knot_off, C.c_int64), _p(path_id, C.c_int32),
                          _p(start, C.c_double), _p(ego_xy, C.c_double), _p(obs, C.c_double),
                          _p(n_obs, C.c_int32), C.c_int(max_obs), C.c_int64(n), C.byref(params),
                          _p(out["index"], C.c_int32), _p(out["cost"], C.c_double), _p(out["flags"], C.c_uint8),
                          _p(out["n_valid"], C.c_int32),
                          _p(out["paths"], C.c_double) if want_paths else None,
                          _p(out["traj"], C.c_double), _p(out["traj_len"], C.c_int32))
    return out
