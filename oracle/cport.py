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
    src = os.path.join(_DIR, "dcarl_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _DIR, "-B", "libdcarl_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
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
