"""Shared test utilities: synthetic workloads and comparison helpers."""
import numpy as np


def synth_streams(n_inst, n_steps, S=20, A=11, seed=1000, r_dtype=np.float32, live_actions=None):
    """Sim-2-shaped instances: idx ~ floor(N(3,1)/6*S) (redrawn until valid), uniform action,
    R = Q*[idx, act] + 50 z with a per-instance Q* ~ U(-50, 100)."""
    rng = np.random.default_rng(seed)
    q = rng.uniform(-50, 100, (n_inst, S, A))
    idx = np.floor(rng.normal(3, 1, (n_inst, n_steps)) / 6 * S).astype(np.int64)
    bad = (idx < 0) | (idx >= S)
    while bad.any():
        idx[bad] = np.floor(rng.normal(3, 1, int(bad.sum())) / 6 * S).astype(np.int64)
        bad = (idx < 0) | (idx >= S)
    act = rng.integers(0, live_actions or A, (n_inst, n_steps))
    R = q[np.arange(n_inst)[:, None], idx, act] + 50 * rng.standard_normal((n_inst, n_steps))
    return idx.astype(np.int32), act.astype(np.int32), R.astype(r_dtype)


def compare_stream(res, ref, value_tol=1e-5, max_near_ties=0):
    """res: GPU StreamResult-like (numpy), ref: oracle dict.  Argmax must be bit-exact except at
    near-ties (both picks within 1e-9 of each other), which are counted and bounded."""
    v, rv = np.asarray(res.value, np.float64), np.asarray(ref["value"], np.float64)
    scale = np.maximum(1.0, np.abs(rv))
    err = np.abs(v - rv) / scale
    assert np.nanmax(err) <= value_tol if err.size else True, f"value error {np.nanmax(err)}"
    mism = np.asarray(res.act) != np.asarray(ref["act"])
    n_mis = int(mism.sum())
    if n_mis:
        assert n_mis <= max_near_ties, f"{n_mis} argmax mismatches"
        assert np.abs(v[mism] - rv[mism]).max() < 1e-9, "argmax mismatch away from a tie"
    else:
        assert np.array_equal(res.activation_step, ref["activation_step"])
    if res.overall is not None and ref.get("overall") is not None and n_mis == 0:
        ro = np.asarray(ref["overall"], np.float64)
        eo = np.abs(np.asarray(res.overall, np.float64) - ro) / np.maximum(1.0, np.abs(ro))
        assert (eo.max() if eo.size else 0.0) <= value_tol, f"overall error {eo.max()}"
    return float(np.nanmax(err)) if err.size else 0.0, n_mis


def stats_reference(value, overall, activation_step, act, A):
    """numpy restatement of dcarl_stream_stats for outputs [n_inst, n_steps]."""
    n_inst, n_steps = value.shape
    st = np.zeros(5 + A)
    if n_steps:
        st[0] = value[:, -1].astype(np.float64).sum()
        st[1] = overall[:, -1].astype(np.float64).sum() if overall is not None else 0.0
    on = activation_step != -1
    st[2] = on.sum()
    st[3] = activation_step[on].sum()
    st[4] = n_inst * n_steps
    st[5:] = np.bincount(act.ravel()[act.ravel() >= 0], minlength=A)[:A]
    return st


def recompute_from_traces(idx, value, act, S, rule_act=0, bonus=0.9):
    """Size-independent consistency: activation_step and overall_value re-derived from the
    per-row (value, act) traces of ONE instance (test_DCARL.py:96-105)."""
    n = len(idx)
    activation = np.full(S, -1, np.int64)
    last = np.zeros(S)
    activated = np.zeros(S, bool)
    count = np.zeros(S, np.int64)
    overall = np.empty(n)
    for k in range(n):
        s = idx[k]
        count[s] += 1
        last[s] = value[k]
        if not activated[s] and act[k] != rule_act:
            activated[s] = True
            activation[s] = count[s]
        overall[k] = (last[activated] + bonus).sum()
    return activation, overall


def face_queries(states, dist, n, seed, eps=1e-5):
    """n query observations that sit within +-eps of a face of some stored row's box: the row's own
    observation, with one coordinate moved to (stored value +- half-width) + U(-eps, eps).  Returns
    (queries [n, D], action of the row [n])."""
    rng = np.random.default_rng(seed)
    D = states.shape[1] - 1
    rows = rng.integers(0, len(states), n)
    dim = rng.integers(0, D, n)
    sign = rng.choice([-1.0, 1.0], n)
    q = states[rows, :D].astype(np.float64).copy()
    q[np.arange(n), dim] += sign * dist[dim] + rng.uniform(-eps, eps, n)
    return q, states[rows, D].astype(np.int64)
