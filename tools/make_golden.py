#!/usr/bin/env python
"""Generate tests/golden/*.npz by RUNNING THE REFERENCE in this container.

The reference (/root/reference, read-only) does not travel to the GPU box, so
its inputs and the outputs it produces here are frozen as small fixtures.
Nothing under tests/, bench.py or the package reads /root/reference at run
time; this script is the only place that does and it is run by hand:

    python tools/make_golden.py            # needs /root/reference

Sources executed / parsed (paths relative to /root/reference):
  Simulation_testing/Simulation_1/test_DCARL.py      (runpy, __main__)
  Simulation_testing/Simulation_2/test_DCARL.py      (runpy, __main__)
  Simulation_testing/Simulation_Data_Collection/Data_Sampling/data_sampling.py
  Field_testing/.../tools/DCARL/{visited_state,visited_value,driving_record}.txt
"""
import contextlib
import hashlib
import io
import os
import random
import runpy
import sys
import tempfile
import types

import numpy as np

REF = "/root/reference"
RL = os.path.join(REF, "Field_testing/Software_and_Raw_Data_on_Self-Driving_Vehicle/"
                       "software/src/tools/DCARL")
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")


def _stub_matplotlib():
    """matplotlib is not installed; the scripts only plot with it."""
    class _Any(types.ModuleType):
        def __getattr__(self, name):
            return lambda *a, **k: _Any("x")

        def __call__(self, *a, **k):
            return _Any("x")

        def __iter__(self):
            return iter(())
    m, p = _Any("matplotlib"), _Any("matplotlib.pyplot")
    m.pyplot = p
    sys.modules["matplotlib"] = m
    sys.modules["matplotlib.pyplot"] = p


def _sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()[:16]


def _run_sim(script):
    cwd = os.getcwd()
    os.chdir(REF)
    try:
        with contextlib.redirect_stdout(io.StringIO()) as so:
            g = runpy.run_path(script, run_name="__main__")
    finally:
        os.chdir(cwd)
    return g, so.getvalue()


def _time_order(g, n_rows):
    """The scripts keep per-state lists; rebuild the per-row (time-ordered) trace."""
    data = g["data"][:n_rows]
    S = g["state_num"]
    pos = [0] * S
    val = np.empty(n_rows)
    act = np.empty(n_rows, np.int64)
    tru = np.empty(n_rows)
    for k in range(n_rows):
        s = int(data[k, 0])
        val[k] = g["step_TSRL_value"][s][pos[s]]
        act[k] = g["step_TSRL_act"][s][pos[s]]
        tru[k] = g["true_step_TSRL_value"][s][pos[s]]
        pos[s] += 1
    return val, act, tru


def golden_sim(which):
    script = f"Simulation_testing/Simulation_{which}/test_DCARL.py"
    g, printed = _run_sim(script)
    n_rows = 20000
    data = g["data"]
    val, act, tru = _time_order(g, n_rows)
    state_major_act = np.concatenate([np.asarray(x, np.int64) for x in g["step_TSRL_act"]])
    state_major_val = np.concatenate([np.asarray(x, np.float64) for x in g["step_TSRL_value"]])
    out = dict(
        idx=data[:, 0].astype(np.uint8), act=data[:, 2].astype(np.uint8),
        R=data[:, 3].astype(np.float64), state=data[:, 1].astype(np.float64),
        true_action_values=g["true_action_values"].astype(np.float64),
        S=np.int64(g["state_num"]), A=np.int64(g["action_num"]),
        n_thres=np.int64(g["n_thres"]), rule_act=np.int64(g["rule_act"]),
        n_rows=np.int64(n_rows),
        step_value=val, step_act=act, true_step_value=tru,
        activation_step=np.asarray(g["activation_step"], np.int64),
        final_value=np.asarray(g["TSRL_value"], np.float64),
        sha_act_state_major=np.array(_sha(state_major_act)),
        sha_val_state_major=np.array(_sha(state_major_val)),
        printed=np.array(printed),
    )
    assert np.array_equal(data[:, 0], out["idx"]) and np.array_equal(data[:, 2], out["act"])
    if "overall_value" in g and len(g["overall_value"]) == n_rows:
        out["overall_value"] = np.asarray(g["overall_value"], np.float64)
    np.savez_compressed(os.path.join(OUT, f"sim{which}.npz"), **out)
    print(f"sim{which}: activation_step={out['activation_step'].tolist()} "
          f"sha_act={out['sha_act_state_major']} sha_val={out['sha_val_state_major']}")


def golden_bounds():
    """Known-answer vectors for the four bound functions (test_DCARL.py:10-28)."""
    g, _ = _run_sim("Simulation_testing/Simulation_1/test_DCARL.py")
    rng = np.random.default_rng(20240917)
    cases, lens, outs = [], [], []
    for n in [1, 2, 3, 7, 8, 9, 11, 16, 127, 128, 129, 255, 1000, 1917, 4096]:
        for loc, sc in [(0.0, 50.0), (80.0, 1.0), (-30.0, 120.0)]:
            x = rng.normal(loc, sc, n)
            cases.append(x)
            lens.append(n)
            outs.append([g["upper_bound"](x), g["lower_bound"](x),
                         g["CI_lower_bound"](x), g["mean_value"](x)])
    # constant arrays: std must be exactly 0
    for n, c in [(11, -50.0), (40, 100.0), (5, 0.1)]:
        x = np.full(n, c)
        cases.append(x)
        lens.append(n)
        outs.append([g["upper_bound"](x), g["lower_bound"](x),
                     g["CI_lower_bound"](x), g["mean_value"](x)])
    # non-default keyword arguments
    x = rng.normal(10, 30, 57)
    kw = dict(alpha=0.1, loc=-50, scale=75)
    cases.append(x)
    lens.append(57)
    outs.append([g["upper_bound"](x, **kw), g["lower_bound"](x, **kw),
                 g["CI_lower_bound"](x, **kw), g["mean_value"](x, **kw)])
    np.savez_compressed(os.path.join(OUT, "bounds.npz"),
                        flat=np.concatenate(cases), lens=np.asarray(lens, np.int64),
                        out=np.asarray(outs, np.float64),
                        last_kwargs=np.array([0.1, -50.0, 75.0]))
    print(f"bounds: {len(lens)} cases")


def golden_sampling():
    """Run Data_Generation() with the global RNGs seeded, then replay the same
    draws to recover the underlying uniform/normal deviates (data_sampling.py:30-67).
    scipy's norm.rvs/uniform.rvs draw standard deviates from numpy's global
    RandomState and apply loc+scale*x, so a replay with the same seed in the same
    order yields the deviates the reference consumed."""
    path = os.path.join(REF, "Simulation_testing/Simulation_Data_Collection/"
                             "Data_Sampling/data_sampling.py")
    mod = runpy.run_path(path, run_name="not_main")
    seed_np, seed_py = 1234, 5678
    cwd = os.getcwd()
    with tempfile.TemporaryDirectory() as td:
        os.makedirs(os.path.join(td, "Simulation_testing/Simulation_Data_Collection"))
        os.chdir(td)
        try:
            np.random.seed(seed_np)
            random.seed(seed_py)
            import warnings
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                mod["Data_Generation"]()
            d = "Simulation_testing/Simulation_Data_Collection/"
            data = np.load(d + "data.npy")
            av = np.load(d + "action_value.npy")
            st = np.load(d + "states.npy")
        finally:
            os.chdir(cwd)
    S, A, N = 20, 11, 50000
    # replay
    np.random.seed(seed_np)
    random.seed(seed_py)
    rs = np.random.mtrand._rand
    u_states = rs.uniform(0.0, 1.0, S)
    u_av = np.stack([rs.uniform(0.0, 1.0, A) for _ in range(S)])
    z_state = rs.standard_normal(N)
    idx_all = np.floor((3 + 1 * z_state) / 6 * S).astype(int)
    valid = (idx_all >= 0) & (idx_all < S)
    acts = np.array([random.randint(0, A - 1) for _ in range(int(valid.sum()))], np.int64)
    z_ret = np.array([rs.standard_normal(1)[0] for _ in range(int(valid.sum()))])
    # the replay must reproduce what the reference wrote
    assert np.array_equal(st, 0 + 1 * u_states)
    assert np.array_equal(av, -50 + 150 * u_av)
    assert np.array_equal(data[:, 0], idx_all[valid])
    assert np.array_equal(data[:, 2], acts)
    assert np.array_equal(data[:, 3], av[idx_all[valid], acts] + 50 * z_ret)
    np.savez_compressed(os.path.join(OUT, "sampling.npz"),
                        u_states=u_states, u_action_values=u_av, z_state=z_state,
                        act=acts, z_ret=z_ret, data=data, action_value=av, states=st,
                        seeds=np.array([seed_np, seed_py]))
    print(f"sampling: {len(data)} rows of {N}; replay of the deviates verified")


def golden_rls():
    """Hard-negative subset of the shipped 209 600-row buffer + the 107 logged
    queries (RLS.py:161-181, record layout RLS.py:232-241).  The subset keeps every
    row that matches, or misses by exactly one dimension, any logged query, every
    non-rule-action row and a 2 % random sample of the rest, so the logged
    (n, mean, var) hold on the subset exactly as on the full buffer (asserted)."""
    vs = np.loadtxt(os.path.join(RL, "visited_state.txt"))
    vv = np.loadtxt(os.path.join(RL, "visited_value.txt"))
    dr = np.loadtxt(os.path.join(RL, "driving_record.txt"))
    dist = np.array([1, 0.3, 2, 50, 10, 0.3, 2, 50, 10, 0.3, 2, 50, 10, 0.3, 2, 50,
                     10, 0.3, 2, 50, 0.1])
    lo, hi = vs - dist, vs + dist
    keep = vs[:, 20] != 0
    full = []
    for r in dr:
        q = r[:21]
        nf = (~((lo <= q) & (q <= hi))).sum(1)
        keep |= nf <= 1
        m = nf == 0
        full.append([m.sum(), vv[m, 1].mean() if m.any() else -1,
                     vv[m, 1].var() if m.any() else -1])
    full = np.asarray(full)
    assert np.array_equal(full[:, 0], dr[:, 23])
    assert np.abs(full[:, 1:] - dr[:, 24:26]).max() < 1e-5
    keep |= np.random.default_rng(0).random(len(vs)) < 0.02
    svs, svv = vs[keep], vv[keep]
    slo, shi = svs - dist, svs + dist
    for r, f in zip(dr, full):
        m = np.all((slo <= r[:21]) & (r[:21] <= shi), axis=1)
        assert m.sum() == f[0]
    np.savez_compressed(os.path.join(OUT, "rls_box.npz"),
                        visited_state=svs, visited_value=svv, driving_record=dr,
                        dist=dist, full_rows=np.int64(len(vs)),
                        exact_nmv=full,
                        action_counts=np.bincount(vs[:, 20].astype(int), minlength=8))
    print(f"rls_box: kept {keep.sum()} of {len(vs)} rows; 107/107 logged queries reproduce")


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    _stub_matplotlib()
    golden_sim(1)
    golden_sim(2)
    golden_bounds()
    golden_sampling()
    golden_rls()
