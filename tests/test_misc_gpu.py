"""Bound functions, sampler and rollout surrogate against the reference's golden outputs / the oracle."""
import random

import numpy as np
import pytest
import torch

import dcarl_b200 as dc
from oracle import cport

pytestmark = pytest.mark.gpu


def test_bound_functions_golden(golden):
    g = golden("bounds")
    off = np.concatenate([[0], np.cumsum(g["lens"])])
    arrays = [g["flat"][off[i]:off[i + 1]] for i in range(len(g["lens"]))]
    got = dc.bounds_batch(arrays[:-1])
    want = g["out"][:-1]
    assert np.abs(got - want).max() / 1.0 < 1e-9
    x = arrays[7]
    assert abs(dc.upper_bound(x) - g["out"][7, 0]) < 1e-10 and abs(dc.lower_bound(x) - g["out"][7, 1]) < 1e-10
    assert abs(dc.CI_lower_bound(x) - g["out"][7, 2]) < 1e-10 and abs(dc.mean_value(x) - g["out"][7, 3]) < 1e-10
    alpha, loc, scale = g["last_kwargs"]
    x = arrays[-1]
    got = [dc.upper_bound(x, alpha, loc, scale), dc.lower_bound(x, alpha=alpha, scale=scale),
           dc.CI_lower_bound(x, alpha, loc, scale), dc.mean_value(list(x))]
    assert np.abs(np.array(got) - g["out"][-1]).max() < 1e-10
    # constant samples: population std is exactly zero
    c = dc.bounds_batch([np.full(11, -50.0)])[0]
    assert c[2] == cport.bounds(np.full(11, -50.0))[2]
    assert np.isnan(dc.bounds_batch([np.zeros(0)])).all()
    # float32 device tensor in, device tensor out
    t = torch.randn(1000, device="cuda")
    o = dc.bounds_batch(flat=t, offsets=torch.tensor([0, 400, 1000], device="cuda"))
    assert o.is_cuda and o.shape == (2, 4)
    w = cport.bounds(t[:400].cpu().numpy().astype(np.float64))
    assert np.abs(o[0].cpu().numpy() - np.array(w)).max() < 1e-9


def test_cell_value_closed_form_is_accurate_to_a_few_ulp():
    """The rsqrt-based closed form of the stream kernel against the reference's divide/sqrt chain
    (test_DCARL.py:12,16,24) in float64: this bounds how close a tie must be to flip an argmax."""
    import math
    rng = np.random.default_rng(1)
    m = 200000
    n = np.concatenate([rng.integers(1, 300, m // 2), rng.integers(300, 2_000_000, m // 2)]).astype(np.int64)
    mean = rng.uniform(-200, 200, m)
    sd = rng.uniform(0, 80, m) * (rng.random(m) > 0.05)
    s1 = mean * n
    s2 = (sd ** 2 + (s1 / n) ** 2) * n
    act = rng.integers(0, 3, m).astype(np.int32)
    got = dc.cell_values(n, s1, s2, act)
    L = math.log(1 / 0.05)
    mu = s1 / n
    var = np.maximum(s2 / n - mu * mu, 0.0)
    ub = np.minimum(100.0, mu + 150 * np.sqrt(L / 2 / n))
    lb = mu - 150 * np.sqrt(L / 2 / n)
    ci = s1 / n / (n + 1) - 4 * np.sqrt(var) / (n + 1) + s1 / (n + 1) - 150 * np.sqrt(L / 2 / (n + 1))
    want = np.where(act == 0, ub, np.minimum(lb, ci))
    # sigma = sqrt(s2/n - mean^2) carries that subtraction's cancellation error on BOTH sides
    # (eps*(mean^2+var), divided by 2*sigma, times 4/(n+1)); everything else is a few ulp
    eps = 2.2e-16
    sig = np.sqrt(var)
    tol = 2e-13 * np.maximum(1.0, np.abs(want))
    nz = (act != 0) & (sig > 0)
    tol[nz] += 32 * eps * (mu[nz] ** 2 + var[nz]) / sig[nz] / (n[nz] + 1)
    tol[(act != 0) & (sig == 0)] += 1e-5
    err = np.abs(got - want)
    assert (err <= tol).all(), (err / tol).max()
    assert (err[act == 0] <= 2e-13 * np.maximum(1.0, np.abs(want[act == 0]))).all()
    assert np.isnan(dc.cell_values([0], [1.0], [1.0], [0]))[0]


def test_data_generation_reproduces_the_reference_files(golden, tmp_path):
    """Seed numpy / random like tools/make_golden.py did before calling the reference's
    Data_Generation(): the three .npy files must come out bit-identical."""
    g = golden("sampling")
    np.random.seed(int(g["seeds"][0]))
    random.seed(int(g["seeds"][1]))
    out = tmp_path / "Simulation_testing" / "Simulation_Data_Collection"
    dc.Data_Generation(str(out))
    assert np.array_equal(np.load(out / "data.npy"), g["data"])
    assert np.array_equal(np.load(out / "action_value.npy"), g["action_value"])
    assert np.array_equal(np.load(out / "states.npy"), g["states"])


def test_sampler_helpers_and_float32_mode(golden):
    g = golden("sampling")
    np.random.seed(7)
    idx = dc.random_state_norm(20, 5000)
    np.random.seed(7)
    assert np.array_equal(idx, np.floor((3 + np.random.standard_normal(5000)) / 6 * 20).astype(int))
    np.random.seed(9)
    r = dc.add_an_act_data(2, g["action_value"][4])
    np.random.seed(9)
    assert r == g["action_value"][4][2] + 50 * np.random.standard_normal(1)[0]
    assert len(dc.random_state_manual(20, 50)) == 50
    # float32 deviates (the 16 B/row bench layout) against the oracle on the same inputs
    rng = np.random.default_rng(0)
    n = 100000
    q = rng.uniform(-50, 100, (20, 11)).astype(np.float32)
    zs, zr = rng.standard_normal(n).astype(np.float32), rng.standard_normal(n).astype(np.float32)
    act = rng.integers(0, 11, n).astype(np.int32)
    idx, R = dc.sample_returns(q, zs, act, zr)
    oi, oR, nv = cport.sample_returns(q.astype(np.float64), zs.astype(np.float64), act, zr.astype(np.float64))
    assert np.array_equal(idx, oi) and R.dtype == np.float32
    assert np.array_equal(R, oR.astype(np.float32))                    # float64 arithmetic, one final rounding


def test_philox_sampler_is_reproducible_and_distributed_right():
    q = torch.rand((20, 11), device="cuda") * 150 - 50
    n = 1 << 20
    i1, a1, r1 = dc.sample_returns_philox(q, n, seed=1234)
    i2, a2, r2 = dc.sample_returns_philox(q, n, seed=1234)
    i3, _, _ = dc.sample_returns_philox(q, n, seed=1235)
    assert torch.equal(i1, i2) and torch.equal(a1, a2) and torch.equal(r1, r2) and not torch.equal(i1, i3)
    idx, act, R = i1.cpu().numpy(), a1.cpu().numpy(), r1.cpu().numpy().astype(np.float64)
    valid = idx >= 0
    # acceptance = P(|z| < 3) ; bins follow the normal cdf ; actions uniform ; residual ~ N(0, 50)
    from scipy.stats import norm
    assert abs(valid.mean() - (norm.cdf(3) - norm.cdf(-3))) < 1e-3
    edges = np.arange(21) * 0.3 - 3.0
    p = np.diff(norm.cdf(edges))
    assert np.abs(np.bincount(idx[valid], minlength=20) / n - p).max() < 2e-3
    assert np.abs(np.bincount(act, minlength=11) / n - 1 / 11).max() < 2e-3
    resid = R[valid] - q.cpu().numpy().astype(np.float64)[idx[valid], act[valid]]
    assert abs(resid.mean()) < 0.3 and abs(resid.std() - 50) < 0.3


def test_rollout_surrogate_matches_the_oracle_bitwise():
    rng = np.random.default_rng(4)
    n = 20000
    s0 = np.zeros((n, 20), np.float32)
    s0[:, 2] = rng.uniform(2, 8, n)
    for j in range(3):
        b = 5 + 5 * j
        s0[:, b] = rng.uniform(-30, 30, n); s0[:, b + 1] = rng.uniform(-6, 6, n)
        s0[:, b + 2] = rng.uniform(-5, 5, n); s0[:, b + 3] = rng.uniform(-1, 1, n)
        s0[:, b + 4] = np.arctan2(s0[:, b + 3], s0[:, b + 2])
    action = rng.integers(0, 11, n).astype(np.int32)
    ret, st, tm = dc.rollout_cv(s0, action)
    oret, ost, otm = cport.rollout_cv(s0, action, 200, 0.05, 1.0, 30.0, dc.rollout.DEFAULT_SPEED_SCALE)
    assert np.array_equal(tm, otm) and np.array_equal(st, ost) and np.array_equal(ret, oret)
    assert set(np.unique(tm)) == {0, 1, 2}
    with pytest.raises(IndexError):
        dc.rollout_cv(s0[:4], np.array([0, 1, 99, 2], np.int32))
