"""The decision logic of k_stream_rows, restated batch by batch in numpy and checked against the
sequential oracle.  CPU only.

The CUDA kernel (dcarl_b200/csrc/stream.cu) does not replay the reference's loop row by row: it
resolves the 32 rows of a batch in parallel under three claims that are properties of the algorithm,
not of the hardware —
  (1) a row of the argmax cell keeps the argmax while its value stays above the bound in force;
      a row of another cell keeps it while its value stays below the state's max; if that value
      rises above the bound it only lifts the bound for the later rows of its state ("lifter");
  (2) the state's max just before a row is the value of the latest earlier argmax-cell row of its
      state in the batch (else the max at the start of the batch);
  (3) the argmax only changes at rows that fail (1), so activation steps can be recorded there.
This file models exactly that (one "lane" per row, masks instead of ballots) so the claims are
tested for themselves, on the same inputs as the GPU tests, without a device.  It is test
infrastructure: the product never imports it.
"""
import math

import numpy as np
import pytest

from oracle import cport
from helpers import synth_streams


def _value(n, s1, s2, is_rule, c_h, cap):
    mean = s1 / n
    hn = c_h / math.sqrt(n)
    if is_rule:
        return min(cap, mean + hn)                                              # ST1:12
    var = max(s2 / n - mean * mean, 0.0)
    ci = (mean - 4.0 * math.sqrt(var) + s1) / (n + 1) - c_h / math.sqrt(n + 1)  # ST1:24 (algebraically)
    return min(mean - hn, ci)                                                   # ST1:16, :89


def _bound(second, mx):
    """runner_up_bound of the kernel: the midpoint of [second, max)."""
    if second == -math.inf:
        return second
    if not second < mx:
        return mx
    b = second + 0.5 * (mx - second)
    return b if b < mx else second


def _top2(vals):
    g = int(np.argmax(vals))                       # first maximum, like np.argmax / the kernel
    rest = np.delete(vals, g)
    return float(vals[g]), g, float(rest.max()) if rest.size else -math.inf


def emulate(idx, act, R, S, A, alpha=0.05, scale=150.0, cap=100.0, prior_rule=100.0, prior_other=-50.0,
            bonus=0.9, n_thres=10, rule_act=0, batch=32):
    """One instance.  Returns value, act, activation_step, overall like the kernel, plus counters."""
    N = len(idx)
    c_h = scale * math.sqrt(math.log(1.0 / alpha) / 2.0)
    cnt = np.zeros((S, A), np.int64)
    s1 = np.zeros((S, A))
    s2 = np.zeros((S, A))
    val = np.full((S, A), prior_other)
    val[:, rule_act] = prior_rule
    m0, g0, sec0 = _top2(val[0])
    smax, sarg, sub2 = np.full(S, m0), np.full(S, g0), np.full(S, _bound(sec0, m0))
    actstep = np.full(S, -1, np.int64)
    force_first = g0 != rule_act
    out_v, out_a, out_o = np.empty(N), np.empty(N, np.int64), np.empty(N)
    total = 0.0
    stats = dict(replays=0, lifters=0, rounds=0)
    for b0 in range(0, N, batch):
        rows = list(range(b0, min(b0 + batch, N)))
        L = len(rows)
        s_, a_ = idx[rows], act[rows]
        # ---- moment stage: running moments in time order, value per row
        v = np.empty(L)
        upd = np.zeros(L, bool)
        for i, k in enumerate(rows):
            s, a = s_[i], a_[i]
            cnt[s, a] += 1
            s1[s, a] += R[k]
            s2[s, a] += R[k] * R[k]
            upd[i] = cnt[s, a] > n_thres
            v[i] = _value(cnt[s, a], s1[s, a], s2[s, a], a == rule_act, c_h, cap)
        state_rows_after = {int(s): cnt[s].sum() for s in set(s_.tolist())}   # rows of the state incl. this batch
        # ---- selector stage
        m_out, mprev, g_out = np.zeros(L), np.zeros(L), np.full(L, -1)
        act_prev, jump = np.zeros(L, bool), np.zeros(L, bool)
        pend = np.ones(L, bool)
        while pend.any():
            stats["rounds"] += 1
            mine = pend.copy()
            c_ms0, c_ub, c_g, c_act = smax[s_], sub2[s_], sarg[s_], actstep[s_]
            T = mine & upd & (a_ == c_g)
            m_prev = c_ms0.copy()
            for i in range(L):                                   # claim (2)
                for j in range(i - 1, -1, -1):
                    if T[j] and s_[j] == s_[i]:
                        m_prev[i] = v[j]
                        break
            m_k = np.where(T, v, m_prev)
            forced = mine & (c_act == -1) if force_first else np.zeros(L, bool)
            lifter = mine & upd & ~T & (v > c_ub) & (v < m_k) & ~forced
            B = c_ub.copy()
            for l in np.nonzero(lifter)[0]:                      # claim (1): lifters raise the bound for later rows
                later = (np.arange(L) > l) & (s_ == s_[l])
                B[later] = np.maximum(B[later], v[l])
            doubt = (mine & upd & np.where(T, ~(v > B), ~(v < m_k))) | forced
            first = int(np.nonzero(doubt)[0][0]) if doubt.any() else L
            done = mine & (np.arange(L) < first)
            g_out[done], m_out[done], mprev[done] = c_g[done], m_k[done], m_prev[done]
            act_prev[done] = c_act[done] != -1
            for i in np.nonzero(done)[0]:
                if upd[i]:
                    val[s_[i], a_[i]] = v[i]                     # the last row of a cell wins (time order)
                smax[s_[i]] = m_k[i]                             # the last row of a state wins
            for s in set(s_[done].tolist()):
                rs = np.nonzero(done & (s_ == s))[0]
                lf = rs[lifter[rs]]
                if lf.size:                                      # last finalized row of a lifted state re-centres
                    last = rs[-1]
                    bself = max(B[last], v[last]) if lifter[last] else B[last]
                    sub2[s] = _bound(bself, m_k[last])
                    stats["lifters"] += lf.size
            pend &= ~done
            if first == L:
                break
            # ---- exact replay of row `first`                     claim (3)
            j = first
            stats["replays"] += 1
            sj = s_[j]
            if upd[j]:
                val[sj, a_[j]] = v[j]
            old_g, old_act = sarg[sj], actstep[sj]
            mx, arg, second = _top2(val[sj])
            g_out[j], m_out[j], mprev[j] = arg, mx, m_prev[j]
            act_prev[j] = old_act != -1
            if old_act == -1 and arg != rule_act:
                later_same = int(((np.arange(L) > j) & (s_ == sj)).sum())
                actstep[sj] = state_rows_after[int(sj)] - later_same      # 1-based row count of the state at row j
                jump[j] = True
            smax[sj], sarg[sj], sub2[sj] = mx, arg, _bound(second, mx)
            pend[j] = False
        # ---- outputs: overall_value as a running sum of per-row differences (ST2:99-105)
        d = np.where(act_prev, m_out - mprev, 0.0) + np.where(jump, m_out + bonus, 0.0)
        pre = np.cumsum(d)
        out_o[rows] = total + pre
        total += pre[-1]
        out_v[rows], out_a[rows] = m_out, g_out
    return out_v, out_a, actstep, out_o, stats


CASES = [
    (20, 11, 4000, {}),
    (1, 30, 3000, {}),
    (4, 6, 1500, {"n_thres": 0}),
    (4, 6, 1500, {"prior_other": 100.0}),
    (4, 6, 1500, {"rule_act": 3, "prior_other": 100.0}),     # priors' argmax is not the rule action: forced first rows
    (4, 6, 1500, {"prior_rule": -80.0}),
    (3, 5, 1500, {"alpha": 0.2, "scale": 40.0, "cap": 60.0, "activation_bonus": 2.5}),
]


@pytest.mark.parametrize("case", CASES)
def test_batched_selector_equals_the_sequential_loop(case):
    S, A, N, kw = case
    idx, act, R = synth_streams(3, N, S, A, seed=17 + N + S, r_dtype=np.float64)
    ref = cport.stream(idx, act, R, S, A, cport.default_params(**kw))
    ekw = dict(kw)
    if "activation_bonus" in ekw:
        ekw["bonus"] = ekw.pop("activation_bonus")
    for i in range(3):
        v, a, ast, ov, stats = emulate(idx[i], act[i], R[i], S, A, **ekw)
        assert np.array_equal(a, ref["act"][i])
        assert np.array_equal(ast, ref["activation_step"][i])
        assert np.abs(v - ref["value"][i]).max() < 1e-9
        assert np.abs(ov - ref["overall"][i]).max() < 1e-7
        assert stats["replays"] < 0.05 * N          # the exact path stays rare


def test_integer_returns_with_exact_ties():
    rng = np.random.default_rng(3)
    S, A, N = 3, 5, 1500
    idx = rng.integers(0, S, (2, N)).astype(np.int32)
    act = rng.integers(0, A, (2, N)).astype(np.int32)
    R = np.round(rng.normal(40, 3, (2, N)) / 20) * 20
    ref = cport.stream(idx, act, R, S, A, cport.default_params(n_thres=2))
    for i in range(2):
        v, a, ast, ov, _ = emulate(idx[i], act[i], R[i], S, A, n_thres=2)
        assert np.array_equal(a, ref["act"][i]) and np.array_equal(ast, ref["activation_step"][i])
        assert np.abs(v - ref["value"][i]).max() < 1e-9
