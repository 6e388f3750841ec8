"""Parity of the CUDA stream kernels (through the C ABI) with the oracle and the reference's
golden outputs.  Integer outputs (argmax, activation step) bit-exact; values within 1e-5."""
import numpy as np
import pytest
import torch

import dcarl_b200 as dc
from dcarl_b200 import _abi
from oracle import cport
from helpers import compare_stream, recompute_from_traces, stats_reference, synth_streams

pytestmark = pytest.mark.gpu


def _oracle(idx, act, R, S, A, **kw):
    return cport.stream(idx, act, np.asarray(R, np.float64), S, A, cport.default_params(**kw))


@pytest.mark.parametrize("which", [1, 2])
@pytest.mark.parametrize("variant", [0, 1])
def test_golden_reference_traces(golden, which, variant):
    g = golden(f"sim{which}")
    S, A, N = int(g["S"]), int(g["A"]), int(g["n_rows"])
    res = dc.stream_bounds(g["idx"][:N].astype(np.int32), g["act"][:N].astype(np.int32), g["R"][:N], S, A,
                           variant=variant)
    assert res.value.dtype == np.float64
    assert np.array_equal(res.act, g["step_act"])                      # bit-exact argmax, 20 000 rows
    assert np.array_equal(res.activation_step, g["activation_step"])
    assert np.abs(res.value - g["step_value"]).max() < 1e-9
    if "overall_value" in g.files:
        assert np.abs(res.overall - g["overall_value"]).max() < 1e-8


def test_simulate_mirrors_the_script_globals(golden):
    g = golden("sim2")
    N = int(g["n_rows"])
    data = np.stack([g["idx"], g["state"], g["act"], g["R"]], axis=1).astype(np.float64)
    out = dc.simulate(data, g["true_action_values"], 20, 11)
    idx = g["idx"][:N]
    for s in range(20):
        rows = np.nonzero(idx == s)[0]
        assert out["step_TSRL_act"][s] == g["step_act"][rows].tolist()
        assert out["true_step_TSRL_value"][s] == g["true_step_value"][rows].tolist()
        assert np.allclose(out["step_TSRL_value"][s], g["step_value"][rows], rtol=0, atol=1e-9)
    assert out["activation_step"].tolist() == g["activation_step"].tolist()
    assert abs(out["overall_value"][-1] - 597.7193818873668) < 1e-8


SHAPES = [
    # n_inst, n_steps, S, A, params
    (48, 2500, 20, 11, {}),
    (33, 1999, 1, 30, {}),                       # Sim-1 shape, ragged tail
    (7, 31, 20, 11, {}),                         # shorter than one batch
    (5, 700, 3, 40, {}),                         # A > 32
    (5, 900, 40, 5, {"rule_act": 3}),            # S > 32, rule action moved
    (6, 800, 4, 6, {"n_thres": 0}),              # refresh from the first sample
    (6, 800, 4, 6, {"prior_other": 100.0}),      # priors tie with the rule action
    (4, 600, 2, 3, {"alpha": 0.2, "scale": 40.0, "cap": 60.0, "activation_bonus": 2.5}),
    (3, 4000, 30, 256, {}),                      # config-4-sized action set: 7680 cells, one instance per CTA
    (2, 300, 1, 1, {}),                          # a single cell
    # the priors' argmax is not the rule action: every state activates on its first row
    (6, 800, 4, 6, {"rule_act": 3, "prior_other": 100.0}),
    (6, 800, 4, 6, {"prior_rule": -80.0}),
    (6, 800, 4, 6, {"prior_rule": -80.0, "n_thres": 0}),
    (5, 700, 3, 40, {"rule_act": 2, "prior_other": 100.0}),
]


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("shape", SHAPES)
def test_random_instances_vs_oracle_f64(shape, variant):
    n_inst, n_steps, S, A, kw = shape
    idx, act, R = synth_streams(n_inst, n_steps, S, A, seed=7 + n_steps, r_dtype=np.float64)
    res = dc.stream_bounds(idx, act, R, S, A, _abi.Params(**kw), variant=variant)
    compare_stream(res, _oracle(idx, act, R, S, A, **kw), value_tol=1e-9)


def test_table_larger_than_shared_memory_takes_the_walker():
    """40 states x 256 actions = 10 240 cells do not fit one CTA's shared memory: the default variant
    must switch to the global-memory walker by itself (and variant 2 must refuse)."""
    idx, act, R = synth_streams(2, 1500, 40, 256, seed=77, r_dtype=np.float64)
    assert not _abi.lib().dcarl_stream_rows_fit(40, 256) and _abi.lib().dcarl_stream_rows_fit(30, 256)
    res = dc.stream_bounds(idx, act, R, 40, 256)
    compare_stream(res, _oracle(idx, act, R, 40, 256), value_tol=1e-9)
    with pytest.raises(dc.DcarlError):
        dc.stream_bounds(idx, act, R, 40, 256, variant=2)


@pytest.mark.parametrize("variant", [0, 1])
def test_float32_inputs_and_outputs(variant):
    idx, act, R = synth_streams(40, 3000, 20, 11, seed=99, r_dtype=np.float32)
    ref = _oracle(idx, act, R.astype(np.float64), 20, 11)           # same inputs, widened
    res = dc.stream_bounds(idx, act, R, 20, 11, variant=variant)       # float32 in -> float32 out
    assert res.value.dtype == np.float32 and res.overall.dtype == np.float32
    compare_stream(res, ref, value_tol=1e-5)
    res64 = dc.stream_bounds(idx, act, R, 20, 11, out_dtype=torch.float64, variant=variant)
    compare_stream(res64, ref, value_tol=1e-9)
    no_ov = dc.stream_bounds(idx, act, R, 20, 11, want_overall=False, variant=variant)
    assert no_ov.overall is None and np.array_equal(no_ov.act, res.act) and np.array_equal(no_ov.value, res.value)


def test_compact_wire_format_matches_int32_keys():
    """dcarl_stream_bounds_u8: one-byte keys in, one-byte argmax out — same bits as the int32 format."""
    for (n_inst, n_steps, S, A, kw) in [(40, 3000, 20, 11, {}), (9, 777, 1, 30, {}), (5, 1200, 3, 200, {"rule_act": 7}),
                                        (6, 800, 4, 6, {"prior_rule": -80.0})]:
        idx, act, R = synth_streams(n_inst, n_steps, S, A, seed=31 + S, r_dtype=np.float32)
        p = _abi.Params(**kw)
        wide = dc.stream_bounds(idx, act, R, S, A, p, want_stats=True)
        for ov in (True, False):
            c = dc.stream_bounds(idx.astype(np.uint8), act.astype(np.uint8), R, S, A, p, key_dtype=torch.uint8,
                                 want_overall=ov, want_stats=True)
            assert c.act.dtype == np.uint8 and c.value.dtype == np.float32
            assert np.array_equal(c.act, wide.act) and np.array_equal(c.value, wide.value)
            assert np.array_equal(c.activation_step, wide.activation_step)
            assert np.array_equal(c.act_hist, wide.act_hist)
            if ov:
                assert np.array_equal(c.overall, wide.overall)
    # host path of a plan with one-byte keys: pinned staging carries 6 + 9 bytes per step
    idx, act, R = synth_streams(64, 2048, 20, 11, seed=8, r_dtype=np.float32)
    plan = dc.StreamPlan(64, 2048, 20, 11, key_dtype=torch.uint8)
    assert plan.h2d_bytes() == 64 * 2048 * 6 and plan.d2h_bytes() == 64 * 2048 * 9 + 64 * 20 * 8
    h = plan.run_host(idx, act, R)
    ref = dc.stream_bounds(idx, act, R, 20, 11)
    assert np.array_equal(h.act, ref.act) and np.array_equal(h.value, ref.value) and np.array_equal(h.overall, ref.overall)
    with pytest.raises(ValueError):
        dc.StreamPlan(4, 100, 300, 4, key_dtype=torch.uint8)             # S does not fit one byte
    with pytest.raises(IndexError):
        dc.stream_bounds(np.full((2, 50), 300), np.zeros((2, 50), np.int64), np.zeros((2, 50), np.float32), 20, 11,
                         key_dtype=torch.uint8)


def test_bad_rows_are_counted_per_row():
    idx, act, R = synth_streams(3, 500, 4, 5, seed=2, r_dtype=np.float32)
    idx[0, 7] = 4; act[0, 7] = 9            # one row, both keys bad: still one row
    idx[1, 33:40] = -1                      # seven rows of one warp batch
    act[2, 499] = 5                         # the last row of the ragged tail
    for variant in (0, 1):
        plan = dc.StreamPlan(3, 500, 4, 5, variant=variant)
        plan.run_device(*[torch.from_numpy(x).cuda() for x in (idx, act, R)])
        assert int(plan.bad_rows.item()) == 9
        with pytest.raises(IndexError, match="9 rows"):
            plan.raise_on_bad_rows()


def test_constant_and_nearly_constant_cells():
    """A cell whose returns are all equal has sigma == 0 exactly in the reference (two-pass np.std); the
    kernel's one-pass variance must not turn rounding residue of s2/n - mean^2 into a spurious sigma."""
    rng = np.random.default_rng(11)
    n_inst, n_steps, S, A = 8, 1500, 2, 4
    idx = rng.integers(0, S, (n_inst, n_steps)).astype(np.int32)
    act = rng.integers(0, A, (n_inst, n_steps)).astype(np.int32)
    base = np.array([[37.3, 91.7, -12.9, 64.123456789], [1e3 / 7, 55.5, 80.1, 80.1]])   # non-dyadic constants
    R = base[idx, act]
    R[:, ::97] += 1e-3                                                   # a few rows make some cells nearly constant
    for variant in (0, 1):
        res = dc.stream_bounds(idx, act, R, S, A, _abi.Params(n_thres=3), variant=variant)
        # nearly constant cells: the one-pass variance resolves sigma only down to ~sqrt(n) 1e-8 |mean|,
        # which moves the bound by less than 1e-6 (the path's tolerance is 1e-5)
        compare_stream(res, _oracle(idx, act, R, S, A, n_thres=3), value_tol=1e-6, max_near_ties=2)
    Rc = base[idx, act]                                                  # exactly constant cells: sigma = 0 exactly
    for variant in (0, 1):
        res = dc.stream_bounds(idx, act, Rc, S, A, _abi.Params(n_thres=3), variant=variant)
        compare_stream(res, _oracle(idx, act, Rc, S, A, n_thres=3), value_tol=1e-11, max_near_ties=2)
    # exactly constant: the closed form of the bound with sigma = 0
    n = np.arange(4, 400, dtype=np.int32)
    c = 64.123456789
    s1 = np.cumsum(np.full(400, c))[n - 1]
    s2 = np.cumsum(np.full(400, c * c))[n - 1]
    out = torch.empty(len(n), dtype=torch.float64, device="cuda")
    t = lambda a, d: torch.from_numpy(np.ascontiguousarray(a)).to("cuda", d)
    dn, d1, d2, da = t(n, torch.int32), t(s1, torch.float64), t(s2, torch.float64), t(np.ones(len(n)), torch.int32)
    _abi.check(_abi.lib().dcarl_cell_values(dc._dev.ptr(dn), dc._dev.ptr(d1), dc._dev.ptr(d2), dc._dev.ptr(da),
                                            len(n), _abi.Params(), dc._dev.ptr(out), dc._dev.stream_handle()))
    torch.cuda.synchronize()
    ch = 150.0 * np.sqrt(np.log(20.0) / 2)
    want = np.minimum(c - ch / np.sqrt(n), c - ch / np.sqrt(n + 1.0))   # min(lower, CI lower) with sigma = 0
    assert np.abs(out.cpu().numpy() - want).max() < 1e-11


def test_integer_returns_make_exact_ties():
    rng = np.random.default_rng(3)
    n_inst, n_steps, S, A = 16, 1500, 3, 5
    idx = rng.integers(0, S, (n_inst, n_steps)).astype(np.int32)
    act = rng.integers(0, A, (n_inst, n_steps)).astype(np.int32)
    R = np.round(rng.normal(40, 3, (n_inst, n_steps)) / 20) * 20       # few distinct values -> equal cells
    for variant in (0, 1):
        res = dc.stream_bounds(idx, act, R, S, A, _abi.Params(n_thres=2), variant=variant)
        compare_stream(res, _oracle(idx, act, R, S, A, n_thres=2), value_tol=1e-9, max_near_ties=0)


def test_two_kernels_agree_bitwise_on_indices():
    idx, act, R = synth_streams(64, 4000, 20, 11, seed=5, r_dtype=np.float32)
    a = dc.stream_bounds(idx, act, R, 20, 11, variant=2, out_dtype=torch.float64)
    b = dc.stream_bounds(idx, act, R, 20, 11, variant=1, out_dtype=torch.float64)
    assert np.array_equal(a.act, b.act) and np.array_equal(a.activation_step, b.activation_step)
    assert np.abs(a.value - b.value).max() < 1e-10 and np.abs(a.overall - b.overall).max() < 1e-8


def test_empty_and_invalid_inputs():
    z = np.zeros((3, 0), np.int32)
    res = dc.stream_bounds(z, z, np.zeros((3, 0)), 4, 5)
    assert res.value.shape == (3, 0) and (res.activation_step == -1).all()
    res = dc.stream_bounds(np.zeros((0, 10), np.int32), np.zeros((0, 10), np.int32), np.zeros((0, 10)), 4, 5)
    assert res.value.shape == (0, 10)
    idx, act, R = synth_streams(2, 100, 4, 5, seed=1, r_dtype=np.float64)
    idx[1, 50] = 4                                                      # == S: out of range
    with pytest.raises(IndexError):
        dc.stream_bounds(idx, act, R, 4, 5)
    with pytest.raises(dc.DcarlError):
        dc.stream_bounds(idx, act, R, 4, 5, _abi.Params(rule_act=9))


def test_device_tensors_in_device_tensors_out_and_stats():
    idx, act, R = synth_streams(96, 2048, 20, 11, seed=21, r_dtype=np.float32)
    plan = dc.StreamPlan(96, 2048, 20, 11)
    d = [torch.from_numpy(x).cuda() for x in (idx, act, R)]
    res = plan.run_device(*d)
    assert res.value.is_cuda and res.act.dtype == torch.int32
    torch.cuda.synchronize()
    v, a, o, st = res.value.cpu().numpy(), res.act.cpu().numpy(), res.overall.cpu().numpy(), res.activation_step.cpu().numpy()
    want = stats_reference(v, o, st, a, 11)
    got = res.stats.cpu().numpy()
    assert np.array_equal(got[2:], want[2:])                            # counts and histogram: exact
    assert np.allclose(got[:2], want[:2], rtol=1e-12)
    assert np.array_equal(res.act_hist.cpu().numpy(), np.bincount(a.ravel(), minlength=11))
    # the pinned host path returns the same bits
    h = plan.run_host(idx, act, R)
    assert np.array_equal(h.value, v) and np.array_equal(h.act, a) and np.array_equal(h.overall, o)
    # a second run of the same plan is bitwise reproducible (no order-dependent atomics on outputs)
    res2 = plan.run_device(*d)
    torch.cuda.synchronize()
    assert torch.equal(res2.value, res.value) and torch.equal(res2.stats, res.stats)


def test_full_size_config2_properties_and_sampled_parity():
    """BASELINE config 2: 4096 instances x 20 000 steps, S=20, A=11, float32 rows."""
    n_inst, n_steps, S, A = 4096, 20000, 20, 11
    g = torch.Generator(device="cuda").manual_seed(2)
    q = torch.rand((n_inst, S, A), generator=g, device="cuda") * 150 - 50
    idx = torch.floor((3 + torch.randn((n_inst, n_steps), generator=g, device="cuda")) / 6 * S).clamp_(0, S - 1).to(torch.int32)
    act = torch.randint(0, A, (n_inst, n_steps), generator=g, device="cuda", dtype=torch.int32)
    R = q.reshape(n_inst, -1).gather(1, (idx.long() * A + act.long())) + 50 * torch.randn((n_inst, n_steps), generator=g, device="cuda")
    del q
    plan = dc.StreamPlan(n_inst, n_steps, S, A)
    res = plan.run_device(idx, act, R)
    torch.cuda.synchronize()
    plan.raise_on_bad_rows()
    a = res.act
    assert int(a.min()) >= 0 and int(a.max()) < A
    assert torch.equal(res.act_hist, torch.bincount(a.reshape(-1).long(), minlength=A))
    assert bool(torch.isfinite(res.value).all()) and bool(torch.isfinite(res.overall).all())
    # sharding invariance: the same instances run as two half batches give the same bits
    half = dc.StreamPlan(n_inst // 2, n_steps, S, A)
    for lo in (0, n_inst // 2):
        sl = slice(lo, lo + n_inst // 2)
        r2 = half.run_device(idx[sl].contiguous(), act[sl].contiguous(), R[sl].contiguous())
        torch.cuda.synchronize()
        assert torch.equal(r2.value, res.value[sl]) and torch.equal(r2.act, res.act[sl])
        assert torch.equal(r2.overall, res.overall[sl]) and torch.equal(r2.activation_step, res.activation_step[sl])
    # sampled instances against the oracle, and the trace-derived consistency on a few
    pick = np.random.default_rng(0).choice(n_inst, 24, replace=False)
    pi = torch.from_numpy(pick).cuda()
    hi, ha, hr = idx[pi].cpu().numpy(), act[pi].cpu().numpy(), R[pi].cpu().numpy()
    class _R:
        pass
    sub = _R()
    sub.value, sub.act = res.value[pi].cpu().numpy(), res.act[pi].cpu().numpy()
    sub.overall, sub.activation_step = res.overall[pi].cpu().numpy(), res.activation_step[pi].cpu().numpy()
    compare_stream(sub, _oracle(hi, ha, hr.astype(np.float64), S, A), value_tol=1e-5, max_near_ties=1)
    for j in range(3):
        astep, ov = recompute_from_traces(hi[j], sub.value[j].astype(np.float64), sub.act[j], S)
        assert np.array_equal(astep, sub.activation_step[j])
        assert np.abs(ov - sub.overall[j]).max() < 1e-2 * 1e-1   # float32 traces: 20 states x ~6e-6 each


def test_config5_mixed_sim1_sim2_batch_reduced():
    """BASELINE config 5 at reduced size: even instances Simulation_1-shaped (S=1, A=30, 11 live
    actions), odd instances Simulation_2-shaped (S=20, A=11), 500 steps; one merged stats vector."""
    n_each, n_steps = 96, 500
    i1, a1, R1 = synth_streams(n_each, n_steps, 1, 30, seed=4000, r_dtype=np.float32, live_actions=11)
    i2, a2, R2 = synth_streams(n_each, n_steps, 20, 11, seed=4001, r_dtype=np.float32)
    p1 = dc.StreamPlan(n_each, n_steps, 1, 30)
    p2 = dc.StreamPlan(n_each, n_steps, 20, 11)
    r1 = p1.run_device(*[torch.from_numpy(x).cuda() for x in (i1, a1, R1)])
    r2 = p2.run_device(*[torch.from_numpy(x).cuda() for x in (i2, a2, R2)])
    merged = dc.merge_stats([r1.stats, r2.stats]).cpu().numpy()
    want = np.zeros(5 + 30)
    for (ii, aa, RR, S, A, res) in ((i1, a1, R1, 1, 30, r1), (i2, a2, R2, 20, 11, r2)):
        ref = _oracle(ii, aa, RR.astype(np.float64), S, A)
        class _R:
            pass
        h = _R()
        h.value, h.act = res.value.cpu().numpy(), res.act.cpu().numpy()
        h.overall, h.activation_step = res.overall.cpu().numpy(), res.activation_step.cpu().numpy()
        compare_stream(h, ref, value_tol=1e-5)
        st = stats_reference(ref["value"].astype(np.float32), ref["overall"].astype(np.float32),
                             ref["activation_step"], ref["act"], A)
        want[:5 + A] += st
    assert np.array_equal(merged[2:], want[2:])
    assert np.allclose(merged[:2], want[:2], rtol=1e-6)
