#!/usr/bin/env python
"""Small invocations of every kernel, meant to run under compute-sanitizer:
    compute-sanitizer --tool memcheck  python tools/sanitize_cases.py
    compute-sanitizer --tool racecheck python tools/sanitize_cases.py
Results are also checked against the oracle so a silent corruption fails the run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import dcarl_b200 as dc  # noqa: E402
from helpers import compare_stream, synth_streams  # noqa: E402
from oracle import cport  # noqa: E402

for (n_inst, n_steps, S, A, kw) in [(9, 700, 20, 11, {}), (5, 333, 1, 30, {}), (3, 200, 3, 40, {}), (4, 300, 40, 5, {"rule_act": 3}),
                                    (4, 300, 4, 6, {"rule_act": 3, "prior_other": 100.0})]:
    idx, act, R = synth_streams(n_inst, n_steps, S, A, seed=n_steps, r_dtype=np.float64)
    ref = cport.stream(idx, act, R, S, A, cport.default_params(**kw))
    for variant in (0, 1):
        res = dc.stream_bounds(idx, act, R, S, A, dc.Params(**kw), variant=variant, want_stats=True)
        compare_stream(res, ref, value_tol=1e-9)
# the summary kernel as a cluster of 8 CTAs (>= 16384 instances)
idx, act, R = synth_streams(16400, 40, 20, 11, seed=5, r_dtype=np.float32)
res = dc.stream_bounds(idx, act, R, 20, 11, want_stats=True)
assert res.stats[4] == 16400 * 40 and res.stats[5:].sum() == 16400 * 40
print("stream ok")
rng = np.random.default_rng(0)
dist = dc.VISITED_STATE_DIST[0]
vs = np.hstack([rng.normal(0, 0.4, (3001, 20)) * dist[:20], rng.integers(0, 8, (3001, 1)).astype(np.float64)])
ret = -rng.random(3001)
buf = dc.BoxBuffer(20, capacity=64)
buf.append(vs, ret)
for Q in (1, 7):
    q = vs[:Q, :20]
    n, mean, var, pick = buf.query(q, 8)
    ref = cport.box_moments(vs.astype(np.float32).astype(np.float64), ret.astype(np.float32).astype(np.float64), dist, q, 8)
    assert np.array_equal(n, ref[:, :, 0])
run = buf.query_graph(1, 8)
run(vs[:1, :20])
# many queries: the bitmap-filtered scan (unfused and as the one-launch query, repeated: the accumulators alternate),
# the float64 storage mode, more queries than one filtered launch takes
from dcarl_b200 import dist as dd  # noqa: E402
q64 = np.vstack([vs[:60, :20], vs[:6, :20] + 1e3])
ref = cport.box_raw(vs.astype(np.float32), ret.astype(np.float32), dist, q64, 8)
assert np.array_equal(buf.moments(q64, 8).cpu().numpy()[:, :, 0], ref[:, :, 0])
fq = dd.ShardedBoxQuery(buf, 64, 8)
for _ in range(3):
    assert np.array_equal(fq(q64[:64])[0].cpu().numpy()[:, :, 0], ref[:64, :, 0])
fq = dd.ShardedBoxQuery(buf, 64, 40)                       # 64 x 40 bins: global atomics
assert np.array_equal(fq(q64[:64])[0].cpu().numpy()[:, :8, 0], ref[:64, :, 0])
ib = dc.BoxBuffer(20)                                    # sorted index + unsorted tail: one warp per query over its row range
ib.append(vs[:2500], ret[:2500])
ib.build_index()
ib.append(vs[2500:], ret[2500:])
assert np.array_equal(ib.moments_indexed(q64, 8).cpu().numpy()[:, :, 0], ref[:, :, 0])
b64 = dc.BoxBuffer(20, dtype=torch.float64)
b64.append(vs, ret)
assert np.array_equal(b64.moments(q64[:9], 8).cpu().numpy()[:, :, 0], cport.box_raw(vs, ret, dist, q64[:9], 8)[:, :, 0])
print("box ok")
x = [rng.normal(0, 50, n) for n in (1, 17, 300, 5000)]
got = dc.bounds_batch(x)
assert np.allclose(got[2], cport.bounds(x[2]), rtol=1e-12)
dc.cell_values([12, 40], [100.0, -50.0], [2000.0, 900.0], [0, 2])
z = rng.standard_normal(5000)
dc.sample_returns(rng.uniform(-50, 100, (20, 11)), z, rng.integers(0, 11, 5000), rng.standard_normal(5000))
dc.sample_returns_philox(torch.rand((20, 11), device="cuda"), 5000, 3)
s0 = rng.uniform(-5, 5, (1000, 20)).astype(np.float32)
dc.rollout_cv(s0, rng.integers(0, 11, 1000))
# Frenet planner: spline build + one planning tick, against the oracle
from helpers_frenet import synth_routes, synth_scenarios  # noqa: E402
knots, off = synth_routes(rng, 4)
pid, start, ego, obs, n_obs = synth_scenarios(rng, 300, knots, off, max_obs=5)
routes = dc.Routes([knots[off[i]:off[i + 1]] for i in range(4)])
res = dc.frenet_plan(routes, pid, start, ego, obs, n_obs, want_paths=True)
coef = np.concatenate([cport.spline_build(knots[off[i]:off[i + 1]]) for i in range(4)])
ref = cport.frenet_plan(coef, off, pid, start, ego, obs, n_obs, cport.frenet_params())
assert np.array_equal(res.index, ref["index"])
print("frenet ok")
torch.cuda.synchronize()
print("all kernels ran")
