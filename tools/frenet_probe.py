#!/usr/bin/env python
"""k_frenet_plan on 262 144 scenarios, a few launches (for a light ncu pass; profiles/README.md)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench, dcarl_b200 as dc
dev = torch.device("cuda:0")
n, n_paths = 262144, 64
knots, off, pid, start, ego, obs, n_obs = bench.synth_frenet(n, n_paths, 5000)
routes = dc.Routes(knots)
p = dc.frenet_params()
t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
d = (t(pid, torch.int32), t(start, torch.float64), t(ego, torch.float64), t(obs, torch.float64), t(n_obs, torch.int32))
for _ in range(3):
    res = dc.frenet_plan(routes, *d, p)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    res = dc.frenet_plan(routes, *d, p)
e1.record()
torch.cuda.synchronize()
print("k_frenet_plan", e0.elapsed_time(e1) / 10, "ms per", n, "scenarios; picks", int((res.index > 0).sum()))
