#!/usr/bin/env python
"""Debug: per-instance start/end times, SM and hardware warp slot of k_stream_rows on config 2.
Needs the library built with `make -C dcarl_b200/csrc EXTRA=-DDCARL_DEBUG_TIMING`."""
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import dcarl_b200 as dc
from dcarl_b200 import _abi

n_inst, n_steps = 4096, 20000
dev = torch.device("cuda:0")
idx, act, R = bench.synth_device(n_inst, n_steps, 1234, dev)
plan = dc.StreamPlan(n_inst, n_steps, bench.S, bench.A, want_stats=False, variant=0, device=dev)
for _ in range(3):
    plan.run_device(idx, act, R)
torch.cuda.synchronize()
buf = np.zeros((n_inst, 4), np.uint64)
L = _abi.lib()
L.dcarl_debug_rows_timing.argtypes = [C.c_void_p, C.c_int64]
assert L.dcarl_debug_rows_timing(buf.ctypes.data, n_inst) == 0
t0 = buf[:, 0].min()
st = (buf[:, 0] - t0).astype(np.float64) / 1e3
en = (buf[:, 1] - t0).astype(np.float64) / 1e3
sm = buf[:, 2].astype(np.int64)
wid = buf[:, 3].astype(np.int64)
dur = en - st
print(f"kernel span {en.max():.1f} us; start spread {st.max():.1f} us; "
      f"warp duration mean {dur.mean():.1f} min {dur.min():.1f} max {dur.max():.1f} p5 {np.percentile(dur,5):.1f} p95 {np.percentile(dur,95):.1f}")
print(f"end time: mean {en.mean():.1f} p5 {np.percentile(en,5):.1f} p50 {np.percentile(en,50):.1f} p95 {np.percentile(en,95):.1f} max {en.max():.1f}")
per_sm = np.bincount(sm, minlength=148)
print("warps per SM histogram:", np.bincount(per_sm).tolist())
# per (sm, smsp) warp counts, smsp = hw warp id % 4
key = sm * 4 + (wid % 4)
cnt = np.bincount(key, minlength=148 * 4)
print("warps per SMSP histogram (hw warpid % 4):", np.bincount(cnt).tolist())
# finish time by SMSP load
for c in sorted(set(cnt.tolist())):
    sel = np.isin(key, np.nonzero(cnt == c)[0])
    if sel.any():
        print(f"  SMSPs with {c} warps: {int((cnt==c).sum())} smsp, mean end {en[sel].mean():.1f} us, max end {en[sel].max():.1f}, mean dur {dur[sel].mean():.1f}")
sm_end = np.array([en[sm == s].max() if (sm == s).any() else 0 for s in range(148)])
print(f"per-SM last finish: min {sm_end[sm_end>0].min():.1f} mean {sm_end[sm_end>0].mean():.1f} max {sm_end.max():.1f}")
