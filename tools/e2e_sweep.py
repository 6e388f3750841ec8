#!/usr/bin/env python
"""Host-in / host-out time of StreamPlan.run_pinned on config 2 for several chunk counts."""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
import dcarl_b200 as dc

dev = torch.device("cuda:0")
idx, act, R = bench.synth_device(bench.N_INST, bench.N_STEPS, 1000, dev)
key = torch.uint8 if os.environ.get("KEYS", "u8") == "u8" else torch.int32     # KEYS=i32: the int32 wire format
plan = dc.StreamPlan(bench.N_INST, bench.N_STEPS, bench.S, bench.A, key_dtype=key)
pin = plan.pinned()
pin["idx"].copy_(idx.to(key)); pin["act"].copy_(act.to(key)); pin["R"].copy_(R)
torch.cuda.synchronize()
for chunks in [int(c) for c in (sys.argv[1] if len(sys.argv) > 1 else "4,8,16,32").split(",")]:
    plan.run_pinned(chunks=chunks)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(5):
        plan.run_pinned(chunks=chunks)
    torch.cuda.synchronize()
    ms = (time.perf_counter() - t0) * 1e3 / 5
    gb = (plan.h2d_bytes() + plan.d2h_bytes()) / 1e9
    print(f"chunks={chunks:3d}  {ms:7.2f} ms  {bench.N_INST * bench.N_STEPS / ms / 1e6:7.2f} G steps/s  {gb / ms * 1e3:6.1f} GB/s both ways")
