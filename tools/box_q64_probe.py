#!/usr/bin/env python
"""Config-4 buffer (1e7 rows, 256 actions), Q = 64 and Q = 1 fused query launches: a few calls for ncu."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import dcarl_b200 as dc  # noqa: E402
from dcarl_b200 import dist as dd  # noqa: E402

dev = torch.device("cuda:0")
chunks = int(os.environ.get("CHUNKS", bench.C4_CHUNKS))        # fewer chunks = a smaller buffer (1.25e6 rows each)
buf = dc.BoxBuffer(20, capacity=bench.C4_ROWS)
for c in range(chunks):
    centers, st, ret = bench.c4_chunk(c, dev)
    buf.append(st, ret)
    del st, ret
for Q in (64, 1):
    print("Q", Q, "rows", len(buf), flush=True)
    fq = dd.ShardedBoxQuery(buf, Q, 256)
    q = centers[:Q].double().contiguous()
    for _ in range(3):
        fq(q)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        fq(q)
    e1.record()
    torch.cuda.synchronize()
    print(Q, e0.elapsed_time(e1) / 5, "ms fused (scan + totals + select)")
    buf.moments(q, 256)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(5):
        buf.moments(q, 256)
    e1.record()
    torch.cuda.synchronize()
    print(Q, e0.elapsed_time(e1) / 5, "ms unfused scan (memset + thresholds + scan)")
