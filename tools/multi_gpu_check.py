#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/multi_gpu_check.py — N-GPU run of the two sharding patterns
 against a single-GPU run of the same data on rank 0:
  (i)  instance-sharded stream (configs 2/5): per-scenario outputs bit-identical, stats all-reduced;
  (ii) row-sharded box moments (config 4): summed raw moments, identical picks.
Also times config 4 (1e7 rows / N per GPU) and config 5 (32768 mixed instances x 500 steps per GPU)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import dcarl_b200 as dc  # noqa: E402
from dcarl_b200 import dist as dd  # noqa: E402

rank, world, local = dd.init_from_env()
dev = torch.device(f"cuda:{local}")
out = {"world": world}


def timed(fn, reps):
    fn(); fn()
    dd.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    dd.allreduce_max(t)
    return float(t[0])


# ---- (i) stream, instance-sharded ------------------------------------------------------------
S, A, n_inst, n_steps = 20, 11, 512, 3000
g = torch.Generator(device=dev).manual_seed(123)           # same seed on every rank: same global batch
idx = torch.randint(0, S, (n_inst, n_steps), generator=g, device=dev, dtype=torch.int32)
act = torch.randint(0, A, (n_inst, n_steps), generator=g, device=dev, dtype=torch.int32)
R = torch.randn((n_inst, n_steps), generator=g, device=dev) * 50 + 20
lo, hi = dd.shard_range(n_inst, rank, world)
plan = dc.StreamPlan(hi - lo, n_steps, S, A)
res, stats = dd.stream_sharded(plan, idx[lo:hi].contiguous(), act[lo:hi].contiguous(), R[lo:hi].contiguous())
full = dc.StreamPlan(n_inst, n_steps, S, A).run_device(idx, act, R)
torch.cuda.synchronize()
ok = bool(torch.equal(res.value, full.value[lo:hi]) and torch.equal(res.act, full.act[lo:hi])
          and torch.equal(res.overall, full.overall[lo:hi]) and torch.equal(res.activation_step, full.activation_step[lo:hi]))
rel = float(((stats - full.stats).abs() / full.stats.abs().clamp_min(1)).max())
exact_counts = bool(torch.equal(stats[2:], full.stats[2:]))
flag = torch.tensor([float(ok and exact_counts and rel < 1e-9)], device=dev, dtype=torch.float64)
torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN) if world > 1 else None
out["stream_sharded_identical"] = bool(flag.item() == 1.0)
out["stream_stats_max_rel_diff"] = rel

# ---- (ii) box moments, row-sharded ---------------------------------------------------------------
M, NA, Q = 400_000, 64, 8
dist = torch.tensor(dc.VISITED_STATE_DIST[0], dtype=torch.float64, device=dev)
centers = torch.randn((64, 20), generator=g, device=dev) * dist[:20].float()
st = centers[torch.randint(0, 64, (M,), generator=g, device=dev)] + torch.randn((M, 20), generator=g, device=dev) * (0.25 * dist[:20].float())
rows = torch.cat([st, torch.randint(0, NA, (M, 1), generator=g, device=dev).float()], 1)
ret = -torch.rand(M, generator=g, device=dev)
q = centers[:Q].double().contiguous()
lo, hi = dd.shard_range(M, rank, world)
shard = dc.BoxBuffer(20, capacity=hi - lo)
shard.append(rows[lo:hi], ret[lo:hi])
rp = dc.RlsParams(visited_times_thres=10)
mom, mv, pick = dd.box_query_sharded(shard, q, NA, rp)            # scan + all-reduce + select (three steps)
fq = dd.ShardedBoxQuery(shard, Q, NA, rp)                          # one launch, sum over ranks through peer memory
for _ in range(3):                                                 # repeated calls: the workspace must come back zeroed
    fmom, fmv, fpick = fq(q)
torch.cuda.synchronize()
out["box_fused_path"] = fq.path
out["box_fused_equals_unfused"] = bool(torch.equal(fmom[:, :, 0], mom[:, :, 0]) and torch.equal(fpick, pick)
                                       and float((fmom - mom).abs().max()) < 1e-9)
# the peer-memory sum against NCCL on random vectors, twenty calls back to back (both parities, epochs)
pg = dd.PeerGroup(4096)
peer_ok = True
for it in range(20):
    v = torch.randn(1000 + it, generator=torch.Generator(device=dev).manual_seed(100 * rank + it), device=dev, dtype=torch.float64)
    w = v.clone()
    pg.allreduce(v)
    torch.distributed.all_reduce(w) if world > 1 else None
    peer_ok = peer_ok and float((v - w).abs().max()) < 1e-12
gath = dd.allgather(v)
out["peer_backend"] = pg.backend
out["peer_allreduce_matches_nccl"] = bool(peer_ok)
out["peer_allreduce_bitwise_equal_on_all_ranks"] = bool(all(torch.equal(gath[0], x) for x in gath))
whole = dc.BoxBuffer(20, capacity=M)
whole.append(rows, ret)
mom1 = whole.moments(q, NA)
mv1, pick1 = whole.select(mom1, rp)
torch.cuda.synchronize()
out["box_counts_identical"] = bool(torch.equal(mom[:, :, 0], mom1[:, :, 0]))
out["box_sums_max_rel_diff"] = float(((mom - mom1).abs() / mom1.abs().clamp_min(1e-9)).max())
out["box_picks_identical"] = bool(torch.equal(pick, pick1))

# ---- (iii) planner (row f3), scenario-sharded ------------------------------------------------------
import bench  # noqa: E402  (synthetic planner inputs: same seed on every rank = same global batch)
n_sc = 20000
knots, off, pid, start_unused, ego_xy, obs, n_obs = bench.synth_frenet(n_sc, 16, 6000)
rng = np.random.default_rng(6001)
ego = np.concatenate([ego_xy, rng.normal(0, 4, (n_sc, 2)), rng.uniform(0.2, 12, (n_sc, 1))], axis=1)   # x, y, vx, vy, v
routes = dc.Routes(knots)
tt = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
lo, hi = dd.shard_range(n_sc, rank, world)
mine, st_f = dd.frenet_sharded(routes, tt(pid[lo:hi], torch.int32), tt(ego[lo:hi], torch.float64), tt(obs[lo:hi], torch.float64),
                               tt(n_obs[lo:hi], torch.int32))
d_ego = tt(ego, torch.float64)
whole_f = dc.frenet_plan(routes, tt(pid, torch.int32), dc.frenet_start(routes, tt(pid, torch.int32), d_ego), d_ego[:, :2].contiguous(),
                         tt(obs, torch.float64), tt(n_obs, torch.int32))
want = dd.frenet_stats(whole_f.index, whole_f.cost)
torch.cuda.synchronize()
ok = bool(torch.equal(mine.index, whole_f.index[lo:hi]) and torch.equal(mine.cost, whole_f.cost[lo:hi])
          and torch.equal(mine.traj_len, whole_f.traj_len[lo:hi]))
ok = ok and bool(torch.equal(st_f[:2], want[:2]) and torch.equal(st_f[3:], want[3:])) and \
    float((st_f[2] - want[2]).abs() / want[2].abs().clamp_min(1)) < 1e-9
flag = torch.tensor([float(ok)], device=dev, dtype=torch.float64)
torch.distributed.all_reduce(flag, op=torch.distributed.ReduceOp.MIN) if world > 1 else None
out["frenet_sharded_identical"] = bool(flag.item() == 1.0)

# ---- timings: config 4 (1e7 rows over all GPUs) and config 5 (mixed, 32768 instances per GPU) ------
if os.environ.get("TIMINGS", "1") == "1":
    M4 = 10_000_000 // world
    b4 = dc.BoxBuffer(20, capacity=M4)
    st = centers[torch.randint(0, 64, (M4,), generator=g, device=dev)] + torch.randn((M4, 20), generator=g, device=dev) * (0.25 * dist[:20].float())
    b4.append(torch.cat([st, torch.randint(0, 256, (M4, 1), generator=g, device=dev).float()], 1), -torch.rand(M4, generator=g, device=dev))
    del st
    q1 = centers[:1].double().contiguous()
    ms = timed(lambda: dd.box_query_sharded(b4, q1, 256), 50)
    f4 = dd.ShardedBoxQuery(b4, 1, 256)
    ms_f = timed(lambda: f4(q1), 50)
    out["config4"] = {"rows_total": M4 * world, "ms_per_query_pass": ms_f, "rows_per_s": M4 * world / (ms_f * 1e-3),
                      "achieved_GBs_per_gpu": M4 * 88 / (ms_f * 1e-3) / 1e9, "path": f4.path,
                      "ms_three_step_path": ms, "three_step_path": "scan + NCCL all-reduce of 256x3 float64 moments + select"}
    del b4
    ne, ns = 16384, 500                                     # per GPU: 16384 Sim1-shaped + 16384 Sim2-shaped
    i1 = torch.zeros((ne, ns), dtype=torch.int32, device=dev)
    a1 = torch.randint(0, 11, (ne, ns), generator=g, device=dev, dtype=torch.int32)
    R1 = torch.randn((ne, ns), generator=g, device=dev) * 50 + 30
    i2 = torch.floor((3 + torch.randn((ne, ns), generator=g, device=dev).clamp_(-2.99, 2.99)) / 6 * 20).to(torch.int32)
    a2 = torch.randint(0, 11, (ne, ns), generator=g, device=dev, dtype=torch.int32)
    R2 = torch.randn((ne, ns), generator=g, device=dev) * 50 + 30
    p1, p2 = dc.StreamPlan(ne, ns, 1, 30), dc.StreamPlan(ne, ns, 20, 11)

    def c5():
        r1 = p1.run_device(i1, a1, R1)
        r2 = p2.run_device(i2, a2, R2)
        return dd.allreduce_stats(dc.merge_stats([r1.stats, r2.stats]))
    ms_old = timed(c5, 50)
    want5 = c5().clone()
    mb = dd.MixedStreamBatch([p1, p2], keep_local=False)
    ms = timed(lambda: mb.run([(i1, a1, R1), (i2, a2, R2)]), 50)
    got5 = mb.run([(i1, a1, R1), (i2, a2, R2)])
    torch.cuda.synchronize()
    out["config5_fused_stats_match"] = bool(torch.equal(got5[2:], want5[2:]) and
                                            float(((got5[:2] - want5[:2]).abs() / want5[:2].abs().clamp_min(1)).max()) < 1e-12)
    out["config5_ms_unfused"] = ms_old
    steps = 2 * ne * ns * world
    out["config5"] = {"scenario_steps_total": steps, "ms_per_batch": ms, "scenario_steps_per_s": steps / (ms * 1e-3),
                      "achieved_GBs_per_gpu": 2 * ne * ns * 24 / (ms * 1e-3) / 1e9}
if rank == 0:
    print(json.dumps(out), flush=True)
if world > 1:
    torch.distributed.destroy_process_group()
