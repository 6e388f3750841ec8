"""world_size-2 gloo test of the multi-GPU host logic: contiguous instance shards,
no data-path exchange, one SUM all-reduce of the float64 stats vector.  Per-shard results come
from the oracle here (no GPU in this container); on the GPU box the same code path runs over NCCL."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as tdist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dcarl_b200 import dist as ddist
    from helpers import stats_reference, synth_streams
    from oracle import cport
    r, w, _ = ddist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    n_inst, n_steps, S, A = 10, 400, 20, 11          # 10 instances over 2 ranks and (below) 3 shards
    idx, act, R = synth_streams(n_inst, n_steps, S, A, seed=77, r_dtype=np.float64)
    lo, hi = ddist.shard_range(n_inst, rank, world)
    out = cport.stream(idx[lo:hi], act[lo:hi], R[lo:hi], S, A)
    st = torch.from_numpy(stats_reference(out["value"], out["overall"], out["activation_step"], out["act"], A))
    ddist.allreduce_stats(st)
    t = torch.tensor([float(rank + 1)], dtype=torch.float64)
    ddist.allreduce_max(t)
    ddist.barrier()
    q.put((rank, lo, hi, st.numpy(), float(t[0]), out["act"]))
    tdist.destroy_process_group()


def test_shard_range_partitions():
    from dcarl_b200.dist import shard_range
    for n, w in [(10, 3), (4096, 8), (7, 8), (0, 2), (262144, 8)]:
        parts = [shard_range(n, r, w) for r in range(w)]
        assert parts[0][0] == 0 and parts[-1][1] == n
        assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
        sizes = [b - a for a, b in parts]
        assert max(sizes) - min(sizes) <= 1


def test_two_rank_gloo_allreduce_matches_single_process():
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from helpers import stats_reference, synth_streams
    from oracle import cport
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    got = sorted((q.get(timeout=120) for _ in range(world)), key=lambda x: x[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    n_inst, n_steps, S, A = 10, 400, 20, 11
    idx, act, R = synth_streams(n_inst, n_steps, S, A, seed=77, r_dtype=np.float64)
    full = cport.stream(idx, act, R, S, A)
    want = stats_reference(full["value"], full["overall"], full["activation_step"], full["act"], A)
    assert got[0][1:3] == (0, 5) and got[1][1:3] == (5, 10)
    for rank, lo, hi, st, mx, acts in got:
        assert np.array_equal(st[2:], want[2:])                    # counts / histogram: exact
        assert np.allclose(st[:2], want[:2], rtol=1e-12, atol=0)   # float sums: order of two partials
        assert mx == 2.0
        assert np.array_equal(acts, full["act"][lo:hi])            # per-scenario outputs do not depend on the sharding
    assert np.array_equal(got[0][3], got[1][3])                    # every rank holds the same reduced vector


def test_row_sharded_box_moments_sum_to_the_unsharded_answer(golden):
    """raw moments (n, sum, sum^2) of row shards add up; the select epilogue on the
    summed moments equals the unsharded pick (host logic, oracle as the per-shard scanner)."""
    import numpy as np
    from dcarl_b200.dist import shard_range
    from oracle import cport
    g = golden("rls_box")
    vs, ret, dr, dist = g["visited_state"], g["visited_value"][:, 1], g["driving_record"], g["dist"]
    q = dr[::7, :20]
    full = cport.box_moments(vs, ret, dist, q, 8)

    def raw(nmv_rows_state, nmv_rows_ret):
        nmv = cport.box_moments(nmv_rows_state, nmv_rows_ret, dist, q, 8)
        n = nmv[:, :, 0]
        s1 = np.where(n > 0, nmv[:, :, 1] * n, 0.0)
        s2 = np.where(n > 0, (nmv[:, :, 2] + nmv[:, :, 1] ** 2) * n, 0.0)
        return np.stack([n, s1, s2], -1)
    tot = np.zeros((len(q), 8, 3))
    for r in range(8):
        lo, hi = shard_range(len(vs), r, 8)
        tot += raw(vs[lo:hi], ret[lo:hi])
    n = tot[:, :, 0]
    assert np.array_equal(n, full[:, :, 0])
    mean = np.where(n > 0, tot[:, :, 1] / np.maximum(n, 1), -1.0)
    var = np.where(n > 0, np.maximum(tot[:, :, 2] / np.maximum(n, 1) - mean ** 2, 0.0), -1.0)
    assert np.abs(mean - full[:, :, 1]).max() < 1e-12 and np.abs(var - full[:, :, 2]).max() < 1e-12
    assert np.array_equal(cport.box_select(np.stack([n, mean, var], -1)), cport.box_select(full))


def _frenet_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank),
                      MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    from dcarl_b200 import dist as ddist
    from helpers_frenet import synth_routes, synth_scenarios
    from oracle import cport
    ddist.init_from_env(backend="gloo")
    rng = np.random.default_rng(31)                      # same seed on every rank: same global batch
    knots, off = synth_routes(rng, 5)
    pid, start, ego, obs, n_obs = synth_scenarios(rng, 301, knots, off)
    coef = np.concatenate([cport.spline_build(knots[off[i]:off[i + 1]]) for i in range(5)])
    lo, hi = ddist.shard_range(301, rank, world)
    out = cport.frenet_plan(coef, off, pid[lo:hi], start[lo:hi], ego[lo:hi], obs[lo:hi], n_obs[lo:hi], cport.frenet_params())
    st = ddist.frenet_stats(torch.from_numpy(out["index"]), torch.from_numpy(out["cost"]))
    ddist.allreduce_stats(st)
    ddist.barrier()
    q.put((rank, lo, hi, st.numpy(), out["index"]))
    tdist.destroy_process_group()


def test_two_rank_planner_shards_and_stats_allreduce():
    """Row f3 on N > 1: scenarios block-partitioned, no exchange, one all-reduce of the tick's summary."""
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    from dcarl_b200 import dist as ddist
    from helpers_frenet import synth_routes, synth_scenarios
    from oracle import cport
    rng = np.random.default_rng(31)
    knots, off = synth_routes(rng, 5)
    pid, start, ego, obs, n_obs = synth_scenarios(rng, 301, knots, off)
    coef = np.concatenate([cport.spline_build(knots[off[i]:off[i + 1]]) for i in range(5)])
    whole = cport.frenet_plan(coef, off, pid, start, ego, obs, n_obs, cport.frenet_params())
    want = ddist.frenet_stats(torch.from_numpy(whole["index"]), torch.from_numpy(whole["cost"])).numpy()
    assert want[0] == 301 and want[3:].sum() == 301 and want[1] == (whole["index"] == 0).sum()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_frenet_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = sorted([q.get(timeout=120) for _ in range(2)])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert got[0][1] == 0 and got[0][2] == got[1][1] and got[1][2] == 301
    for _, lo, hi, st, index in got:
        assert np.array_equal(index, whole["index"][lo:hi])           # per-scenario results do not depend on the sharding
        assert np.array_equal(st[[0, 1]], want[[0, 1]]) and np.array_equal(st[3:], want[3:])   # counts: exact
        assert abs(st[2] - want[2]) <= 1e-9 * abs(want[2])
