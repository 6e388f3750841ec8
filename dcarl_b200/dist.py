"""Multi-GPU plumbing: scenarios shard across ranks with no data-path exchange; one small
all-reduce of the float64 summary-statistics vector per batch.

One process per GPU (torchrun).  NCCL on GPUs, gloo for the CPU tests of this host logic.
"""
import os

import torch
import torch.distributed as dist


def init_from_env(backend=None):
    """Initialise torch.distributed from RANK / WORLD_SIZE / MASTER_* if a launcher set them.
    Returns (rank, world_size, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend, device_id=torch.device(f"cuda:{local}"))
        else:
            dist.init_process_group(backend)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def shard_range(n_total, rank, world):
    """Contiguous block partition: rank r owns [lo, hi); sizes differ by at most one."""
    base, rem = divmod(int(n_total), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def allreduce_stats(stats, group=None):
    """In-place SUM all-reduce of the float64 stats vector (6 KB at most): latency-bound, one call
    per batch, issued on the compute stream."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM, group=group)
    return stats


def allreduce_max(t, group=None):
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return t


def barrier():
    if dist.is_available() and dist.is_initialized():
        dist.barrier()


def box_query_sharded(buf, queries, n_actions, rls_params=None, group=None):
    """Config-4 pattern: the {state, action, return} buffer is ROW-sharded (each rank's ``buf`` holds
    M/G rows), the queries are replicated.  Every rank scans its shard once, the raw moments
    (n, sum R, sum R^2) [Q, n_actions, 3] are summed across ranks with ONE all-reduce, and every
    rank runs the small select epilogue redundantly (no broadcast).  Returns CUDA tensors
    (moments, mean_var, action)."""
    mom = buf.moments(queries, n_actions)
    allreduce_stats(mom, group)
    mv, pick = buf.select(mom, rls_params)
    return mom, mv, pick


def stream_sharded(plan, idx, act, R, group=None):
    """Config-2/5 pattern: scenario instances are block-partitioned over ranks (``plan`` and the
    inputs are this rank's shard, see ``shard_range``); no data-path exchange, one all-reduce of the
    summary statistics.  Returns (this rank's StreamResult, all-reduced stats)."""
    res = plan.run_device(idx, act, R)
    allreduce_stats(res.stats, group)
    return res, res.stats


def frenet_stats(index, cost):
    """Summary vector of one planning tick, float64 [3 + C + 1]: scenarios, brake ticks (index 0), summed
    cost of the selected candidates, histogram of the selected index 0..C.  Tensor ops only (any device)."""
    idx = index.long()
    C = cost.shape[1]
    hist = torch.bincount(idx.clamp_min(0), minlength=C + 1)[:C + 1].to(torch.float64)
    picked = idx > 0
    sel_cost = cost.gather(1, (idx - 1).clamp_min(0)[:, None])[:, 0].to(torch.float64)
    head = torch.stack([torch.tensor(float(idx.numel()), dtype=torch.float64, device=idx.device),
                        (idx == 0).sum().to(torch.float64), (sel_cost * picked).sum()])
    return torch.cat([head, hist])


def frenet_sharded(routes, path_id, ego, obs, n_obs, params=None, group=None):
    """Planner pattern: scenarios are block-partitioned over ranks (the arguments are this rank's shard,
    see ``shard_range``; the routes are replicated), each rank projects its egos and plans its scenarios with
    no data-path exchange, and the summary vector of ``frenet_stats`` is all-reduced once.
    Returns (this rank's FrenetResult, all-reduced stats).  ego [n, 5] = x, y, vx, vy, v (CUDA tensors)."""
    from .frenet import frenet_plan, frenet_start
    start = frenet_start(routes, path_id, ego)
    res = frenet_plan(routes, path_id, start, ego[:, :2].contiguous(), obs, n_obs, params)
    stats = frenet_stats(res.index, res.cost)
    allreduce_stats(stats, group)
    return res, stats
