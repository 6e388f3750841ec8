"""Multi-GPU plumbing: scenarios shard across ranks with no data-path exchange; one small
all-reduce of the float64 summary-statistics vector per batch.

One process per GPU (torchrun).  NCCL on GPUs, gloo for the CPU tests of this host logic.
"""
import ctypes as C
import os

import torch
import torch.distributed as dist

from . import _abi, _dev

STATS_HEADER = 5   # DCARL_STATS_HEADER


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


def allgather(t, group=None):
    """List of every rank's copy of `t` (rank order); [t] without a process group."""
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        out = [torch.empty_like(t) for _ in range(dist.get_world_size(group))]
        dist.all_gather(out, t, group=group)
        return out
    return [t]


def _world(group=None):
    return dist.get_world_size(group) if (dist.is_available() and dist.is_initialized()) else 1


class PeerGroup:
    """This rank's peer-mapped all-reduce workspace and the mapped workspaces of the other ranks of the
    NVLink box (include/dcarl_b200.h, dcarl_peer_ctx).  The mapping is plumbing: torch's symmetric-memory
    allocator does the CUDA IPC exchange; the sum itself runs inside this package's kernels (peer stores
    over NVLink, a flag per rank, rank-ordered adds).  ``ctx`` is None for a single rank, and when the
    ranks cannot map each other's memory (``backend`` then says "nccl": callers fall back to
    torch.distributed.all_reduce)."""

    def __init__(self, n_max, group=None, device=None):
        self.group, self.n_max = group, int(n_max)
        self.world = _world(group)
        self.rank = dist.get_rank(group) if self.world > 1 else 0
        self.ctx, self.backend, self.error = None, "single", None
        if self.world == 1:
            return
        self.backend = "nccl"
        if self.world > _abi.MAX_PEERS or not torch.cuda.is_available():
            return
        try:
            import torch.distributed._symmetric_memory as symm
            dev = torch.device(device or f"cuda:{torch.cuda.current_device()}")
            L = _abi.lib()
            nbytes = int(L.dcarl_peer_ws_bytes(self.world, self.n_max))
            pg = group if group is not None else dist.group.WORLD
            if hasattr(symm, "enable_symm_mem_for_group"):
                try:
                    symm.enable_symm_mem_for_group(pg.group_name)
                except Exception:   # noqa: BLE001  (deprecated no-op on newer torch)
                    pass
            self.buf = symm.empty(nbytes, dtype=torch.uint8, device=dev)
            self.hdl = symm.rendezvous(self.buf, pg)
            self.buf.zero_()
            torch.cuda.synchronize(dev)
            dist.barrier(group)
            so, fo, eo = C.c_size_t(), C.c_size_t(), C.c_size_t()
            L.dcarl_peer_ws_layout(self.world, self.n_max, C.byref(so), C.byref(fo), C.byref(eo))
            ptrs = [int(p) for p in self.hdl.buffer_ptrs]
            ctx = _abi.PeerCtx()
            for r in range(self.world):
                ctx.slots[r] = ptrs[r] + so.value
                ctx.flags[r] = ptrs[r] + fo.value
            ctx.epoch = self.buf.data_ptr() + eo.value
            ctx.rank, ctx.world, ctx.n_max = self.rank, self.world, self.n_max
            self.ctx, self.backend = ctx, "peer-memory"
        except Exception as e:   # noqa: BLE001  (no P2P mapping on this box: NCCL carries the sum)
            self.error = repr(e)
        # every rank must take the same path
        ok = torch.tensor([1 if self.ctx is not None else 0], device="cuda")
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)
        if int(ok.item()) == 0:
            self.ctx, self.backend = None, "nccl"

    def allreduce(self, t, stream=None):
        """In-place SUM of a float64 CUDA tensor (<= n_max elements) over the ranks."""
        if self.world == 1:
            return t
        if self.ctx is None:
            dist.all_reduce(t, op=dist.ReduceOp.SUM, group=self.group)
            return t
        if t.dtype != torch.float64 or not t.is_contiguous() or t.numel() > self.n_max:
            raise ValueError("PeerGroup.allreduce: contiguous float64 tensor of at most n_max elements expected")
        _abi.check(_abi.lib().dcarl_peer_allreduce(_dev.ptr(t), t.numel(), C.byref(self.ctx), _dev.stream_handle(stream)))
        return t


def nccl_comm_ptr(group=None):
    """The ncclComm_t of torch's NCCL process group as an integer (for dcarl_allreduce_stats), or None."""
    try:
        pg = group if group is not None else dist.group.WORLD
        be = pg._get_backend(torch.device("cuda"))
        return int(be._comm_ptr())
    except Exception:   # noqa: BLE001
        return None


def allreduce_stats_abi(stats, group=None, stream=None):
    """The C-ABI form of the stats all-reduce (dcarl_allreduce_stats: ncclAllReduce on torch's communicator,
    issued on `stream`).  Falls back to torch.distributed.all_reduce when the communicator is not exposed."""
    if _world(group) == 1:
        return stats
    comm = nccl_comm_ptr(group)
    if comm is None:
        return allreduce_stats(stats, group)
    _abi.check(_abi.lib().dcarl_allreduce_stats(_dev.ptr(stats), stats.numel(), C.c_void_p(comm), _dev.stream_handle(stream)))
    return stats


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


class ShardedBoxQuery:
    """Config-4 pattern for repeated queries of one shape: ``run(queries) -> (moments, mean_var, action)``.
    The buffer is row-sharded over the ranks of `group`, the queries are replicated.  One call is ONE
    kernel launch (dcarl_box_query): thresholds, the scan of this rank's rows, the sum of the raw moments
    over ranks through peer memory, and the act_test pick, the last three in the kernel's epilogue.
    Static output tensors (overwritten by the next call).  Without peer mapping between the ranks the sum
    goes through NCCL between a scan launch and a select launch."""

    def __init__(self, buf, Q, n_actions, rls_params=None, group=None, peers=None, graph=False):
        """graph=True replays the launch from a CUDA graph (captured on the first call, re-captured when the
        buffer grew or moved): the per-query host cost drops to one graph launch.  The peer all-reduce keeps
        its call counter on the device, so it is captured with the kernel; every rank must use the same setting."""
        self.buf, self.Q, self.na, self.group = buf, int(Q), int(n_actions), group
        self.use_graph, self._graph, self._key = bool(graph), None, None
        self.rp = rls_params or _abi.RlsParams()
        dev = buf.device
        self.q = torch.zeros((self.Q, buf.D), dtype=torch.float64, device=dev)
        self.mom = torch.empty((self.Q, self.na, 3), dtype=torch.float64, device=dev)
        self.mv = torch.empty((self.Q, self.na, 2), dtype=torch.float64, device=dev)
        self.pick = torch.empty((self.Q,), dtype=torch.int32, device=dev)
        self.world = _world(group)
        L = _abi.lib()
        self.ws = torch.zeros(int(L.dcarl_box_query_ws(self.Q, buf.D, self.na)) + 16, dtype=torch.uint8, device=dev)
        self.peers = peers if peers is not None else PeerGroup(min(self.Q, 64) * self.na * 3, group, dev)
        many = min(self.Q, 64) * self.na * 3 > 4096          # (kSoloEpilogueMax in csrc/box.cu: too much for the fused epilogue's one CTA)
        if self.world == 1 and many:
            self.path = "two launches: thresholds + scan, then one CTA per query for the select"
        elif self.world == 1:
            self.path = "one launch: thresholds + scan + totals + select"
        elif self.peers.ctx is not None and many:
            self.path = ("two launches: thresholds + scan, then one CTA per query for the peer-memory sum over ranks "
                         "and the select")
        elif self.peers.ctx is not None:
            self.path = "one launch: thresholds + scan + peer-memory sum over ranks + select"
        else:
            self.path = "scan + NCCL all-reduce of the moments + select"

    def __call__(self, queries):
        buf = self.buf
        q = queries if (isinstance(queries, torch.Tensor) and queries.is_cuda and queries.dtype == torch.float64
                        and queries.is_contiguous()) else _dev.as_device(queries, torch.float64, buf.device)
        q = q.reshape(self.Q, buf.D)
        if self.world > 1 and self.peers.ctx is None:
            buf.moments(q, self.na, out=self.mom)
            allreduce_stats(self.mom, self.group)
            buf.select(self.mom, self.rp, out=(self.mv, self.pick))
            return self.mom, self.mv, self.pick
        ctx = self.peers.ctx if self.world > 1 else None
        out = (self.mom, self.mv, self.pick)
        if not self.use_graph:
            buf.query_fused(q, self.na, self.rp, self.ws, out, ctx)
            return out
        key = (len(buf), buf._state.data_ptr(), buf._ret.data_ptr())
        self.q.copy_(q, non_blocking=True)
        if self._graph is None or self._key != key:
            buf.query_fused(self.q, self.na, self.rp, self.ws, out, ctx)      # warm-up outside the capture
            torch.cuda.synchronize(buf.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                buf.query_fused(self.q, self.na, self.rp, self.ws, out, ctx)
            self._graph, self._key = g, key
        self._graph.replay()
        return out


class MixedStreamBatch:
    """Config-5 pattern: this rank's shard of a batch that mixes scenario shapes (one StreamPlan per shape,
    e.g. Simulation_1-shaped S=1/A=30 and Simulation_2-shaped S=20/A=11).  ``run(data)`` launches every
    plan on its rows; the summary kernels chain into ONE vector (the action histogram zero-padded to the
    widest shape) and the last of them ends with the sum over ranks through peer memory — one collective
    per batch, fused into the kernel that produces the vector.  ``local_stats`` is this rank's vector
    before the sum (kept only when ``keep_local``)."""

    def __init__(self, plans, group=None, peers=None, keep_local=True):
        self.plans, self.group, self.keep_local = list(plans), group, keep_local
        self.A_max = max(p.A for p in self.plans)
        self.width = STATS_HEADER + self.A_max
        dev = self.plans[0].device
        self.local_stats = torch.zeros(self.width, dtype=torch.float64, device=dev)
        self.stats = torch.zeros_like(self.local_stats)
        self.world = _world(group)
        self.peers = peers if peers is not None else PeerGroup(self.width, group, dev)
        for p in self.plans:
            p.want_stats_kernel = False           # the batch runs the summary kernels itself

    def run(self, data, stream=None):
        """data: one (idx, act, R) triple of CUDA tensors per plan.  Returns the all-reduced stats tensor."""
        L = _abi.lib()
        fused = self.world > 1 and self.peers.ctx is not None and not self.keep_local
        target = self.stats if fused else self.local_stats
        last = len(self.plans) - 1
        for k, (plan, (i, a, r)) in enumerate(zip(self.plans, data)):
            plan.run_device(i, a, r, stream=stream)
            with torch.cuda.device(plan.device):
                _abi.check(L.dcarl_stream_stats_reduce(
                    _dev.ptr(plan.value), int(plan.out_dtype == torch.float64), _dev.ptr(plan.overall),
                    _dev.ptr(plan.activation_step), _dev.ptr(plan.act_hist), plan.n_inst, plan.n_steps, plan.S, plan.A,
                    _dev.ptr(target), self.width, int(k > 0),
                    C.byref(self.peers.ctx) if (fused and k == last) else None, _dev.stream_handle(stream)))
        if not fused:
            self.stats.copy_(self.local_stats)
            self.peers.allreduce(self.stats, stream)
        return self.stats


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
    # (torch.where, not a product: the selected-cost slot of a scenario without a pick can hold NaN)
    head = torch.stack([torch.tensor(float(idx.numel()), dtype=torch.float64, device=idx.device),
                        (idx == 0).sum().to(torch.float64), torch.where(picked, sel_cost, torch.zeros_like(sel_cost)).sum()])
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
