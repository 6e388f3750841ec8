"""The field-deployed DCARL estimator ("RLS"), with the buffer scan on the GPU.

Drop-in for class RLS of Field_testing/.../tools/DCARL/stable_baselines/deepq/RLS.py: same
constructor keywords, same methods (act / act_train / act_test / should_use_rule / add_data /
state_with_action / _calculate_visited_times / _calculate_statistics_index), same files
(visited_state.txt, visited_value.txt, driving_record.txt in the working directory).  The rtree
(RLS.py:61-69) is replaced by a device copy of the buffer (float64 like the reference's boxes; BoxBuffer
also offers a float32 throughput layout) that one kernel pass scans for all candidate actions at once
(csrc/box.cu); `state_index.dat/.idx` are neither read nor written.
"""
import ctypes as C
import os
import os.path as osp
import random
from collections import deque

import numpy as np
import torch

from . import _abi, _dev

# RLS.py:68
VISITED_STATE_DIST = np.array([[1, 0.3, 2, 50, 10, 0.3, 2, 50, 10, 0.3, 2, 50, 10, 0.3, 2, 50,
                                10, 0.3, 2, 50, 0.1]], dtype=np.float64)


class BoxBuffer:
    """Append-only device buffer {state (+) action [M, D+1], return [M]} + the scan.

    dtype=torch.float32 (default) is the throughput layout, 4*(D+2) bytes per row; dtype=torch.float64 keeps
    the reference's own precision (RLS.py:195-196 builds its boxes from float64 rows): a float32 copy moves a
    stored coordinate by up to 4e-6 at |x| ~ 100, which flips queries that sit that close to a box face
    (measured: DESIGN.md section 5).  The drop-in RLS class stores float64."""

    def __init__(self, obs_dimension=20, dist=VISITED_STATE_DIST, capacity=1 << 16, device=None, dtype=torch.float32):
        _dev.require_cuda()
        if dtype not in (torch.float32, torch.float64):
            raise ValueError("BoxBuffer dtype must be torch.float32 or torch.float64")
        self.dtype = dtype
        self._esize = 8 if dtype == torch.float64 else 4
        self.D = int(obs_dimension)
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.dist = _dev.as_device(np.asarray(dist, np.float64).reshape(-1), torch.float64, self.device)
        if self.dist.numel() != self.D + 1:
            raise ValueError("dist must have obs_dimension + 1 entries")
        self.M = 0
        self._state = torch.empty((capacity, self.D + 1), dtype=dtype, device=self.device)
        self._ret = torch.empty((capacity,), dtype=dtype, device=self.device)
        self._ws = None
        self._qws = None            # workspace of the one-launch query (query()): call counter + two accumulators
        self._index_dim = None      # see build_index(): rows [0, _index_rows) are sorted by this dim
        self._index_rows = 0
        self._index_col = None
        self.last_scanned_rows = 0

    def __len__(self):
        return self.M

    @property
    def state(self):
        return self._state[:self.M]

    @property
    def ret(self):
        return self._ret[:self.M]

    def append(self, states, returns):
        """states [m, D+1] (or [D+1]), returns [m]; numpy or tensors."""
        s = _dev.as_device(states, self.dtype, self.device).reshape(-1, self.D + 1)
        r = _dev.as_device(returns, self.dtype, self.device).reshape(-1)
        m = s.shape[0]
        if r.numel() != m:
            raise ValueError("one return per stored state expected")
        self.reserve(self.M + m)
        self._state[self.M:self.M + m] = s
        self._ret[self.M:self.M + m] = r
        self.M += m

    def append_episodes(self, obs, action, rew, done, ep_offsets, gamma=0.95, lag=10):
        """Batched RLS.add_data (RLS.py:185-215) for many concurrent transition streams: builds the
        lagged / terminal-discounted returns on the device and appends the recorded transitions
        (stream-major, time order).  obs [N, D], action/rew/done [N], ep_offsets [E+1] (stream e owns
        transitions ep_offsets[e]:ep_offsets[e+1]).  Returns the number of rows appended."""
        L = _abi.lib()
        o = _dev.as_device(obs, torch.float32, self.device).reshape(-1, self.D).contiguous()
        a = _dev.as_device(action, torch.int32, self.device).reshape(-1)
        r = _dev.as_device(rew, torch.float32, self.device).reshape(-1)
        if isinstance(done, torch.Tensor):
            d = done.to(device=self.device, dtype=torch.uint8).reshape(-1).contiguous()
        else:
            d = torch.from_numpy(np.ascontiguousarray(np.asarray(done).astype(np.uint8)).reshape(-1)).to(self.device)
        off = _dev.as_device(ep_offsets, torch.int64, self.device).reshape(-1)
        N, E = o.shape[0], off.numel() - 1
        if not (a.numel() == r.numel() == d.numel() == N):
            raise ValueError("obs, action, rew and done must describe the same N transitions")
        if E <= 0 or N == 0:
            return 0
        sh = _dev.stream_handle()
        counts = torch.empty(E, dtype=torch.int64, device=self.device)
        _abi.check(L.dcarl_add_data_counts(_dev.ptr(r), _dev.ptr(d), _dev.ptr(off), E, lag, _dev.ptr(counts), sh))
        row_off = torch.zeros(E + 1, dtype=torch.int64, device=self.device)
        torch.cumsum(counts, 0, out=row_off[1:])
        total = int(row_off[-1].item())
        if total == 0:
            return 0
        if int(off[0].item()) != 0 or int(off[-1].item()) != N or bool((off[1:] < off[:-1]).any()):
            raise ValueError("ep_offsets must start at 0, end at the number of transitions and be non-decreasing")
        ep_of_row = torch.repeat_interleave(torch.arange(E, dtype=torch.int32, device=self.device), off[1:] - off[:-1])
        self.reserve(self.M + total)
        f32 = self.dtype == torch.float32       # the writer produces float32 rows (its inputs are float32)
        st_out = self._state[self.M:self.M + total] if f32 else torch.empty((total, self.D + 1), dtype=torch.float32, device=self.device)
        rt_out = self._ret[self.M:self.M + total] if f32 else torch.empty((total,), dtype=torch.float32, device=self.device)
        _abi.check(L.dcarl_add_data_rows(_dev.ptr(o), _dev.ptr(a), _dev.ptr(r), _dev.ptr(d), _dev.ptr(off),
                                         _dev.ptr(row_off), _dev.ptr(ep_of_row), N, self.D, lag, float(gamma),
                                         _dev.ptr(st_out), _dev.ptr(rt_out), sh))
        if not f32:
            self._state[self.M:self.M + total] = st_out
            self._ret[self.M:self.M + total] = rt_out
        self.M += total
        return total

    def reserve(self, rows):
        """Grow the device arrays to hold at least `rows` rows (amortised doubling)."""
        if rows > self._state.shape[0]:
            cap = max(2 * self._state.shape[0], rows)
            ns = torch.empty((cap, self.D + 1), dtype=self.dtype, device=self.device)
            nr = torch.empty((cap,), dtype=self.dtype, device=self.device)
            ns[:self.M] = self._state[:self.M]
            nr[:self.M] = self._ret[:self.M]
            self._state, self._ret = ns, nr

    def _scan(self, row_lo, row_hi, q, n_actions, out, stream=None, ws=None):
        """One kernel pass over rows [row_lo, row_hi) (row_lo a multiple of 4: 16-byte aligned tiles).
        `ws`: a caller-owned threshold workspace (CUDA graphs keep their own: the shared one is replaced
        when a later call needs more room, which would leave a captured pointer dangling)."""
        L = _abi.lib()
        Q = q.shape[0]
        need = L.dcarl_box_moments_ws(Q, self.D, n_actions)
        if ws is None:
            if self._ws is None or self._ws.numel() < need:
                self._ws = torch.empty((max(need, 16),), dtype=torch.uint8, device=self.device)
            ws = self._ws
        elif ws.numel() < need:
            raise ValueError("threshold workspace too small for this query shape")
        sp = C.c_void_p(self._state.data_ptr() + row_lo * (self.D + 1) * self._esize)
        rp = C.c_void_p(self._ret.data_ptr() + row_lo * self._esize)
        with torch.cuda.device(self.device):
            _abi.check(L.dcarl_box_moments(sp, rp, int(self._esize == 8), row_hi - row_lo, self.D, _dev.ptr(self.dist), _dev.ptr(q), Q,
                                           n_actions, _dev.ptr(out), _dev.ptr(ws), ws.numel(),
                                           _dev.stream_handle(stream)))
        self.last_scanned_rows += (row_hi - row_lo) if Q else 0

    def query_fused(self, q, n_actions, rls_params, ws, out, peer_ctx=None, stream=None):
        """One launch for a whole act_test query pass (dcarl_box_query; a second small launch when there are many
        (query, action) pairs): q float64 CUDA [Q, D]; ws a uint8 CUDA tensor of at least dcarl_box_query_ws() + 16
        bytes, zero-initialised ONCE and then owned by the calls that use it (it holds the call counter and the two
        alternating accumulators); out = (moments [Q, A, 3], mean_var [Q, A, 2], action int32 [Q]) CUDA tensors;
        peer_ctx sums the moments over the ranks holding the other row shards."""
        mom, mv, pick = out
        L = _abi.lib()
        with torch.cuda.device(self.device):
            _abi.check(L.dcarl_box_query(
                _dev.ptr(self._state), _dev.ptr(self._ret), int(self._esize == 8), self.M, self.D, _dev.ptr(self.dist), _dev.ptr(q),
                q.shape[0], n_actions, rls_params or _abi.RlsParams(), _dev.ptr(mom), _dev.ptr(mv), _dev.ptr(pick),
                C.c_void_p((ws.data_ptr() + 15) & ~15), ws.numel() - 16, C.byref(peer_ctx) if peer_ctx is not None else None,
                _dev.stream_handle(stream)))
        self.last_scanned_rows = self.M
        return mom, mv, pick

    def moments(self, queries, n_actions, out=None, stream=None, ws=None):
        """queries float64 [Q, D] (numpy or CUDA tensor) -> CUDA float64 [Q, n_actions, 3] = (n, sum R, sum R^2).
        One streaming pass over the whole buffer, whatever Q is."""
        q = _dev.as_device(queries, torch.float64, self.device).reshape(-1, self.D).contiguous()
        if out is None:
            out = torch.empty((q.shape[0], n_actions, 3), dtype=torch.float64, device=self.device)
        self.last_scanned_rows = 0
        self._scan(0, self.M, q, n_actions, out, stream, ws)
        return out

    # ---- on-device index (SURVEY "next" row f1) ------------------------------------------------
    def build_index(self, dim=None):
        """Sort the stored rows by one observation dimension so that a query only has to scan the
        contiguous run of rows whose value in that dimension lies within the box half-width of the
        query (every other row fails the closed-box test in that dimension anyway).  The scan kernel
        and its exact test are unchanged; only its row range shrinks.  `dim` defaults to the
        dimension with the smallest box-width / data-spread ratio.  Rows appended later form an
        unsorted tail that is scanned in full until the index is rebuilt."""
        if self.M == 0:
            return None
        st = self.state
        if dim is None:
            qs = torch.quantile(st[:: max(1, self.M // 200000), :self.D].double(),
                                torch.tensor([0.01, 0.99], dtype=torch.float64, device=self.device), dim=0)
            spread = (qs[1] - qs[0]).clamp_min(1e-12)
            dim = int(torch.argmin(2 * self.dist[:self.D] / spread).item())
        n_sorted = (self.M // 4) * 4          # keeps the tail 16-byte aligned; <= 3 rows stay in the tail
        order = torch.argsort(st[:n_sorted, dim], stable=True)
        self._state[:n_sorted] = st[:n_sorted][order]
        self._ret[:n_sorted] = self.ret[:n_sorted][order]
        self._index_dim, self._index_rows = int(dim), n_sorted
        self._index_col = self._state[:n_sorted, dim].contiguous()
        return self._index_dim

    def moments_indexed(self, queries, n_actions, out=None, stream=None):
        """Same result as ``moments`` (counts identical, sums up to float64 addition order), scanning per query only
        the index range that can match plus the unsorted tail: ONE launch for all queries (one warp per query,
        dcarl_box_moments_ranges), the ranges found on the device.  ``last_scanned_rows`` reports the rows read."""
        if self._index_dim is None:
            return self.moments(queries, n_actions, out)
        q = _dev.as_device(queries, torch.float64, self.device).reshape(-1, self.D).contiguous()
        Q = q.shape[0]
        if out is None:
            out = torch.empty((Q, n_actions, 3), dtype=torch.float64, device=self.device)
        d = self._index_dim
        w = self.dist[d]
        # bracket in the buffer's element type, widened by 2 ulp: a prefilter only, the kernel applies the exact test
        ninf = torch.tensor(-float("inf"), device=self.device, dtype=self.dtype)
        pinf = torch.tensor(float("inf"), device=self.device, dtype=self.dtype)
        lo = torch.nextafter(torch.nextafter((q[:, d] - w).to(self.dtype), ninf), ninf)
        hi = torch.nextafter(torch.nextafter((q[:, d] + w).to(self.dtype), pinf), pinf)
        r0 = torch.searchsorted(self._index_col, lo, right=False).to(torch.int64).contiguous()
        r1 = torch.maximum(torch.searchsorted(self._index_col, hi, right=True).to(torch.int64), r0).contiguous()
        self._scanned = (r1 - r0).sum() + Q * (self.M - self._index_rows)       # (a tensor: read lazily, no sync here)
        if Q:
            with torch.cuda.device(self.device):
                _abi.check(_abi.lib().dcarl_box_moments_ranges(
                    _dev.ptr(self._state), _dev.ptr(self._ret), int(self._esize == 8), self.M, self.D, _dev.ptr(self.dist),
                    _dev.ptr(q), _dev.ptr(r0), _dev.ptr(r1), self._index_rows, self.M, Q, n_actions, _dev.ptr(out),
                    _dev.stream_handle(stream)))
        return out

    @property
    def last_scanned_rows(self):
        """Rows read by the last scan (whole buffer per pass, or the index ranges of moments_indexed)."""
        v = self._scanned
        return int(v.item()) if isinstance(v, torch.Tensor) else int(v)

    @last_scanned_rows.setter
    def last_scanned_rows(self, v):
        self._scanned = v

    def select(self, moments, rls_params=None, stream=None, out=None):
        """moments [Q, n_actions, 3] -> (mean_var CUDA float64 [Q, n_actions, 2], action CUDA int32 [Q])."""
        Q, na, _ = moments.shape
        if out is None:
            mv = torch.empty((Q, na, 2), dtype=torch.float64, device=self.device)
            pick = torch.empty((Q,), dtype=torch.int32, device=self.device)
        else:
            mv, pick = out
        _abi.check(_abi.lib().dcarl_box_select(_dev.ptr(moments), Q, na, rls_params or _abi.RlsParams(),
                                               _dev.ptr(mv), _dev.ptr(pick), _dev.stream_handle(stream)))
        return mv, pick

    def query_graph(self, Q, n_actions=8, rls_params=None):
        """A CUDA-graph replay of the fused query launch (thresholds + scan + totals + select in one kernel)
        for repeated queries of one shape: the per-query cost is one graph launch.  Returns
        ``run(queries) -> (moments, mean_var, action)`` writing into static CUDA tensors; the graph is
        re-captured automatically when rows were appended to the buffer or its storage moved."""
        return _BoxQueryGraph(self, int(Q), int(n_actions), rls_params or _abi.RlsParams())

    def query(self, obs, n_actions=8, rls_params=None):
        """obs [D] or [Q, D] -> host (n [Q, A], mean [Q, A], var [Q, A], action [Q]).  Uses the index
        (if built) for up to 8 queries, the single full pass otherwise."""
        q = np.asarray(obs, np.float64).reshape(-1, self.D)
        use_index = self._index_dim is not None and len(q) <= 8
        if use_index:
            mom = self.moments_indexed(q, n_actions)
            mv, pick = self.select(mom, rls_params)
        else:                                   # thresholds + scan + totals + select in one launch
            qd = _dev.as_device(q, torch.float64, self.device).contiguous()
            Q = qd.shape[0]
            need = int(_abi.lib().dcarl_box_query_ws(Q, self.D, n_actions)) + 16
            if self._qws is None or self._qws.numel() < need:
                self._qws = torch.zeros(need, dtype=torch.uint8, device=self.device)
            mom = torch.empty((Q, n_actions, 3), dtype=torch.float64, device=self.device)
            mv = torch.empty((Q, n_actions, 2), dtype=torch.float64, device=self.device)
            pick = torch.empty((Q,), dtype=torch.int32, device=self.device)
            if Q:
                self.query_fused(qd, n_actions, rls_params, self._qws, (mom, mv, pick))
        mom_h, mv_h = mom.cpu().numpy(), mv.cpu().numpy()
        return mom_h[:, :, 0], mv_h[:, :, 0], mv_h[:, :, 1], pick.cpu().numpy()


class _BoxQueryGraph:
    """A captured replay of one fused query launch (dcarl_box_query) on static tensors the graph owns:
    queries in, (moments, mean_var, action) out, its own zeroed workspace.  Re-captured when the buffer
    grew or moved (append / reserve), never affected by other calls on the same buffer."""

    def __init__(self, buf, Q, n_actions, rls_params):
        self.buf, self.Q, self.na, self.rp = buf, Q, n_actions, rls_params
        dev = buf.device
        self.q = torch.zeros((Q, buf.D), dtype=torch.float64, device=dev)
        self.mom = torch.empty((Q, n_actions, 3), dtype=torch.float64, device=dev)
        self.mv = torch.empty((Q, n_actions, 2), dtype=torch.float64, device=dev)
        self.pick = torch.empty((Q,), dtype=torch.int32, device=dev)
        self.ws = torch.zeros(int(_abi.lib().dcarl_box_query_ws(Q, buf.D, n_actions)) + 16, dtype=torch.uint8, device=dev)
        self._graph, self._key = None, None

    def _current_key(self):
        b = self.buf
        return (len(b), b._state.data_ptr(), b._ret.data_ptr())

    def _record(self):
        buf = self.buf
        out = (self.mom, self.mv, self.pick)
        buf.query_fused(self.q, self.na, self.rp, self.ws, out)        # warm-up outside the capture
        torch.cuda.synchronize(buf.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            buf.query_fused(self.q, self.na, self.rp, self.ws, out)
        self._graph, self._key = g, self._current_key()

    def __call__(self, queries):
        if self._graph is None or self._key != self._current_key():
            self._record()
        self.q.copy_(_dev.as_device(queries, torch.float64, self.buf.device).reshape(self.Q, self.buf.D), non_blocking=True)
        self._graph.replay()
        return self.mom, self.mv, self.pick


class RLS(object):
    """RLS.py:11-241.  `n_actions` (8 in the reference, RLS.py:135) is the only added keyword."""

    def __init__(self,
                 visited_times_thres=30,
                 is_training=True,
                 debug=True,
                 save_new_data=True,
                 create_new_train_file=False,
                 create_new_record_file=True,
                 save_new_driving_data=True,
                 n_actions=8):
        self.visited_times_thres = visited_times_thres
        self.is_training = is_training
        self.trajectory_buffer = deque(maxlen=20)
        self.debug = debug
        self.save_new_data = save_new_data
        self.create_new_train_file = create_new_train_file
        self.create_new_record_file = create_new_record_file
        self.save_new_driving_data = save_new_driving_data
        self.obs_dimension = 20
        self.gamma = 0.95
        self.n_actions = n_actions
        self._cache_obs = None
        self._setup_data_saving()

    # RLS.py:34-76
    def _setup_data_saving(self):
        self.visited_state_dist = VISITED_STATE_DIST.copy()
        # float64 storage: the reference's boxes are built from float64 rows (RLS.py:195-196)
        self.visited_state_tree = BoxBuffer(self.obs_dimension, self.visited_state_dist, dtype=torch.float64)
        if self.create_new_train_file:
            for f in ("visited_state.txt", "visited_value.txt"):
                if osp.exists(f):
                    os.remove(f)
            print("Create new data file...")
            self.visited_state_value = []
            self.visited_state_counter = 0
        else:
            print("loading visited value data...")
            self.visited_state_value = np.loadtxt("visited_value.txt").reshape(-1, 2).tolist()
            self.visited_state_counter = len(self.visited_state_value)
            states = np.loadtxt("visited_state.txt").reshape(-1, self.obs_dimension + 1)
            if len(states) != self.visited_state_counter:
                raise ValueError("visited_state.txt and visited_value.txt disagree on the row count")
            if len(states):
                self.visited_state_tree.append(states, np.asarray(self.visited_state_value)[:, 1])
                self.visited_state_tree.build_index()      # stands in for the rtree: queries scan one sorted run
        self.visited_state_outfile = open("visited_state.txt", "a")
        self.visited_state_format = " ".join(("%f",) * (self.obs_dimension + 1)) + "\n"
        self.visited_value_outfile = open("visited_value.txt", "a")
        self.visited_value_format = " ".join(("%f",) * 2) + "\n"
        if self.create_new_record_file and osp.exists("driving_record.txt"):
            os.remove("driving_record.txt")
        self.driving_record_outfile = open("driving_record.txt", "a")
        self.driving_record_format = " ".join(("%f",) * (self.obs_dimension + 9)) + "\n"
        print("All training data loaded...")

    # one scan answers every candidate action of an observation; cached per observation
    def _stats(self, obs):
        obs = np.asarray(obs, np.float64).reshape(-1)[:self.obs_dimension]
        key = (obs.tobytes(), len(self.visited_state_tree))
        if self._cache_obs != key:
            n, mean, var, pick = self.visited_state_tree.query(
                obs, self.n_actions,
                _abi.RlsParams(visited_times_thres=self.visited_times_thres))
            self._cache = (n[0], mean[0], var[0], int(pick[0]))
            self._cache_obs = key
        return self._cache

    def act(self, obs, RL_action):                       # RLS.py:78-82
        if self.is_training:
            return self.act_train(obs, RL_action)
        return self.act_test(obs, RL_action)

    def act_train(self, obs, RL_action):                 # RLS.py:84-89
        if self.should_use_rule(obs):
            return np.array(0)
        return RL_action

    def state_with_action(self, obs, action):            # RLS.py:91-93
        return np.append(obs, action)

    def should_use_rule(self, obs):                      # RLS.py:95-118
        n, mean, _, _ = self._stats(obs)
        visited_times_rule, mean_rule = int(n[0]), float(mean[0])
        if self.debug:
            print("training: rule_visited_times:", visited_times_rule, "Q_rule:", mean_rule)
        if visited_times_rule < self.visited_times_thres:
            return True
        explore_motivation = random.uniform(-1, 0)
        if explore_motivation < mean_rule:
            return True
        return False

    def act_test(self, obs, RL_action, confidence_thres=0.5):     # RLS.py:120-157
        if confidence_thres == 0.5:
            n, mean, _, pick = self._stats(obs)
        else:
            nq, mq, vq, pq = self.visited_state_tree.query(
                np.asarray(obs, np.float64).reshape(-1)[:self.obs_dimension], self.n_actions,
                _abi.RlsParams(visited_times_thres=self.visited_times_thres, confidence_thres=confidence_thres))
            n, mean, pick = nq[0], mq[0], int(pq[0])
        if self.debug:
            print("testing: rule_visited_times:", int(n[0]), "Q_rule:", float(mean[0]))
            if pick:
                print("RL_visited_times:", int(n[pick]), "Q_RL:", float(mean[pick]))
        return np.array(pick)

    # RLS.py:161-181 — the trailing arguments of the reference signatures are accepted and unused
    def _nmv(self, point):
        """(n, mean, var) for one query point obs (+) action (RLS.py:91-93)."""
        point = np.asarray(point, np.float64).reshape(-1)
        a = float(point[self.obs_dimension])
        ai = int(round(a))
        if ai != a or ai < 0:
            raise ValueError("the action coordinate of a query must be a non-negative integer "
                             "(the reference only ever queries integer candidate actions)")
        if ai < self.n_actions:
            ns, means, vars_, _ = self._stats(point)
        else:                                        # beyond the candidate set: one wider scan, not cached
            ns, means, vars_, _ = self.visited_state_tree.query(point[:self.obs_dimension], ai + 1)
            ns, means, vars_ = ns[0], means[0], vars_[0]
        return int(ns[ai]), float(means[ai]), float(vars_[ai])

    def _calculate_visited_times(self, obs, visited_state_tree=None):
        return self._nmv(obs)[0]

    def _calculate_statistics_index(self, obs, visited_state_value=None, visited_state_tree=None):
        n, mean, var = self._nmv(obs)
        if n == 0:
            return -1, -1, -1                        # RLS.py:169-170
        return mean, var, np.sqrt(var)

    # RLS.py:185-241
    def add_data(self, obs, action, rew, new_obs, done):
        self.trajectory_buffer.append((obs, action, rew, new_obs, done))
        new_states, new_values = [], []
        while len(self.trajectory_buffer) > 10:
            obs_left, action_left, rew_left, _, _ = self.trajectory_buffer.popleft()
            new_states.append(self.state_with_action(obs_left, action_left))
            new_values.append([action_left, rew_left])
        if done:
            _, _, rew_right, _, _ = self.trajectory_buffer[-1]
            while len(self.trajectory_buffer) > 0:
                obs_left, action_left, _, _, _ = self.trajectory_buffer.popleft()
                new_states.append(self.state_with_action(obs_left, action_left))
                new_values.append([action_left, rew_right * self.gamma ** len(self.trajectory_buffer)])
        if self.save_new_data and new_states:
            st = np.asarray(new_states, np.float64)
            va = np.asarray(new_values, np.float64)
            self.visited_state_value.extend(va.tolist())
            self.visited_state_tree.append(st, va[:, 1])
            for s_row, v_row in zip(st, va):
                self.visited_state_outfile.write(self.visited_state_format % tuple(s_row))
                self.visited_value_outfile.write(self.visited_value_format % tuple(v_row))
            self.visited_state_counter += len(st)

        if self.save_new_driving_data:
            n, mean, var, _ = self._stats(obs)
            a = int(action)
            if a == 0:
                rl = (-1, -1, -1)
            else:
                rl = (int(n[a]), float(mean[a]), float(var[a]))
            record = np.concatenate([self.state_with_action(obs, a),
                                     [rew, float(done), int(n[0]), float(mean[0]), float(var[0]), *rl]])
            self.driving_record_outfile.write(self.driving_record_format % tuple(record))
