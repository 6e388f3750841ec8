"""The field-deployed DCARL estimator ("RLS"), with the buffer scan on the GPU.

Drop-in for class RLS of Field_testing/.../tools/DCARL/stable_baselines/deepq/RLS.py: same
constructor keywords, same methods (act / act_train / act_test / should_use_rule / add_data /
state_with_action / _calculate_visited_times / _calculate_statistics_index), same files
(visited_state.txt, visited_value.txt, driving_record.txt in the working directory).  The rtree
(RLS.py:61-69) is replaced by a float32 device copy of the buffer that one kernel pass scans for
all candidate actions at once (csrc/box.cu); `state_index.dat/.idx` are neither read nor written.
"""
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
    """Append-only device buffer {state (+) action float32 [M, D+1], return float32 [M]} + the scan."""

    def __init__(self, obs_dimension=20, dist=VISITED_STATE_DIST, capacity=1 << 16, device=None):
        _dev.require_cuda()
        self.D = int(obs_dimension)
        self.device = torch.device(device or f"cuda:{torch.cuda.current_device()}")
        self.dist = _dev.as_device(np.asarray(dist, np.float64).reshape(-1), torch.float64, self.device)
        if self.dist.numel() != self.D + 1:
            raise ValueError("dist must have obs_dimension + 1 entries")
        self.M = 0
        self._state = torch.empty((capacity, self.D + 1), dtype=torch.float32, device=self.device)
        self._ret = torch.empty((capacity,), dtype=torch.float32, device=self.device)
        self._ws = None

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
        s = _dev.as_device(states, torch.float32, self.device).reshape(-1, self.D + 1)
        r = _dev.as_device(returns, torch.float32, self.device).reshape(-1)
        m = s.shape[0]
        if r.numel() != m:
            raise ValueError("one return per stored state expected")
        if self.M + m > self._state.shape[0]:
            cap = max(2 * self._state.shape[0], self.M + m)
            ns = torch.empty((cap, self.D + 1), dtype=torch.float32, device=self.device)
            nr = torch.empty((cap,), dtype=torch.float32, device=self.device)
            ns[:self.M] = self._state[:self.M]
            nr[:self.M] = self._ret[:self.M]
            self._state, self._ret = ns, nr
        self._state[self.M:self.M + m] = s
        self._ret[self.M:self.M + m] = r
        self.M += m

    def moments(self, queries, n_actions, out=None, stream=None):
        """queries float64 [Q, D] (numpy or CUDA tensor) -> CUDA float64 [Q, n_actions, 3] = (n, sum R, sum R^2)."""
        q = _dev.as_device(queries, torch.float64, self.device).reshape(-1, self.D).contiguous()
        Q = q.shape[0]
        L = _abi.lib()
        need = L.dcarl_box_moments_ws(Q, self.D, n_actions)
        if self._ws is None or self._ws.numel() < need:
            self._ws = torch.empty((max(need, 16),), dtype=torch.uint8, device=self.device)
        if out is None:
            out = torch.empty((Q, n_actions, 3), dtype=torch.float64, device=self.device)
        _abi.check(L.dcarl_box_moments(_dev.ptr(self._state), _dev.ptr(self._ret), self.M, self.D,
                                       _dev.ptr(self.dist), _dev.ptr(q), Q, n_actions, _dev.ptr(out),
                                       _dev.ptr(self._ws), self._ws.numel(), _dev.stream_handle(stream)))
        return out

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
        """A CUDA-graph replay of (zero moments, threshold prepare, scan, select) for repeated queries
        of one shape: the per-query cost is one graph launch instead of four kernel launches plus
        their host overhead (the scan of a 1e7-row shard is ~0.15 ms, shorter than that overhead).
        Returns ``run(queries) -> (moments, mean_var, action)`` writing into static CUDA tensors; the
        graph is re-captured automatically when rows were appended to the buffer."""
        return _BoxQueryGraph(self, int(Q), int(n_actions), rls_params or _abi.RlsParams())

    def query(self, obs, n_actions=8, rls_params=None):
        """obs [D] or [Q, D] -> host (n [Q, A], mean [Q, A], var [Q, A], action [Q])."""
        mom = self.moments(np.asarray(obs, np.float64).reshape(-1, self.D), n_actions)
        mv, pick = self.select(mom, rls_params)
        mom_h, mv_h = mom.cpu().numpy(), mv.cpu().numpy()
        return mom_h[:, :, 0], mv_h[:, :, 0], mv_h[:, :, 1], pick.cpu().numpy()


class _BoxQueryGraph:
    def __init__(self, buf, Q, n_actions, rls_params):
        self.buf, self.Q, self.na, self.rp = buf, Q, n_actions, rls_params
        dev = buf.device
        self.q = torch.zeros((Q, buf.D), dtype=torch.float64, device=dev)
        self.mom = torch.empty((Q, n_actions, 3), dtype=torch.float64, device=dev)
        self.mv = torch.empty((Q, n_actions, 2), dtype=torch.float64, device=dev)
        self.pick = torch.empty((Q,), dtype=torch.int32, device=dev)
        self._graph, self._rows = None, -1

    def _record(self):
        buf = self.buf
        buf.moments(self.q, self.na, out=self.mom)            # sizes the workspace outside the capture
        buf.select(self.mom, self.rp, out=(self.mv, self.pick))
        torch.cuda.synchronize(buf.device)
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            buf.moments(self.q, self.na, out=self.mom)
            buf.select(self.mom, self.rp, out=(self.mv, self.pick))
        self._graph, self._rows = g, len(buf)

    def __call__(self, queries):
        if self._graph is None or self._rows != len(self.buf):
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
        self.visited_state_tree = BoxBuffer(self.obs_dimension, self.visited_state_dist)
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
    def _calculate_visited_times(self, obs, visited_state_tree=None):
        obs = np.asarray(obs, np.float64).reshape(-1)
        a = obs[self.obs_dimension]
        n, _, _, _ = self._stats(obs)
        ai = int(round(float(a)))
        if ai != a or not (0 <= ai < self.n_actions):
            return self._scan_single(obs)[0]
        return int(n[ai])

    def _calculate_statistics_index(self, obs, visited_state_value=None, visited_state_tree=None):
        obs = np.asarray(obs, np.float64).reshape(-1)
        a = obs[self.obs_dimension]
        ai = int(round(float(a)))
        if ai != a or not (0 <= ai < self.n_actions):
            n, mean, var = self._scan_single(obs)
        else:
            ns, means, vars_, _ = self._stats(obs)
            n, mean, var = int(ns[ai]), float(means[ai]), float(vars_[ai])
        if n == 0:
            return -1, -1, -1
        return mean, var, np.sqrt(var)

    def _scan_single(self, point):
        """A query whose action coordinate is not one of the integer candidates (never happens in
        the reference's own call sites): scan with that action as candidate 0 by shifting the
        action axis."""
        raise ValueError("query action must be an integer in [0, n_actions)")

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
