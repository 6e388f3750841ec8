"""Literal Python/NumPy restatement of the reference loops — TEST INFRASTRUCTURE ONLY.

Each function follows the cited reference lines statement by statement (same
list-append, same ``np.array(list)`` re-materialisation, same ``np.mean`` /
``np.std`` / ``np.argmax`` calls, same operator order) with the script-level globals
turned into arguments and the plotting removed.  It is as slow as the reference on
purpose: it is what ``bench.py --impl reference`` times.

ST1 = Simulation_testing/Simulation_1/test_DCARL.py, ST2 = .../Simulation_2/test_DCARL.py,
DS = Simulation_testing/Simulation_Data_Collection/Data_Sampling/data_sampling.py,
RLS = Field_testing/.../tools/DCARL/stable_baselines/deepq/RLS.py
Pinned by tests/test_oracle.py against tests/golden/*.npz (outputs of the reference).
"""
import math

import numpy as np


# ST1:10-28 (identical in ST2:9-27)
def upper_bound(data_array, alpha=0.05, loc=-50, scale=150):
    return min(100, np.mean(data_array) + scale * math.sqrt(math.log(1 / alpha) / 2 / len(data_array)))


def lower_bound(data_array, alpha=0.05, loc=-50, scale=150):
    return np.mean(data_array) - scale * math.sqrt(math.log(1 / alpha) / 2 / len(data_array))


def CI_lower_bound(data_array, alpha=0.05, loc=-50, scale=150):
    n = len(data_array)
    dsum = np.sum(data_array)
    sigma = np.std(data_array)
    return dsum / n / (n + 1) - 4 * sigma / (n + 1) + dsum / (n + 1) - scale * math.sqrt(math.log(1 / alpha) / 2 / (n + 1))


def mean_value(data_array, alpha=0.05, loc=-50, scale=150):
    return min(100, np.mean(data_array))


def stream(data, state_num, action_num, rule_act=0, n_thres=10, want_overall=True):
    """ST1:41-99 / ST2:41-105.  data: float64 [N, 4] rows [idx, state, act, R].

    Returns row-ordered value / act traces, activation_step[S], overall[N] (ST2 only
    in the reference) and the final TSRL_value table."""
    data_state_act = [[[] for _ in range(action_num)] for _ in range(state_num)]      # ST1:41
    temp = [-50] * action_num                                                         # ST1:51
    temp[rule_act] = 100                                                              # ST1:52
    TSRL_value = np.array([temp] * state_num).tolist()                                # ST1:53
    activation_step = np.array([-1] * state_num)                                      # ST1:58
    activation_value = np.array([-1] * state_num)                                     # ST1:59
    steps_in_state = [0] * state_num
    N = len(data)
    out_value = np.empty(N)
    out_act = np.empty(N, np.int64)
    overall_value = np.empty(N) if want_overall else None
    k = 0
    for idx_ori, state, act_ori, act_value in data:                                  # ST1:73
        idx = int(idx_ori)
        act = int(act_ori)
        data_state_act[idx][act].append(act_value)                                    # ST1:80
        if len(data_state_act[idx][act]) > n_thres:                                   # ST1:86
            if act == rule_act:
                TSRL_value[idx][act] = upper_bound(np.array(data_state_act[idx][act]))
            else:
                TSRL_value[idx][act] = min(lower_bound(np.array(data_state_act[idx][act])),
                                           CI_lower_bound(np.array(data_state_act[idx][act])))
        out_value[k] = max(np.array(TSRL_value[idx]))                                 # ST1:93
        TSRL_act = np.argmax(np.array(TSRL_value[idx]))                               # ST1:94
        out_act[k] = TSRL_act
        steps_in_state[idx] += 1
        if activation_step[idx] == -1 and TSRL_act != rule_act:                       # ST1:98
            activation_step[idx] = steps_in_state[idx]                                # ST1:99
        if want_overall:                                                              # ST2:99-105
            overall_value_step = 0
            for i in range(state_num):
                if activation_step[i] != -1:
                    overall_value_step = overall_value_step + max(np.array(TSRL_value[i])) - activation_value[i] * 0.9
            overall_value[k] = overall_value_step
        k = k + 1
    return dict(value=out_value, act=out_act, activation_step=activation_step.astype(np.int64),
                overall=overall_value, final_value=np.array(TSRL_value, np.float64))


# RLS:68
VISITED_STATE_DIST = np.array([[1, 0.3, 2, 50, 10, 0.3, 2, 50, 10, 0.3, 2, 50, 10, 0.3, 2, 50,
                                10, 0.3, 2, 50, 0.1]])


class BoxBuffer:
    """Brute-force restatement of the rtree store used by RLS (rtree is a third-party
    dependency, unpinned and not installed): insert stores the closed box
    [state - dist, state + dist] (RLS:195-196), intersection(point) yields the ids of
    the boxes that contain the point (libspatialindex closed-interval test)."""

    def __init__(self, visited_state, visited_value, dist=VISITED_STATE_DIST):
        vs = np.asarray(visited_state, np.float64)
        self.lo = vs - dist
        self.hi = vs + dist
        self.visited_state_value = np.asarray(visited_value, np.float64)

    def intersection(self, point):
        p = np.asarray(point, np.float64)
        return np.nonzero(np.all((self.lo <= p) & (p <= self.hi), axis=1))[0]

    # RLS:161-163
    def calculate_visited_times(self, obs):
        return sum(1 for _ in self.intersection(obs.tolist()))

    # RLS:165-181
    def calculate_statistics_index(self, obs):
        if self.calculate_visited_times(obs) == 0:
            return -1, -1, -1
        value_list = [self.visited_state_value[i] for i in self.intersection(obs.tolist())]
        value_array = np.array(value_list)[:, 1]
        mean = np.mean(value_array)
        var = np.var(value_array)
        return mean, var, np.sqrt(var)

    # RLS:120-157
    def act_test(self, obs, visited_times_thres=10, confidence_thres=0.5, n_actions=8):
        from scipy.stats import norm
        rule_state = np.append(obs, 0)
        visited_times_rule = self.calculate_visited_times(rule_state)
        mean_rule, var_rule, _ = self.calculate_statistics_index(rule_state)
        for candidate_action in range(1, n_actions):
            RL_state = np.append(obs, candidate_action)
            visited_times_RL = self.calculate_visited_times(RL_state)
            mean_RL, var_RL, _ = self.calculate_statistics_index(RL_state)
            if visited_times_rule < visited_times_thres or visited_times_RL < 5 or mean_rule > -0.1:
                continue
            var_diff = var_rule / visited_times_rule + var_RL / visited_times_RL
            with np.errstate(all="ignore"):
                z = (mean_RL - mean_rule) / np.sqrt(var_diff)
            if norm.cdf(z) > confidence_thres:
                return np.array(candidate_action)
        return np.array(0)


def add_data_returns(rewards, dones, gamma=0.95, lag=10):
    """RLS:185-215 return construction for ONE env stream, without the I/O.

    For the transition sequence (rew_t, done_t) returns the list of
    (t_recorded, R) in the order the reference appends them to visited_value."""
    from collections import deque
    buf = deque(maxlen=20)
    out = []
    for t, (rew, done) in enumerate(zip(rewards, dones)):
        buf.append((t, rew))
        while len(buf) > lag:                                                         # RLS:188
            t_left, rew_left = buf.popleft()
            out.append((t_left, rew_left))                                            # RLS:192
        if done:                                                                      # RLS:202
            _, rew_right = buf[-1]
            while len(buf) > 0:
                t_left, _ = buf.popleft()
                out.append((t_left, rew_right * gamma ** len(buf)))                   # RLS:207
    return out


def data_generation(u_states, u_action_values, z_state, acts, z_ret,
                    state_num=20, action_num=11, min_value=-50, max_value=100):
    """DS:30-67 with the random deviates supplied (scipy's rvs = loc + scale*deviate).

    acts / z_ret are consumed only for accepted draws (DS:50-55), in order."""
    states = 0 + 1 * np.asarray(u_states)                                             # DS:39
    action_values = [min_value + (max_value - min_value) * np.asarray(u) for u in u_action_values]   # DS:43
    states_idxes = np.floor((3 + 1 * np.asarray(z_state)) / 6 * state_num).astype(int)    # DS:14-15
    data = []
    j = 0
    for idx in states_idxes:                                                          # DS:49
        if idx < 0 or idx >= state_num:
            continue
        act = int(acts[j])
        data.append([int(idx), states[idx], int(act), float(action_values[idx][act] + 50 * z_ret[j])])  # DS:55,9
        j += 1
    return np.array(data), np.array(action_values), np.array(states)
