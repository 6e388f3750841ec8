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

    # RLS:95-118 (the debug prints removed); draws from the global `random` like the reference
    def should_use_rule(self, obs, visited_times_thres=30):
        import random
        rule_state = np.append(obs, 0)
        visited_times_rule = self.calculate_visited_times(rule_state)
        mean_rule, var_rule, sigma_rule = self.calculate_statistics_index(rule_state)
        if visited_times_rule < visited_times_thres:
            return True
        explore_motivation = random.uniform(-1, 0)
        if explore_motivation < mean_rule:
            return True
        return False

    # RLS:84-89
    def act_train(self, obs, RL_action, visited_times_thres=30):
        if self.should_use_rule(obs, visited_times_thres):
            return np.array(0)
        return RL_action

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


# ---- candidate-trajectory rollout (closed loop of the collector) ---------------------------------------
# DVC = Simulation_testing/Simulation_Data_Collection/Data_From_Carla/Agent/drl_library/dqn/dqn_value_collect.py
# TS  = .../Data_From_Carla/Test_Scenarios/TestScenario_Town03.py
# PRD = .../Data_From_Carla/Agent/zzz/predict.py,  JTP = .../Agent/zzz/JunctionTrajectoryPlanner.py
# Written from those lines (not from the CUDA kernel).  CARLA itself cannot run here: the ego is placed AT
# point k of the locked-in trajectory at time k*DT and its speed is the distance between consecutive
# points / DT; every rule below is the reference's.
def predict_circles(ego_xy, vehicles, considered, move_gap):
    """PRD:64-110.  vehicles: rows [x, y, vx, vy, yaw].  Returns [(x0, y0, vx, vy)] for the front and the
    back circle of each of the `considered` vehicles nearest to the ego (sorted() is stable: PRD:78)."""
    d = [(float(np.linalg.norm(np.array([v[0], v[1]]) - np.asarray(ego_xy))), i) for i, v in enumerate(vehicles)]   # PRD:72-76
    order = sorted(d, key=lambda t: t[0])                                                                            # PRD:78
    out = []
    for _, i in order[:considered]:                                                                                  # PRD:80-83
        x, y, vx, vy, yaw = vehicles[i]
        out.append((x + math.cos(yaw) * move_gap, y + math.sin(yaw) * move_gap, vx, vy))                            # PRD:103-104
        out.append((x - math.cos(yaw) * move_gap, y - math.sin(yaw) * move_gap, vx, vy))                            # PRD:105-106
    return out


def rollout_episode(xs, ys, brake, circles, n_pred, dt=0.3, radius=1.0, ticks_per_point=6.0,
                    collision_reward=-100.0, pass_y=73.7, stuck_speed=0.1, stay_thres=2.0):
    """One episode along the trajectory (xs, ys) [DVC:106-113 keeps following the trajectory locked in at the
    trigger].  `brake`: the backup trajectory, i.e. the same points with every desired speed zeroed
    (JTP:147-150) — the ego stands still at the first point.  Returns (return, points consumed, term) with
    term 0 = trajectory ran out, 1 = collision, 2 = pass, 3 = stuck."""
    n = len(xs)
    episode_reward = 0.0
    stuck_time = 0.0                       # TS:183-185: the first call only arms the clock
    for k in range(n):
        if brake:
            x, y, v = xs[0], ys[0], 0.0
        else:
            x, y = xs[k], ys[k]
            k0 = k if k + 1 < n else k - 1
            v = math.hypot(xs[k0 + 1] - xs[k0], ys[k0 + 1] - ys[k0]) / dt if n >= 2 else 0.0
        reward = ticks_per_point * (math.sqrt(v) * 0.1)                          # TS:403-404, per simulator tick
        done, term = False, 0
        # PRD:49-59: for t in range(2, min(len(fp.x) - 1, len(path.t) - 1), 2), any circle within the radius
        if k in range(2, min(n - 1, n_pred - 1), 2):
            for (cx, cy, vx, vy) in circles:
                px, py = cx + k * dt * vx, cy + k * dt * vy                      # PRD:99-100
                if (px - x) ** 2 + (py - y) ** 2 <= radius ** 2:                 # PRD:57-58
                    reward, done, term = collision_reward, True, 1               # TS:407-410
                    break
        if y < pass_y:                                                           # TS:414-415, :198
            done = True
            term = term or 2
        else:                                                                    # TS:419 elif
            now = k * dt
            if k > 0 and v < stuck_speed:                                        # TS:188
                if now - stuck_time > stay_thres:                                # TS:189-190
                    reward, done, term = 0.0, True, 3                            # TS:420-422
            else:
                stuck_time = now                                                 # TS:192 (and :184 on the first call)
        episode_reward += reward                                                 # DVC:119
        if done:
            return episode_reward, k + 1, term
    return episode_reward, n, 0


def rollout_candidates(paths, n_valid, cost, ego_xy, vehicles, considered=4, move_gap=1.0, **kw):
    """All C+1 actions of one scenario (DVC:144-145 cycles used_action over 0..len(candidates)).
    paths [C, Tm, 2], n_valid [C], cost [C].  Action 0 = brake over the cheapest candidate's points
    (JTP:151-153: sorted by cost, stable), action c+1 = candidate c (JTP:139)."""
    circles = predict_circles(ego_xy, vehicles, considered, move_gap)
    C = len(n_valid)
    cheapest = sorted(range(C), key=lambda c: cost[c])[0]
    out = []
    for act in range(C + 1):
        c = cheapest if act == 0 else act - 1
        n = int(n_valid[c])
        out.append(rollout_episode(list(paths[c, :n, 0]), list(paths[c, :n, 1]), act == 0, circles, **kw))
    return out
