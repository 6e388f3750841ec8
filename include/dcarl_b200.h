/*
 * dcarl_b200.h — C ABI of libdcarl_b200.so: the B200 (sm_100a) implementation of DCARL's
 * data-parallel hot path.  The reference (zhcao92/DCARL) has no FFI of its own — the path
 * is plain Python — so each entry point below names the reference
 * function(s) it replaces; INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add.
 *
 * Reference files (relative to the reference repository root):
 *   ST1 = Simulation_testing/Simulation_1/test_DCARL.py
 *   ST2 = Simulation_testing/Simulation_2/test_DCARL.py
 *   DS  = Simulation_testing/Simulation_Data_Collection/Data_Sampling/data_sampling.py
 *   RLS = Field_testing/Software_and_Raw_Data_on_Self-Driving_Vehicle/software/src/tools/DCARL/
 *         stable_baselines/deepq/RLS.py
 *   JTP / SPL / PRD = Simulation_testing/Simulation_Data_Collection/Data_From_Carla/Agent/zzz/
 *         JunctionTrajectoryPlanner.py / cubic_spline_planner.py / predict.py
 *
 * Conventions
 *   - extern "C", plain pointers and sizes.  Unless a parameter says "host", every pointer is
 *     a caller-owned DEVICE pointer on the current CUDA device.  No hidden allocation: calls
 *     that need scratch take (ws, ws_bytes) sized by the matching *_ws() query.
 *   - every compute call takes a cudaStream_t (passed as void*) and is asynchronous on it.
 *   - return value: 0 = ok; < 0 = invalid argument (-(1-based argument index));
 *     > 0 = cudaError_t of the failing runtime call / launch.
 *     dcarl_last_error() returns a thread-local, human-readable description.
 *   - re-entrant; no global mutable state besides the per-thread error string.
 *   - there is NO CPU fallback: without a CUDA device every compute call returns an error.
 */
#ifndef DCARL_B200_H
#define DCARL_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DCARL_B200_ABI_VERSION 2

typedef void *dcarl_stream_t; /* cudaStream_t */

/* Constants of the bound estimator.  Defaults are the literals of ST1:10-12, 43-52. */
typedef struct dcarl_params {
    double alpha;            /* ST1:10  0.05 */
    double scale;            /* ST1:10  150  */
    double cap;              /* ST1:12  min(100, .) on the rule action */
    double prior_rule;       /* ST1:52  100  */
    double prior_other;      /* ST1:51  -50  */
    double activation_bonus; /* ST2:103 "- activation_value[i]*0.9", activation_value == -1 -> +0.9 */
    int32_t n_thres;         /* ST1:45  10, the cell refreshes when len > n_thres (strict) */
    int32_t rule_act;        /* ST1:43  0   */
} dcarl_params;

/* Constants of RLS.act_test (RLS:120-157) as constructed by dqn.py:106-112. */
typedef struct dcarl_rls_params {
    double rule_mean_gate;        /* RLS:141  -0.1 : skip when mean_rule > gate */
    double confidence_thres;      /* RLS:120  0.5  */
    int32_t visited_times_thres;  /* RLS:141  rule action needs >= this many visits (10 as deployed) */
    int32_t min_candidate_visits; /* RLS:141  5 */
} dcarl_rls_params;

/* ---- multi-GPU: the summary vectors are summed over the ranks of one NVLink / NVSwitch box -----------
 * The path shards with no data-path exchange: scenario instances (stream) or buffer rows
 * (box) are partitioned over the GPUs and one small vector is summed per batch — <= 48 KB, so latency
 * is what counts.  The kernels that produce such a vector can finish with the sum themselves, over
 * peer-mapped memory: rank r owns a workspace that every other rank has mapped into its address space
 * (CUDA IPC or symmetric memory; mapping is the host's business), laid out as
 *   slots [2][world][n_max] float64 | flags [2][world] uint32 | epoch uint32     (dcarl_peer_ws_bytes)
 * zero-initialised once, before the first call on any rank.  A call writes this rank's vector into its
 * slot on EVERY rank, raises its flag there, waits for all flags of its own workspace and adds the slots
 * in rank order (bit-identical sums on every rank, unlike an order-dependent reduction tree).  Two
 * parities alternate per call; the call counter lives on the device, so launches can be graph-captured.
 * All ranks must issue the same sequence of calls with the same lengths.  world <= DCARL_MAX_PEERS. */
#define DCARL_MAX_PEERS 16
typedef struct dcarl_peer_ctx {
    void *slots[DCARL_MAX_PEERS];  /* slots[r]: base of rank r's slot area, as mapped in THIS process */
    void *flags[DCARL_MAX_PEERS];  /* flags[r]: base of rank r's flag area, as mapped in THIS process */
    void *epoch;                   /* this rank's call counter (device memory) */
    int32_t rank, world, n_max, pad_;
} dcarl_peer_ctx;
size_t dcarl_peer_ws_bytes(int32_t world, int32_t n_max);   /* size of one rank's workspace */
/* Offsets of the three areas inside a workspace of dcarl_peer_ws_bytes(world, n_max) bytes (host pointers). */
void dcarl_peer_ws_layout(int32_t world, int32_t n_max, size_t *slots_off, size_t *flags_off, size_t *epoch_off);
/* data[0..n) (device, float64) <- sum over ranks, in place; one CTA, asynchronous on `stream`. */
int dcarl_peer_allreduce(double *data, int32_t n, const dcarl_peer_ctx *peers /* host */, dcarl_stream_t stream);
/* The same sum through NCCL: ncclAllReduce(data, data, n, ncclDouble, ncclSum, comm, stream) on the
 * library's copy of libnccl (resolved at run time; an error if it cannot be loaded).  `comm` is an
 * ncclComm_t passed as void*.  (dcarl_allreduce_stats.) */
int dcarl_allreduce_stats(double *stats, int32_t n, void *comm, dcarl_stream_t stream);

const char *dcarl_last_error(void);
int dcarl_abi_version(void);
/* Fills the current device's SM count and compute capability (host pointers). */
int dcarl_device_info(int *sm_count, int *cc_major, int *cc_minor);
void dcarl_default_params(dcarl_params *p);          /* host pointer */
void dcarl_default_rls_params(dcarl_rls_params *p);  /* host pointer */

/* ---- a1-a4: upper_bound / lower_bound / CI_lower_bound / mean_value, ST1:10-28 -----------
 * Segmented: sample array `seg` occupies x[seg_offsets[seg] .. seg_offsets[seg+1]).
 * out[seg*4 + {0,1,2,3}] = {upper_bound, lower_bound, CI_lower_bound, mean_value}; an empty
 * segment yields NaN (numpy: mean of an empty array).  Two-pass mean / population std in
 * float64 like np.mean / np.std.  x is float64 (x_is_f64 != 0) or float32. */
int dcarl_bounds(const void *x, int x_is_f64, const int64_t *seg_offsets, int64_t n_seg,
                 const dcarl_params *p /* host */, double *out, dcarl_stream_t stream);

/* ---- a5-a7: the streaming estimator + argmax selector, ST1:73-99 / ST2:72-105 ------------
 * n_inst independent scenario instances, each a stream of n_steps rows {idx, act, R}.
 * Layout: row-major [n_inst, n_steps] for idx (int32, 0 <= idx < S), act (int32, 0 <= act < A)
 * and R (float32, or float64 when r_is_f64).  Per row k of instance i the kernel appends R to
 * cell (idx, act), refreshes that cell's bound once it holds more than n_thres samples
 * (rule action: upper bound; others: min(lower, CI lower)), and emits
 *   out_value[i,k] = max_a value[idx][a]          (float32, or float64 when out_is_f64)
 *   out_act[i,k]   = argmax_a value[idx][a]       (first maximum, like np.argmax)
 *   out_overall[i,k] (nullable) = sum over activated states of (max + activation_bonus), ST2:99-105
 *   out_activation_step[i,s] = per-state 1-based step at which argmax first left rule_act, else -1
 *   act_hist[A] (nullable, uint64, ACCUMULATED into) = histogram of out_act over all rows.
 * Rows with idx/act out of range are an input error (the reference raises IndexError): they never
 * index out of bounds, leave the outputs of their instance unspecified, and are counted, one per
 * row, into *bad_rows (nullable, uint64, accumulated); callers must treat *bad_rows != 0 as a
 * failed call.
 * variant: 0 = default (row-parallel shared-memory kernel when the S*A table fits in shared
 * memory, else the global-memory walker), 1 = force the walker, 2 = force the row-parallel kernel
 * (error if it does not fit).  The walker needs a workspace of dcarl_stream_bounds_ws() bytes;
 * dcarl_stream_rows_fit() tells whether variant 0 will use the row-parallel kernel (1) or the walker (0).
 */
size_t dcarl_stream_bounds_ws(int64_t n_inst, int64_t n_steps, int32_t S, int32_t A);
int dcarl_stream_rows_fit(int32_t S, int32_t A);
int dcarl_stream_bounds(const int32_t *idx, const int32_t *act, const void *R, int r_is_f64,
                        int64_t n_inst, int64_t n_steps, int32_t S, int32_t A,
                        const dcarl_params *p /* host */,
                        void *out_value, int out_is_f64, int32_t *out_act,
                        int64_t *out_activation_step, void *out_overall,
                        uint64_t *act_hist, uint64_t *bad_rows,
                        void *ws, size_t ws_bytes, int variant, dcarl_stream_t stream);

/* Compact wire format of the same call (ST2:72-105 rows whose keys fit one byte, S <= 256 and
 * A <= 256): idx and act are uint8, R float32, out_value / out_overall (nullable) float32 and
 * out_act uint8 — 6 bytes in and 9 bytes out per scenario-step instead of 12 + 12, which is what
 * a host <-> device link carries in the end-to-end path.  Row-parallel kernel only: the S*A table
 * must fit shared memory (dcarl_stream_rows_fit).  Results are identical to dcarl_stream_bounds on
 * the widened keys. */
int dcarl_stream_bounds_u8(const uint8_t *idx, const uint8_t *act, const float *R,
                           int64_t n_inst, int64_t n_steps, int32_t S, int32_t A,
                           const dcarl_params *p /* host */, float *out_value, uint8_t *out_act,
                           int64_t *out_activation_step, float *out_overall,
                           uint64_t *act_hist, uint64_t *bad_rows, dcarl_stream_t stream);

/* The bound of a cell from caller-held moments, exactly as the stream kernel evaluates it:
 * cell i holds n[i] > 0 samples with sum[i] and sumsq[i]; act[i] == rule_act selects the upper
 * bound (ST1:12), anything else min(lower, CI lower) (ST1:16, 24).  out[i] float64; NaN for n <= 0.
 * (n_thres is not applied here: the caller decides whether the cell refreshes, ST1:86.) */
int dcarl_cell_values(const int32_t *n, const double *sum, const double *sumsq, const int32_t *act,
                      int64_t count, const dcarl_params *p /* host */, double *out,
                      dcarl_stream_t stream);

/* Summary statistics of one batch, the vector that is all-reduced across GPUs:
 * stats[0] = sum_i out_value[i, n_steps-1]       stats[1] = sum_i out_overall[i, n_steps-1] (0 if null)
 * stats[2] = # activated (instance,state) pairs  stats[3] = sum of their activation_step
 * stats[4] = n_inst*n_steps                      stats[5+a] = act_hist[a]  (0 if act_hist null)
 * Fixed-order float64 reduction: bitwise reproducible for a given input. */
#define DCARL_STATS_HEADER 5
int dcarl_stream_stats(const void *out_value, int out_is_f64, const void *out_overall,
                       const int64_t *out_activation_step, const uint64_t *act_hist,
                       int64_t n_inst, int64_t n_steps, int32_t S, int32_t A,
                       double *stats /* [DCARL_STATS_HEADER + A] */, dcarl_stream_t stream);
/* The same reduction ending with the sum over ranks (one kernel: local reduction, then the peer-memory
 * all-reduce of dcarl_peer_ctx).  stats_width >= DCARL_STATS_HEADER + A: the vector is zero-padded to
 * that width so that batches mixing shapes with different action counts add up (BASELINE config 5);
 * accumulate != 0 adds this call's local vector to what `stats` already holds before the sum over ranks
 * is taken — the caller chains the shapes of a mixed batch and lets only the LAST call of the chain
 * carry `peers` (the earlier ones pass NULL), so the batch costs one collective. */
int dcarl_stream_stats_reduce(const void *out_value, int out_is_f64, const void *out_overall,
                              const int64_t *out_activation_step, const uint64_t *act_hist,
                              int64_t n_inst, int64_t n_steps, int32_t S, int32_t A,
                              double *stats, int32_t stats_width, int accumulate,
                              const dcarl_peer_ctx *peers /* host, nullable */, dcarl_stream_t stream);

/* ---- a8: the {s,a,R} sampler of Data_Generation, DS:12-17, 45-55 -------------------------
 * Draw i: x = loc + sd*z_state[i]; idx = floor(x/span*S); rejected (out_idx = -1, R = 0) when
 * idx is outside [0,S) (DS:50-51); else R = q_star[idx, act[i]] + sigma*z_ret[i] (DS:9, 55).
 * z_state / z_ret are standard normal deviates supplied by the caller (float32, or float64 when
 * z_is_f64); arithmetic is float64; out_R is float32 (float64 when out_is_f64). */
int dcarl_sample_returns(const void *q_star, int q_is_f64, const void *z_state, const void *z_ret,
                         int z_is_f64, const int32_t *act, int64_t n, int32_t S, int32_t A,
                         double loc, double sd, double span, double sigma,
                         int32_t *out_idx, void *out_R, int out_is_f64, dcarl_stream_t stream);
/* Same draw with on-device Philox4x32-10 + Box-Muller deviates and uniform actions; row i uses
 * counter (i, 0, 0, 0) and key (seed_lo, seed_hi): reproducible for any grid shape. */
int dcarl_sample_returns_philox(const void *q_star, int q_is_f64, uint64_t seed, int64_t n,
                                int32_t S, int32_t A, double loc, double sd, double span,
                                double sigma, int32_t *out_idx, int32_t *out_act, void *out_R,
                                int out_is_f64, dcarl_stream_t stream);

/* ---- a9-a11: box-kernel confidence, RLS:161-181 and RLS.act_test, RLS:120-157 ------------
 * Buffer: M stored rows, buf_state [M, D+1] (D observation dims, then the action) and buf_ret [M] (the
 * return, column 1 of visited_value, RLS:174), both float32 (buf_is_f64 == 0: the throughput layout,
 * 4*(D+2) = 88 bytes per row) or both float64 (buf_is_f64 != 0: the reference's own precision,
 * RLS:195-196 — a float32 copy moves a stored coordinate by up to 4e-6 at |x| ~ 100, enough to flip a
 * query that sits that close to a box face).  Row i covers the closed box [s_i - dist, s_i + dist]
 * evaluated in float64; a query point obs (+) a (RLS:91-93) matches the rows whose box contains it.
 * For each of the Q queries (float64 [Q, D]) and each candidate action a in [0, n_actions) writes the
 * raw moments out_moments[q, a, {0,1,2}] = {n, sum R, sum R^2} (float64; zeroed by the call).
 * One streaming pass over the buffer, independent of Q (eight or more queries on a buffer of 20 observation dims go
 * through a per-dimension bitmap pre-filter, 64 queries per pass, so that the pass stays close to
 * HBM-bound). */
size_t dcarl_box_moments_ws(int32_t Q, int32_t D, int32_t n_actions);
int dcarl_box_moments(const void *buf_state, const void *buf_ret, int buf_is_f64, int64_t M, int32_t D,
                      const double *dist /* [D+1] */, const double *queries /* [Q, D] */,
                      int32_t Q, int32_t n_actions, double *out_moments,
                      void *ws, size_t ws_bytes, dcarl_stream_t stream);
/* One act_test query pass as ONE launch (RLS:120-157 for Q observations): thresholds in the kernel's
 * prologue, the scan, and — in the last CTA to finish — the totals, their sum over the ranks that hold the
 * other row shards (peers != NULL and peers->world > 1; see dcarl_peer_ctx), (mean, var) and the pick.
 * ws: at least dcarl_box_query_ws() bytes, 16-byte aligned, zero-initialised ONCE by the caller and then
 * owned by the calls that use it (same pointer and ws_bytes every time; stream-ordered): it holds a call
 * counter and two accumulators that alternate, each call zeroing the other one for the next call on its way
 * in (no memset and no copy per query).  out_moments [Q, n_actions, 3] holds the totals over all
 * ranks; out_mean_var / out_action as in dcarl_box_select (each nullable). */
/* The same raw moments for queries that each come with a row range [row_lo[q], row_hi[q]) outside of which no
 * row can match (SURVEY next-row f1: e.g. the run of a buffer sorted by one observation dim that lies within the
 * box half-width of the query; RLS.py:61-69 keeps an rtree for this).  [tail_lo, tail_hi) is a second range every
 * query scans (the unsorted rows appended since the sort; empty when tail_hi <= tail_lo).  One launch, one warp
 * per query, exact test as in dcarl_box_moments; out_moments [Q, n_actions, 3] is written, not added to. */
int dcarl_box_moments_ranges(const void *buf_state, const void *buf_ret, int buf_is_f64, int64_t M, int32_t D,
                             const double *dist /* [D+1] */, const double *queries /* [Q, D] */,
                             const int64_t *row_lo /* [Q] */, const int64_t *row_hi /* [Q] */,
                             int64_t tail_lo, int64_t tail_hi, int32_t Q, int32_t n_actions,
                             double *out_moments, dcarl_stream_t stream);
size_t dcarl_box_query_ws(int32_t Q, int32_t D, int32_t n_actions);
int dcarl_box_query(const void *buf_state, const void *buf_ret, int buf_is_f64, int64_t M, int32_t D,
                    const double *dist /* [D+1] */, const double *queries /* [Q, D] */,
                    int32_t Q, int32_t n_actions, const dcarl_rls_params *p /* host */,
                    double *out_moments, double *out_mean_var, int32_t *out_action,
                    void *ws, size_t ws_bytes, const dcarl_peer_ctx *peers /* host, nullable */,
                    dcarl_stream_t stream);
/* Moments -> (mean, var) with the (-1,-1) sentinel for n == 0 (RLS:169-170) and the
 * act_test pick: first candidate 1..n_actions-1 passing the gate and the z-test, else 0.
 * out_mean_var [Q, n_actions, 2] and out_action [Q] are each nullable. */
int dcarl_box_select(const double *moments, int32_t Q, int32_t n_actions,
                     const dcarl_rls_params *p /* host */, double *out_mean_var,
                     int32_t *out_action, dcarl_stream_t stream);

/* ---- a13 / "next" row f2: batched RLS.add_data return construction, RLS:185-215 ----------------
 * E concurrent transition streams, concatenated: stream e owns transitions
 * [ep_offsets[e], ep_offsets[e+1]) of obs (float32 [N, D]), action (int32 [N]), rew (float32 [N]),
 * done (uint8 [N]).  A transition is recorded with its own reward once `lag` (10) newer ones have
 * arrived; a transition with done != 0 flushes everything still queued with its reward times
 * gamma^distance; transitions still queued when the stream stops are not recorded.  Recorded
 * transitions of a stream are a prefix of it.
 *   dcarl_add_data_counts: out_counts[e] = number of recorded transitions of stream e.
 *   dcarl_add_data_rows:   with row_offsets = exclusive scan of the counts and ep_of_row[i] = the
 *     stream of transition i, writes buffer rows out_state (float32 [rows, D+1] = obs (+) action,
 *     RLS:91-93) and out_ret (float32 [rows]) in stream-major, time order. */
int dcarl_add_data_counts(const float *rew, const uint8_t *done, const int64_t *ep_offsets, int64_t E,
                          int32_t lag, int64_t *out_counts, dcarl_stream_t stream);
int dcarl_add_data_rows(const float *obs, const int32_t *action, const float *rew, const uint8_t *done,
                        const int64_t *ep_offsets, const int64_t *row_offsets, const int32_t *ep_of_row,
                        int64_t N, int32_t D, int32_t lag, double gamma, float *out_state, float *out_ret,
                        dcarl_stream_t stream);

/* ---- a14: constant-velocity rollout surrogate (CARLA itself is external) -----
 * state0 float32 [n, 20] = ego + 3 objects x [x, y, vx, vy, yaw]; the ego velocity is scaled
 * by speed_scale[action]; every step advances all four bodies by v*dt, a collision
 * (squared distance <= radius^2) ends the episode with -100, passing pass_x ends it, else the
 * step reward 0.1*sqrt(|v_ego|) accrues.  out_term: 0 horizon, 1 collision, 2 pass. */
int dcarl_rollout_cv(const float *state0, const int32_t *action, int64_t n, int32_t steps,
                     float dt, float radius, float pass_x, const float *speed_scale,
                     int32_t n_actions, float *out_return, int32_t *out_steps, int32_t *out_term,
                     dcarl_stream_t stream);

/* ---- candidate-trajectory rollout: the episode return of every action of a scenario ------------
 * The reference's collector (Data_From_Carla/Agent/drl_library/dqn/dqn_value_collect.py:86-145) locks in
 * one planner candidate per episode (action 0 = the brake trajectory, c+1 = candidate c, JTP:132-140),
 * follows it in CARLA and records (state, action, sum of step rewards).  Here the episode is idealised
 * kinematically on the planner's own outputs (dcarl_frenet_plan: out_paths, out_n_valid, out_cost): the
 * ego is at point k of the trajectory at time k*dt, its speed the distance between consecutive points / dt;
 * the other vehicles move at constant velocity (PRD:96-100).  Per point: reward
 * ticks_per_point * 0.1 * sqrt(v) (TS:403-404; TS = Test_Scenarios/TestScenario_Town03.py); a collision —
 * the planner's own predicate, PRD:49-59, at t = 2, 4, ... < min(len - 1, n_pred - 1) against the front /
 * back circles of the nearest vehicles — replaces it by collision_reward and ends the episode (TS:407-410);
 * y < pass_y ends it (TS:414-415, :196-201); otherwise a speed below stuck_speed for more than stay_thres
 * seconds ends it with reward 0 (TS:419-422, :182-194).  The brake action stands still at the first point
 * over the cheapest candidate's horizon (JTP:147-150).
 *   out_return [n, C+1] float64, out_steps [n, C+1] points consumed,
 *   out_term   [n, C+1] 0 = trajectory ran out, 1 = collision, 2 = pass, 3 = stuck. */
typedef struct dcarl_rollout_params {
    double dt;                 /* JTP:19   0.3 */
    double robot_radius;       /* JTP:29   1 */
    double move_gap;           /* JTP:31   1 */
    double ticks_per_point;    /* dt / simulator tick (TS:84: 0.05 s) = 6 */
    double collision_reward;   /* TS:409   -100 */
    double pass_y;             /* TS:198   73.7 (NaN: never) */
    double stuck_speed;        /* TS:188   0.1 */
    double stay_thres;         /* TS:182   2 s */
    int32_t n_pred;            /* PRD:91   len(np.arange(0, MAXT, DT)) = 15 */
    int32_t obstacles_considered; /* JTP:28 4 */
} dcarl_rollout_params;
void dcarl_default_rollout_params(dcarl_rollout_params *p);   /* host pointer */
int dcarl_rollout_candidates(const double *paths /* [n, C, Tm, 2] */, const int32_t *n_valid /* [n, C] */,
                             const double *cost /* [n, C] */, const double *ego_xy /* [n, 2] */,
                             const double *obs /* [n, max_obs, 5] */, const int32_t *n_obs /* [n] */,
                             int32_t max_obs, int64_t n, int32_t C, int32_t Tm,
                             const dcarl_rollout_params *p /* host */, double *out_return,
                             int32_t *out_steps, int32_t *out_term, dcarl_stream_t stream);

/* ---- "next" row f3: Frenet candidate trajectories + collision-checked selection --------------
 * One planning tick of JunctionTrajectoryPlanner (JTP:103-140) for n independent scenarios.
 *
 * Routes.  A route is the cubic spline x(s), y(s) of SPL:123-178 through its knots (the densified
 * lane centre line, JTP:79-90).  dcarl_spline_build turns P routes, knots concatenated as
 * knots_xy float64 [K, 2] with route r owning rows [knot_off[r], knot_off[r+1]), into
 * coef float64 [K, 9] rows {s, ax, bx, cx, dx, ay, by, cy, dy} (SPL:18-45; natural end conditions).
 *
 * Planning.  Scenario i runs on route path_id[i] from the Frenet start state
 * start[i] = {s0, c_d, c_d_d, c_d_dd, c_speed} (JTP:248-276, 110) with obstacles
 * obs[i, 0..n_obs[i]) = {x, y, vx, vy, yaw} (float64 [n, max_obs, 5]) and the ego position ego_xy[i]
 * (only used to rank obstacles by distance, PRD:64-85).  Candidate c = (id*n_T + iT)*n_v + iv is the
 * quintic d(t) to d_samples[id] and the quartic s(t) to speed tv_samples[iv] over T_samples[iT]
 * (JTP:300-338); its points t = k*dt, k < n_t[iT], are mapped through the spline until one leaves
 * the route (JTP:345-356).  Outputs, C = n_d*n_T*n_v <= 32, Tm = max n_t <= 32:
 *   out_index [n]      0 = no admissible candidate (brake), c+1 = cheapest candidate that passes
 *                      check_paths (JTP:381-394) and predict.check_collision (PRD:21-60); ties in
 *                      cost go to the lower c, like sorted().  -1 = path_id out of range.
 *   out_cost [n, C]    cf (JTP:333);  out_flags [n, C] bit0 = feasible, bit1 = collision-free
 *   out_n_valid [n, C] points of the global path before it leaves the route
 *   out_paths [n, C, Tm, 2] (nullable) global x, y; NaN past out_n_valid
 *   out_traj [n, Tm, 3] x, y, desired speed of the selected trajectory (JTP:154-156); for index 0
 *                      the cheapest candidate with zero speeds (JTP:149-150, the branch without a
 *                      previous trajectory); out_traj_len [n] its number of valid points. */
#define DCARL_FRENET_MAX_D 16
#define DCARL_FRENET_MAX_T 4
#define DCARL_FRENET_MAX_V 8
typedef struct dcarl_frenet_params {
    double d_samples[DCARL_FRENET_MAX_D];  /* JTP:300  np.arange(MAX_LEFT_WIDTH, MAX_RIGHT_WIDTH+1, D_ROAD_W) */
    double T_samples[DCARL_FRENET_MAX_T];  /* JTP:303  np.arange(MINT, MAXT, DT) */
    double tv_samples[DCARL_FRENET_MAX_V]; /* JTP:315  np.arange(target - D_T_S*N, target + D_T_S*N, D_T_S) */
    int32_t n_t[DCARL_FRENET_MAX_T];       /* JTP:308  len(np.arange(0, Ti, DT)) */
    int32_t n_d, n_T, n_v;
    int32_t n_pred;                        /* PRD:91   len(np.arange(0, MAXT, DT)) */
    int32_t obstacles_considered;          /* JTP:28   4 */
    int32_t pad_;
    double dt;                             /* JTP:19   0.3 */
    double target_speed;                   /* JTP:22   30/3.6 */
    double max_speed, max_accel, max_curvature; /* JTP:13-15 */
    double robot_radius, move_gap;         /* JTP:29, 31 */
    double kj, kt, kd, klat, klon;         /* JTP:35-39 */
} dcarl_frenet_params;
void dcarl_default_frenet_params(dcarl_frenet_params *p);   /* host pointer; the shipped constants */
int dcarl_spline_build(const double *knots_xy, const int64_t *knot_off, int32_t n_paths, double *coef,
                       dcarl_stream_t stream);
/* calculate_start_state (JTP:248-276): ego [n, 5] = {x, y, vx, vy, v} projected on route path_id[i]
 * (tools.get_frenet_state, tools.py:138-270, on the route's knots with the zero tangent table of
 * JTP:84) -> out_start [n, 5] = {s0, c_d, c_d_d, 0, c_speed}, the `start` input of dcarl_frenet_plan.
 * NaN for a path_id outside the routes. */
int dcarl_frenet_start(const double *coef, const int64_t *knot_off, int32_t n_paths, const int32_t *path_id,
                       const double *ego, int64_t n, double *out_start, dcarl_stream_t stream);
int dcarl_frenet_plan(const double *coef, const int64_t *knot_off, int32_t n_paths, const int32_t *path_id,
                      const double *start, const double *ego_xy, const double *obs, const int32_t *n_obs,
                      int32_t max_obs, int64_t n, const dcarl_frenet_params *p /* host */,
                      int32_t *out_index, double *out_cost, uint8_t *out_flags, int32_t *out_n_valid,
                      double *out_paths, double *out_traj, int32_t *out_traj_len, dcarl_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* DCARL_B200_H */
