/*
 * dcarl_oracle.c — CPU restatement of the DCARL hot path.  TEST INFRASTRUCTURE ONLY.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call this file.  The product (dcarl_b200/) never does.
 *
 * Parity status: PINNED.  Every function here is checked bit-for-bit against outputs
 * of the reference itself, produced by tools/make_golden.py running the unmodified
 * reference scripts in the build container (the .npz files under tests/golden):
 *   - bound functions      vs tests/golden/bounds.npz           (tests/test_oracle.py)
 *   - stream loop          vs tests/golden/sim1.npz, sim2.npz   (sha256 of the whole trace)
 *   - box query            vs tests/golden/rls_box.npz          (107 logged field queries)
 *   - sampler              vs tests/golden/sampling.npz         (seeded run of Data_Generation)
 *
 * Reference locations (relative to /root/reference):
 *   ST1 = Simulation_testing/Simulation_1/test_DCARL.py
 *   ST2 = Simulation_testing/Simulation_2/test_DCARL.py
 *   DS  = Simulation_testing/Simulation_Data_Collection/Data_Sampling/data_sampling.py
 *   RLS = Field_testing/Software_and_Raw_Data_on_Self-Driving_Vehicle/software/src/tools/
 *         DCARL/stable_baselines/deepq/RLS.py
 *
 * The reference computes with numpy float64.  To be bit-exact the summation below
 * restates numpy's published pairwise-summation scheme for add.reduce on a
 * contiguous float64 vector (numpy 2.3, not vendored in the reference):
 *   reduce:  out = pairwise(a, n)   (verified bit-exact against np.sum for n = 1..20000)
 *   pairwise(n<8): plain loop; (n<=128): 8 partial sums combined as a balanced tree,
 *   then the tail; (n>128): split at n/2 rounded down to a multiple of 8.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    double alpha;            /* ST1:10  alpha=0.05 */
    double scale;            /* ST1:10  scale=150 */
    double cap;              /* ST1:12  min(100, .) on the rule action */
    double prior_rule;       /* ST1:52  100 */
    double prior_other;      /* ST1:51  -50 */
    double activation_bonus; /* ST2:103 "- activation_value[i]*0.9" with activation_value=-1 */
    int32_t n_thres;         /* ST1:45  10 (strict >) */
    int32_t rule_act;        /* ST1:43  0 */
} orc_params;

/* ---- numpy float64 add.reduce --------------------------------------------------- */
static double pairwise(const double *a, long n)
{
    if (n < 8) {
        double res = 0.0;
        for (long i = 0; i < n; i++) res += a[i];
        return res;
    } else if (n <= 128) {
        double r[8];
        long i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8)
            for (int j = 0; j < 8; j++) r[j] += a[i + j];
        double res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    } else {
        long n2 = n / 2;
        n2 -= n2 % 8;
        return pairwise(a, n2) + pairwise(a + n2, n - n2);
    }
}

double orc_np_sum(const double *a, long n)
{
    if (n <= 0) return 0.0;
    return pairwise(a, n);
}

double orc_np_mean(const double *a, long n) { return orc_np_sum(a, n) / (double)n; }

/* np.std (ddof=0): two-pass, sum((x-mean)^2)/n, sqrt.  tmp must hold n doubles. */
double orc_np_var(const double *a, long n, double *tmp)
{
    double m = orc_np_sum(a, n) / (double)n;
    for (long i = 0; i < n; i++) { double d = a[i] - m; tmp[i] = d * d; }
    return orc_np_sum(tmp, n) / (double)n;
}

/* ---- the four bound functions, ST1:10-28 (== ST2:9-27) --------------------------- */
double orc_upper_bound(const double *x, long n, double alpha, double scale, double cap)
{
    double v = orc_np_mean(x, n) + scale * sqrt(log(1 / alpha) / 2 / (double)n);
    return cap < v ? cap : (v < cap ? v : cap); /* python min(100, v) */
}

double orc_lower_bound(const double *x, long n, double alpha, double scale)
{
    return orc_np_mean(x, n) - scale * sqrt(log(1 / alpha) / 2 / (double)n);
}

double orc_ci_lower_bound(const double *x, long n, double alpha, double scale, double *tmp)
{
    double dn = (double)n, dn1 = (double)(n + 1);
    double dsum = orc_np_sum(x, n);
    double sigma = sqrt(orc_np_var(x, n, tmp));
    /* ST1:24, evaluated left to right */
    return dsum / dn / dn1 - 4 * sigma / dn1 + dsum / dn1 - scale * sqrt(log(1 / alpha) / 2 / dn1);
}

double orc_mean_value(const double *x, long n, double cap)
{
    double v = orc_np_mean(x, n);
    return v < cap ? v : cap;
}

/* ---- streaming estimator + selector, ST1:73-99 / ST2:72-105 ---------------------- */
typedef struct { double *v; long n, cap; } cell_t;

static int first_argmax(const double *v, int A, double *mx)
{
    int best = 0;
    for (int a = 1; a < A; a++) if (v[a] > v[best]) best = a;
    *mx = v[best];
    return best;
}

/* One scenario instance.  idx/act/R: the per-row stream (ST1:73).  Outputs are in row
 * (time) order.  out_overall / out_final may be NULL.  incremental != 0 replaces the
 * O(n) re-reduction of the cell by running sums (same algebra the GPU uses; NOT
 * bit-exact with numpy, used for the fair-CPU timing and as a cross-check). */
int orc_stream(const int32_t *idx, const int32_t *act, const double *R, long n_steps,
               int S, int A, const orc_params *p, int incremental,
               double *out_value, int32_t *out_act, int64_t *out_activation_step,
               double *out_overall, double *out_final)
{
    long ncell = (long)S * A;
    cell_t *cells = (cell_t *)calloc((size_t)ncell, sizeof(cell_t));
    double *val = (double *)malloc(sizeof(double) * (size_t)ncell);
    double *s1 = (double *)calloc((size_t)ncell, sizeof(double));
    double *s2 = (double *)calloc((size_t)ncell, sizeof(double));
    long *cnt = (long *)calloc((size_t)ncell, sizeof(long));
    long *steps_in_state = (long *)calloc((size_t)S, sizeof(long));
    double *tmp = NULL;
    long tmp_cap = 0;
    int rc = 0;
    double L = log(1 / p->alpha);
    for (int s = 0; s < S; s++) {
        for (int a = 0; a < A; a++) val[(long)s * A + a] = p->prior_other;   /* ST1:51 */
        val[(long)s * A + p->rule_act] = p->prior_rule;                      /* ST1:52 */
        out_activation_step[s] = -1;                                         /* ST1:58 */
    }
    for (long k = 0; k < n_steps; k++) {
        int s = idx[k], a = act[k];
        if (s < 0 || s >= S || a < 0 || a >= A) { rc = -1; break; }
        long c = (long)s * A + a;
        double v = 0;
        long n;
        if (!incremental) {
            cell_t *ce = &cells[c];
            if (ce->n == ce->cap) {
                ce->cap = ce->cap ? ce->cap * 2 : 16;
                ce->v = (double *)realloc(ce->v, sizeof(double) * (size_t)ce->cap);
            }
            ce->v[ce->n++] = R[k];                                           /* ST1:80 */
            n = ce->n;
            if (n > p->n_thres) {                                            /* ST1:86 */
                if (n > tmp_cap) { tmp_cap = n * 2; tmp = (double *)realloc(tmp, sizeof(double) * (size_t)tmp_cap); }
                if (a == p->rule_act) {
                    v = orc_upper_bound(ce->v, n, p->alpha, p->scale, p->cap);   /* ST1:88 */
                } else {
                    double lb = orc_lower_bound(ce->v, n, p->alpha, p->scale);
                    double ci = orc_ci_lower_bound(ce->v, n, p->alpha, p->scale, tmp);
                    v = lb < ci ? lb : (ci < lb ? ci : lb);                  /* ST1:90 min(lb, ci) */
                }
            }
        } else {
            n = ++cnt[c];
            s1[c] += R[k];
            s2[c] += R[k] * R[k];
            if (n > p->n_thres) {
                double dn = (double)n, dn1 = (double)(n + 1);
                double mean = s1[c] / dn;
                if (a == p->rule_act) {
                    v = mean + p->scale * sqrt(L / 2 / dn);
                    if (!(v < p->cap)) v = p->cap;
                } else {
                    double var = s2[c] / dn - mean * mean;
                    double sigma = var > 0 ? sqrt(var) : 0.0;
                    double lb = mean - p->scale * sqrt(L / 2 / dn);
                    double ci = s1[c] / dn / dn1 - 4 * sigma / dn1 + s1[c] / dn1 - p->scale * sqrt(L / 2 / dn1);
                    v = lb < ci ? lb : ci;
                }
            }
        }
        if (n > p->n_thres) val[c] = v;
        double mx;
        int am = first_argmax(val + (long)s * A, A, &mx);                    /* ST1:93-94 */
        out_value[k] = mx;
        out_act[k] = am;
        steps_in_state[s]++;
        if (out_activation_step[s] == -1 && am != p->rule_act)               /* ST1:98-99 */
            out_activation_step[s] = steps_in_state[s];
        if (out_overall) {                                                   /* ST2:99-105 */
            double ov = 0;
            for (int i = 0; i < S; i++) {
                if (out_activation_step[i] != -1) {
                    double mi;
                    first_argmax(val + (long)i * A, A, &mi);
                    ov = ov + mi - (-1 * p->activation_bonus);
                }
            }
            out_overall[k] = ov;
        }
    }
    if (out_final) memcpy(out_final, val, sizeof(double) * (size_t)ncell);
    for (long c = 0; c < ncell; c++) free(cells[c].v);
    free(cells); free(val); free(s1); free(s2); free(cnt); free(steps_in_state); free(tmp);
    return rc;
}

/* Many instances, row-major [n_inst, n_steps] streams; a plain loop (callers thread it). */
int orc_stream_batch(const int32_t *idx, const int32_t *act, const double *R, long n_inst,
                     long n_steps, int S, int A, const orc_params *p, int incremental,
                     double *out_value, int32_t *out_act, int64_t *out_activation_step,
                     double *out_overall)
{
    for (long i = 0; i < n_inst; i++) {
        int rc = orc_stream(idx + i * n_steps, act + i * n_steps, R + i * n_steps, n_steps, S, A,
                            p, incremental, out_value + i * n_steps, out_act + i * n_steps,
                            out_activation_step + i * S,
                            out_overall ? out_overall + i * n_steps : NULL, NULL);
        if (rc) return rc;
    }
    return 0;
}

/* ---- box-kernel confidence, RLS:161-181 ----------------------------------------- */
/* Stored row i covers the closed box [s_i - dist, s_i + dist] (RLS:195-196, computed in
 * float64 at insert time); rtree.intersection(point) returns the boxes that contain the
 * point.  For each query obs (D dims) and each candidate action a (appended as dim D,
 * RLS:91-93) return n, mean and population variance (np.mean / np.var, RLS:177-178) of
 * the matched returns; (n=0) -> mean=var=-1 (RLS:169-170). */
int orc_box_moments(const double *buf_state, const double *buf_ret, long M, int D,
                    const double *dist, const double *queries, int Q, int n_actions,
                    double *out_nmv)
{
    double *vals = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1));
    double *tmp = (double *)malloc(sizeof(double) * (size_t)(M > 0 ? M : 1));
    int W = D + 1;
    for (int q = 0; q < Q; q++) {
        for (int a = 0; a < n_actions; a++) {
            long n = 0;
            for (long i = 0; i < M; i++) {
                const double *s = buf_state + i * W;
                int ok = 1;
                for (int d = 0; d < W && ok; d++) {
                    double x = d < D ? queries[(long)q * D + d] : (double)a;
                    ok = (s[d] - dist[d] <= x) && (x <= s[d] + dist[d]);
                }
                if (ok) vals[n++] = buf_ret[i];
            }
            double *o = out_nmv + ((long)q * n_actions + a) * 3;
            o[0] = (double)n;
            if (n == 0) { o[1] = -1; o[2] = -1; }
            else { o[1] = orc_np_mean(vals, n); o[2] = orc_np_var(vals, n, tmp); }
        }
    }
    free(vals); free(tmp);
    return 0;
}

/* Same box test (RLS:161-163, 195-196) restated for large buffers: ONE pass over the stored rows per
 * query instead of one per (query, candidate action).  A stored row whose observation box contains
 * the query's observation matches exactly the candidate actions a with
 * s_act - dist[D] <= a <= s_act + dist[D] (the action is just the last coordinate of the 21-D point,
 * RLS:91-93), so the count per (query, action) is identical to orc_box_moments.  The buffer is the
 * float32 copy the device holds, widened to double per element (buf_is_f32 != 0), or the float64
 * rows themselves.  Output: raw moments out[q, a, {0,1,2}] = {n, sum R, sum R^2} in double, summed in
 * row order — counts are exact, the sums are compared with a tolerance (tests/test_oracle.py pins this
 * function against orc_box_moments). */
int orc_box_raw(const void *buf_state, const void *buf_ret, int buf_is_f32, long M, int D,
                const double *dist, const double *queries, int Q, int n_actions, double *out)
{
    const int W = D + 1;
    const float *sf = (const float *)buf_state, *rf = (const float *)buf_ret;
    const double *sd = (const double *)buf_state, *rd = (const double *)buf_ret;
    for (long i = 0; i < (long)Q * n_actions * 3; i++) out[i] = 0.0;
    for (int q = 0; q < Q; q++) {
        const double *x = queries + (long)q * D;
        double *o = out + (long)q * n_actions * 3;
        for (long i = 0; i < M; i++) {
            int ok = 1;
            for (int d = 0; d < D && ok; d++) {
                const double s = buf_is_f32 ? (double)sf[i * W + d] : sd[i * W + d];
                ok = (s - dist[d] <= x[d]) && (x[d] <= s + dist[d]);
            }
            if (!ok) continue;
            const double sa = buf_is_f32 ? (double)sf[i * W + D] : sd[i * W + D];
            const double r = buf_is_f32 ? (double)rf[i] : rd[i];
            const double lo = sa - dist[D], hi = sa + dist[D];
            if (!(lo == lo) || !(hi == hi)) continue;
            long a0 = (long)ceil(lo), a1 = (long)floor(hi);
            if (a0 < 0) a0 = 0;
            if (a1 > n_actions - 1) a1 = n_actions - 1;
            for (long a = a0; a <= a1; a++)
                if (lo <= (double)a && (double)a <= hi) { o[a * 3] += 1.0; o[a * 3 + 1] += r; o[a * 3 + 2] += r * r; }
        }
    }
    return 0;
}

static double norm_cdf(double z) { return 0.5 * erfc(-z * 0.70710678118654752440); }

/* RLS.act_test, RLS:120-157, for Q queries; nmv = [Q, n_actions, 3] from above.
 * Candidates 1..n_actions-1 in order (the reference has n_actions = 8, RLS:135). */
int orc_box_select(const double *nmv, int Q, int n_actions, int visited_times_thres,
                   int min_candidate_visits, double rule_mean_gate, double confidence_thres,
                   int32_t *out_action)
{
    for (int q = 0; q < Q; q++) {
        const double *r = nmv + (long)q * n_actions * 3;
        int pick = 0;
        for (int a = 1; a < n_actions; a++) {
            const double *c = r + a * 3;
            if (r[0] < visited_times_thres || c[0] < min_candidate_visits || r[1] > rule_mean_gate)
                continue;                                                    /* RLS:141 */
            double var_diff = r[2] / r[0] + c[2] / c[0];                    /* RLS:144 */
            double z = (c[1] - r[1]) / sqrt(var_diff);                      /* RLS:145-148 */
            if (norm_cdf(z) > confidence_thres) { pick = a; break; }        /* RLS:150-154 */
        }
        out_action[q] = pick;
    }
    return 0;
}

/* ---- {s,a,R} sampler, DS:12-17, 45-55 with the random deviates supplied ---------- */
/* z_state[n]: standard normals for the state draw; act[n]: uniform action per DRAW
 * (only consumed when the draw is valid); z_ret[n]: standard normals for the return.
 * Writes idx (or -1 when rejected, DS:50-51), R = Q*[idx,act] + sigma*z (DS:9).
 * Returns the number of valid rows. */
long orc_sample_returns(const double *q_star, const double *z_state, const int32_t *act,
                        const double *z_ret, long n, int S, int A, double loc, double sd,
                        double span, double sigma, int32_t *out_idx, double *out_R)
{
    long valid = 0;
    for (long i = 0; i < n; i++) {
        double x = loc + sd * z_state[i];
        double f = floor(x / span * S);                                      /* DS:15 */
        if (f < 0 || f >= S) { out_idx[i] = -1; out_R[i] = 0; continue; }
        int s = (int)f;
        out_idx[i] = s;
        out_R[i] = q_star[(long)s * A + act[i]] + sigma * z_ret[i];
        valid++;
    }
    return valid;
}

/* ---- constant-velocity rollout surrogate ---------------------------- */
/* Semantics restated from Data_From_Carla/Agent/zzz/predict.py:96-100 (x += vx*dt),
 * predict.py:52-59 (collision when squared distance <= r^2), and
 * Test_Scenarios/TestScenario_Town03.py:403-422 (step reward 0.1*sqrt(v), collision
 * -100 & done, pass & done); episode return = sum of step rewards
 * (drl_library/dqn/dqn_value_collect.py:119).  fp32 arithmetic, same op order as
 * the kernel.  state0: [n, 20] = ego + 3 objects x [x, y, vx, vy, yaw].  The ego's
 * speed is scaled by speed_scale[action].  term: 0 = horizon, 1 = collision, 2 = pass. */
int orc_rollout_cv(const float *state0, const int32_t *action, long n, int steps, float dt,
                   float radius, float pass_x, const float *speed_scale, int n_actions,
                   float *out_return, int32_t *out_steps, int32_t *out_term)
{
    for (long i = 0; i < n; i++) {
        const float *s = state0 + i * 20;
        int a = action[i];
        if (a < 0 || a >= n_actions) return -1;
        float ex = s[0], ey = s[1], evx = s[2] * speed_scale[a], evy = s[3] * speed_scale[a];
        float ox[3], oy[3], ovx[3], ovy[3];
        for (int j = 0; j < 3; j++) { ox[j] = s[5 + 5 * j]; oy[j] = s[6 + 5 * j]; ovx[j] = s[7 + 5 * j]; ovy[j] = s[8 + 5 * j]; }
        float v = sqrtf(evx * evx + evy * evy);
        float step_reward = 0.1f * sqrtf(v);
        float ret = 0.0f;
        int t, term = 0;
        for (t = 1; t <= steps; t++) {
            ex = fmaf(evx, dt, ex); ey = fmaf(evy, dt, ey);
            int hit = 0;
            for (int j = 0; j < 3; j++) {
                ox[j] = fmaf(ovx[j], dt, ox[j]); oy[j] = fmaf(ovy[j], dt, oy[j]);
                float dx = ex - ox[j], dy = ey - oy[j];
                float d2 = fmaf(dx, dx, dy * dy);
                hit |= d2 <= radius * radius;
            }
            if (hit) { ret += -100.0f; term = 1; break; }
            ret += step_reward;
            if (ex > pass_x) { term = 2; break; }
        }
        out_return[i] = ret;
        out_steps[i] = t <= steps ? t : steps;
        out_term[i] = term;
    }
    return 0;
}
