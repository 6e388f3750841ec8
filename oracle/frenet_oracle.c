/*
 * frenet_oracle.c — CPU restatement of the reference's Frenet candidate planner
 * (SURVEY section 8 row f3).  TEST INFRASTRUCTURE ONLY: only tests/, __graft_entry__.smoke()
 * and bench.py's CPU legs may build, load or call this file; dcarl_b200/ never does.
 *
 * Parity status: PINNED against outputs of the reference itself — tools/make_golden_frenet.py
 * drives the unmodified planner (65 scenarios on four routes) and freezes inputs and outputs in
 * tests/golden/frenet.npz; tests/test_oracle_frenet.py checks this file against them: selected
 * index, feasibility and collision flags and truncation lengths exactly, costs / paths / spline
 * coefficients to 1e-9 (the reference solves its spline and polynomial systems with LAPACK via
 * np.linalg.solve, a third-party routine whose pivoting order is not restated here).
 *
 * Reference locations, relative to /root/reference/Simulation_testing/Simulation_Data_Collection/
 * Data_From_Carla/Agent/zzz/:
 *   JTP = JunctionTrajectoryPlanner.py     SPL = cubic_spline_planner.py     PRD = predict.py
 */
#define _USE_MATH_DEFINES
#define _GNU_SOURCE
#include <math.h>
#include <stdint.h>
#include <string.h>
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif

#define FR_MAX_D 16
#define FR_MAX_T 4
#define FR_MAX_V 8
#define FR_MAX_CAND 32
#define FR_MAX_PTS 32
#define FR_MAX_OBS 16

/* Mirrors dcarl_frenet_params in include/dcarl_b200.h field for field. */
typedef struct {
    double d_samples[FR_MAX_D];   /* JTP:300  np.arange(MAX_LEFT_WIDTH, MAX_RIGHT_WIDTH+1, D_ROAD_W) */
    double T_samples[FR_MAX_T];   /* JTP:303  np.arange(MINT, MAXT, DT) */
    double tv_samples[FR_MAX_V];  /* JTP:315  np.arange(target - dts*N, target + dts*N, dts) */
    int32_t n_t[FR_MAX_T];        /* JTP:308  len(np.arange(0, Ti, DT)) */
    int32_t n_d, n_T, n_v;
    int32_t n_pred;               /* PRD:91   len(np.arange(0, MAXT, DT)) */
    int32_t obstacles_considered; /* JTP:28 */
    int32_t pad_;
    double dt;                    /* JTP:19 */
    double target_speed;          /* JTP:22 */
    double max_speed, max_accel, max_curvature;   /* JTP:13-15 */
    double robot_radius, move_gap;                /* JTP:29, 31 */
    double kj, kt, kd, klat, klon;                /* JTP:35-39 */
} fr_params;

/* ---- SPL:18-45 / :123-150: natural cubic spline through (x[i], y[i]); rows of coef are
 * [a, b, c, d] per knot, b and d of the last knot unused (0).  The reference builds the dense
 * (n x n) system and calls np.linalg.solve; the system is tridiagonal, solved here directly. */
static void spline1d(const double *x, const double *a, int n, double *b, double *c, double *d, double *w1, double *w2) {
    /* rows 1..n-2: h[i-1] c[i-1] + 2(h[i-1]+h[i]) c[i] + h[i] c[i+1] = B[i]; c[0] = c[n-1] = 0 */
    c[0] = 0.0;
    c[n - 1] = 0.0;
    if (n > 2) {
        double *diag = w1, *rhs = w2;
        for (int i = 1; i <= n - 2; ++i) {
            const double h0 = x[i] - x[i - 1], h1 = x[i + 1] - x[i];
            diag[i] = 2.0 * (h0 + h1);
            rhs[i] = 3.0 * (a[i + 1] - a[i]) / h1 - 3.0 * (a[i] - a[i - 1]) / h0;   /* SPL:118-119 */
        }
        for (int i = 2; i <= n - 2; ++i) {
            const double sub = x[i] - x[i - 1];          /* A[i, i-1] = h[i-1] */
            const double l = sub / diag[i - 1];
            diag[i] -= l * sub;                          /* A[i-1, i] = h[i-1] as well */
            rhs[i] -= l * rhs[i - 1];
        }
        for (int i = n - 2; i >= 1; --i) {
            const double sup = x[i + 1] - x[i];
            c[i] = (rhs[i] - (i + 1 <= n - 2 ? sup * c[i + 1] : 0.0)) / diag[i];
        }
    }
    for (int i = 0; i < n - 1; ++i) {                    /* SPL:38-43 */
        const double h = x[i + 1] - x[i];
        d[i] = (c[i + 1] - c[i]) / (3.0 * h);
        b[i] = (a[i + 1] - a[i]) / h - h * (c[i + 1] + 2.0 * c[i]) / 3.0;
    }
    b[n - 1] = 0.0;
    d[n - 1] = 0.0;
}

/* coef: [n, 9] rows {s, ax, bx, cx, dx, ay, by, cy, dy}; n <= FR_MAX_KNOTS.  Returns 0, or -1 if n is out of range. */
#define FR_MAX_KNOTS 4096
int orc_spline_build(const double *xy, int n, double *coef) {
    static _Thread_local double s[FR_MAX_KNOTS], w1[FR_MAX_KNOTS], w2[FR_MAX_KNOTS];
    static _Thread_local double ax[FR_MAX_KNOTS], bx[FR_MAX_KNOTS], cx[FR_MAX_KNOTS], dxx[FR_MAX_KNOTS];
    static _Thread_local double ay[FR_MAX_KNOTS], by[FR_MAX_KNOTS], cy[FR_MAX_KNOTS], dyy[FR_MAX_KNOTS];
    if (n < 2 || n > FR_MAX_KNOTS) return -1;
    s[0] = 0.0;
    for (int i = 1; i < n; ++i) {                        /* SPL:130-137: s = cumsum(sqrt(dx^2 + dy^2)) */
        const double dx = xy[2 * i] - xy[2 * (i - 1)], dy = xy[2 * i + 1] - xy[2 * (i - 1) + 1];
        s[i] = s[i - 1] + sqrt(dx * dx + dy * dy);
    }
    for (int i = 0; i < n; ++i) { ax[i] = xy[2 * i]; ay[i] = xy[2 * i + 1]; }
    spline1d(s, ax, n, bx, cx, dxx, w1, w2);
    spline1d(s, ay, n, by, cy, dyy, w1, w2);
    for (int i = 0; i < n; ++i) {
        double *r = coef + 9 * (size_t)i;
        r[0] = s[i];
        r[1] = ax[i]; r[2] = bx[i]; r[3] = cx[i]; r[4] = dxx[i];
        r[5] = ay[i]; r[6] = by[i]; r[7] = cy[i]; r[8] = dyy[i];
    }
    return 0;
}

/* SPL:47-66, :103-107: segment index by bisect_right(x, t) - 1; None outside [x0, x_last].
 * t == x_last indexes past the last segment in the reference (IndexError); treated as outside. */
static int spline_seg(const double *coef, int n, double t) {
    if (t < coef[0] || !(t < coef[9 * (size_t)(n - 1)])) return -1;
    int lo = 0, hi = n;   /* first index with s[idx] > t */
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (t < coef[9 * (size_t)mid]) hi = mid; else lo = mid + 1;
    }
    return lo - 1;
}

/* ---- JTP:397-443 quintic, :446-485 quartic.  The reference solves the 3x3 / 2x2 systems with
 * np.linalg.solve; here Gaussian elimination with partial pivoting on the same matrices. */
static void solve3(double A[3][3], double b[3], double x[3]) {
    int p[3] = {0, 1, 2};
    for (int k = 0; k < 3; ++k) {
        int piv = k;
        for (int i = k + 1; i < 3; ++i)
            if (fabs(A[p[i]][k]) > fabs(A[p[piv]][k])) piv = i;
        const int t = p[k]; p[k] = p[piv]; p[piv] = t;
        for (int i = k + 1; i < 3; ++i) {
            const double l = A[p[i]][k] / A[p[k]][k];
            for (int j = k; j < 3; ++j) A[p[i]][j] -= l * A[p[k]][j];
            b[p[i]] -= l * b[p[k]];
        }
    }
    for (int k = 2; k >= 0; --k) {
        double v = b[p[k]];
        for (int j = k + 1; j < 3; ++j) v -= A[p[k]][j] * x[j];
        x[k] = v / A[p[k]][k];
    }
}

static void quintic(double xs, double vxs, double axs, double xe, double vxe, double axe, double T, double a[6]) {
    a[0] = xs; a[1] = vxs; a[2] = axs / 2.0;
    const double T2 = T * T, T3 = T2 * T, T4 = T3 * T, T5 = T4 * T;
    double A[3][3] = {{T3, T4, T5}, {3 * T2, 4 * T3, 5 * T4}, {6 * T, 12 * T2, 20 * T3}};
    double b[3] = {xe - a[0] - a[1] * T - a[2] * T2, vxe - a[1] - 2 * a[2] * T, axe - 2 * a[2]};
    double x[3];
    solve3(A, b, x);
    a[3] = x[0]; a[4] = x[1]; a[5] = x[2];
}

static void quartic(double xs, double vxs, double axs, double vxe, double axe, double T, double a[5]) {
    a[0] = xs; a[1] = vxs; a[2] = axs / 2.0;
    const double T2 = T * T, T3 = T2 * T;
    /* [[3T^2, 4T^3], [6T, 12T^2]]; partial pivoting: |6T| > |3T^2| iff T < 2 */
    double A[2][2] = {{3 * T2, 4 * T3}, {6 * T, 12 * T2}};
    double b[2] = {vxe - a[1] - 2 * a[2] * T, axe - 2 * a[2]};
    int r0 = 0, r1 = 1;
    if (fabs(A[1][0]) > fabs(A[0][0])) { r0 = 1; r1 = 0; }
    const double l = A[r1][0] / A[r0][0];
    const double u11 = A[r1][1] - l * A[r0][1];
    const double y1 = b[r1] - l * b[r0];
    a[4] = y1 / u11;
    a[3] = (b[r0] - A[r0][1] * a[4]) / A[r0][0];
}

static double qp_point5(const double a[6], double t) {   /* JTP:423-427, Python float ** int -> pow */
    return a[0] + a[1] * t + a[2] * pow(t, 2) + a[3] * pow(t, 3) + a[4] * pow(t, 4) + a[5] * pow(t, 5);
}
static double qp_d3_5(const double a[6], double t) { return 6 * a[3] + 24 * a[4] * t + 60 * a[5] * pow(t, 2); }
static double qp_point4(const double a[5], double t) {
    return a[0] + a[1] * t + a[2] * pow(t, 2) + a[3] * pow(t, 3) + a[4] * pow(t, 4);
}
static double qp_d1_4(const double a[5], double t) { return a[1] + 2 * a[2] * t + 3 * a[3] * pow(t, 2) + 4 * a[4] * pow(t, 3); }
static double qp_d2_4(const double a[5], double t) { return 2 * a[2] + 6 * a[3] * t + 12 * a[4] * pow(t, 2); }
static double qp_d3_4(const double a[5], double t) { return 6 * a[3] + 24 * a[4] * t; }

/*
 * One planning tick for n independent scenarios (JTP:103-140 with the helpers it calls).
 *   coef [K,9], knot_off [P+1]: the routes' splines; path_id [n]
 *   start [n,5] = {s0, c_d, c_d_d, c_d_dd, c_speed}   (JTP:248-276, :110)
 *   ego_xy [n,2]; obs [n, max_obs, 5] = {x, y, vx, vy, yaw}; n_obs [n]
 * Outputs (C = n_d*n_T*n_v candidates, Tm = max n_t):
 *   index [n]: 0 = no candidate (brake), i+1 = candidate i     (JTP:123-130)
 *   cost [n,C] = cf; flags [n,C]: bit0 feasible (check_paths), bit1 collision-free; n_valid [n,C]
 *   paths [n,C,Tm,2] (nullable): global x, y, NaN past n_valid
 *   traj [n,Tm,3]: x, y, desired speed of the selected trajectory (JTP:132-156; index 0 = the
 *   lowest-cost candidate with zero speeds, the stateless branch of get_backup_trajectory), traj_len [n]
 */
void orc_frenet_plan(const double *coef, const int64_t *knot_off, const int32_t *path_id, const double *start,
                     const double *ego_xy, const double *obs, const int32_t *n_obs, int max_obs, int64_t n,
                     const fr_params *p, int32_t *index, double *cost, uint8_t *flags, int32_t *n_valid,
                     double *paths, double *traj, int32_t *traj_len) {
    const int C = p->n_d * p->n_T * p->n_v;
    int Tm = 0;
    for (int i = 0; i < p->n_T; ++i) if (p->n_t[i] > Tm) Tm = p->n_t[i];
    for (int64_t sc = 0; sc < n; ++sc) {
        const double *st = start + 5 * sc;
        const double s0 = st[0], c_d = st[1], c_d_d = st[2], c_d_dd = st[3], c_speed = st[4];
        const int64_t k0 = knot_off[path_id[sc]];
        const int nk = (int)(knot_off[path_id[sc] + 1] - k0);
        const double *cf = coef + 9 * k0;
        double X[FR_MAX_CAND][FR_MAX_PTS], Y[FR_MAX_CAND][FR_MAX_PTS], SD[FR_MAX_CAND][FR_MAX_PTS];
        int NV[FR_MAX_CAND], NT[FR_MAX_CAND];
        uint8_t FL[FR_MAX_CAND];
        double CF[FR_MAX_CAND];

        /* PRD:64-110: the obstacles_considered nearest vehicles (stable sort by distance to ego),
         * a front and a back circle each, constant velocity */
        int order[FR_MAX_OBS];
        const int no = n_obs[sc] < max_obs ? n_obs[sc] : max_obs;
        double dist[FR_MAX_OBS];
        for (int i = 0; i < no; ++i) {
            const double *o = obs + ((size_t)sc * max_obs + i) * 5;
            const double dx = o[0] - ego_xy[2 * sc], dy = o[1] - ego_xy[2 * sc + 1];
            dist[i] = sqrt(dx * dx + dy * dy);
            order[i] = i;
        }
        for (int i = 1; i < no; ++i) {   /* insertion sort: stable */
            const int oi = order[i];
            int j = i - 1;
            while (j >= 0 && dist[order[j]] > dist[oi]) { order[j + 1] = order[j]; --j; }
            order[j + 1] = oi;
        }
        const int nuse = no < p->obstacles_considered ? no : p->obstacles_considered;

        int ci = 0;
        for (int id = 0; id < p->n_d; ++id) {
            for (int iT = 0; iT < p->n_T; ++iT) {
                const double Ti = p->T_samples[iT];
                const int nt = p->n_t[iT];
                double lat[6];
                quintic(c_d, c_d_d, c_d_dd, p->d_samples[id], 0.0, 0.0, Ti, lat);          /* JTP:306 */
                for (int iv = 0; iv < p->n_v; ++iv, ++ci) {
                    double lon[5];
                    quartic(s0, c_speed, 0.0, p->tv_samples[iv], 0.0, Ti, lon);             /* JTP:318 */
                    double Jp = 0.0, Js = 0.0, d_last = 0.0, sd_last = 0.0;
                    int feasible = 1, nv = -1;
                    for (int it = 0; it < nt; ++it) {
                        const double t = it * p->dt;                                         /* np.arange(0, Ti, DT) */
                        const double d = qp_point5(lat, t), s = qp_point4(lon, t);
                        const double s_d = qp_d1_4(lon, t), s_dd = qp_d2_4(lon, t);
                        const double j5 = qp_d3_5(lat, t), j4 = qp_d3_4(lon, t);
                        Jp += j5 * j5;                                                       /* JTP:325-326: sum(), in order */
                        Js += j4 * j4;
                        d_last = d; sd_last = s_d;
                        SD[ci][it] = s_d;
                        if (s_d > p->max_speed) feasible = 0;                                /* JTP:385 */
                        if (fabs(s_dd) > p->max_accel) feasible = 0;                         /* JTP:387 */
                        if (nv < 0) {                                                        /* JTP:345-356 */
                            const int sg = spline_seg(cf, nk, s);
                            if (sg < 0) { nv = it; }
                            else {
                                const double *r = cf + 9 * (size_t)sg;
                                const double u = s - r[0];
                                const double ix = r[1] + r[2] * u + r[3] * pow(u, 2.0) + r[4] * pow(u, 3.0);
                                const double iy = r[5] + r[6] * u + r[7] * pow(u, 2.0) + r[8] * pow(u, 3.0);
                                const double ddx = r[2] + 2.0 * r[3] * u + 3.0 * r[4] * pow(u, 2.0);
                                const double ddy = r[6] + 2.0 * r[7] * u + 3.0 * r[8] * pow(u, 2.0);
                                const double yaw = atan2(ddy, ddx);
                                X[ci][it] = ix + d * cos(yaw + M_PI / 2.0);
                                Y[ci][it] = iy + d * sin(yaw + M_PI / 2.0);
                            }
                        }
                    }
                    if (nv < 0) nv = nt;
                    const double ds = (p->target_speed - sd_last) * (p->target_speed - sd_last);
                    const double cd = p->kj * Jp + p->kt * Ti + p->kd * pow(d_last, 2);     /* JTP:331 */
                    const double cv = p->kj * Js + p->kt * Ti + p->kd * ds;                 /* JTP:332 */
                    CF[ci] = p->klat * cd + p->klon * cv;
                    NV[ci] = nv; NT[ci] = nt;
                    /* JTP:358-377: yaw / ds of the global path, curvature check of check_paths (:389) */
                    if (nv >= 2) {
                        double yw[FR_MAX_PTS], dl[FR_MAX_PTS];
                        for (int i = 0; i + 1 < nv; ++i) {
                            const double dx = X[ci][i + 1] - X[ci][i], dy = Y[ci][i + 1] - Y[ci][i];
                            yw[i] = atan2(dy, dx);
                            dl[i] = sqrt(dx * dx + dy * dy);
                        }
                        yw[nv - 1] = yw[nv - 2];
                        dl[nv - 1] = dl[nv - 2];
                        for (int i = 0; i + 1 < nv; ++i) {
                            if (dl[i] < 0.00001) dl[i] = 0.1;
                            const double c = (yw[i + 1] - yw[i]) / dl[i];
                            if (fabs(c) > p->max_curvature) feasible = 0;
                        }
                    }
                    /* PRD:21-60: collision-free unless some circle comes within the radius at a
                     * sampled step t = 2, 4, ... < min(len(fp.x)-1, len(path.t)-1) */
                    int freeflag = 1;
                    if (nuse > 0 && nt >= 2) {
                        int lim = nv - 1 < p->n_pred - 1 ? nv - 1 : p->n_pred - 1;
                        for (int q = 0; q < nuse && freeflag; ++q) {
                            const double *o = obs + ((size_t)sc * max_obs + order[q]) * 5;
                            const double gx = cos(o[4]) * 1.0 * p->move_gap, gy = sin(o[4]) * 1.0 * p->move_gap;
                            for (int sgn = 0; sgn < 2 && freeflag; ++sgn) {
                                for (int t = 2; t < lim; t += 2) {
                                    const double xp = o[0] + (t * p->dt) * o[2], yp = o[1] + (t * p->dt) * o[3];
                                    const double px = sgn ? xp - gx : xp + gx, py = sgn ? yp - gy : yp + gy;
                                    const double dd = (px - X[ci][t]) * (px - X[ci][t]) + (py - Y[ci][t]) * (py - Y[ci][t]);
                                    if (dd <= p->robot_radius * p->robot_radius) { freeflag = 0; break; }
                                }
                            }
                        }
                    }
                    FL[ci] = (uint8_t)((feasible ? 1 : 0) | (freeflag ? 2 : 0));
                }
            }
        }
        /* JTP:125-130: candidates by ascending cost (stable), first feasible and collision-free */
        int best = -1, lowest = 0;
        for (int i = 1; i < C; ++i) if (CF[i] < CF[lowest]) lowest = i;
        for (int i = 0; i < C; ++i) {
            if (FL[i] != 3) continue;
            if (best < 0 || CF[i] < CF[best]) best = i;
        }
        index[sc] = best + 1;
        const int pick = best >= 0 ? best : lowest;
        for (int i = 0; i < C; ++i) {
            cost[sc * C + i] = CF[i];
            flags[sc * C + i] = FL[i];
            n_valid[sc * C + i] = NV[i];
            if (paths) {
                for (int t = 0; t < Tm; ++t) {
                    double *o = paths + (((size_t)sc * C + i) * Tm + t) * 2;
                    o[0] = t < NV[i] ? X[i][t] : NAN;
                    o[1] = t < NV[i] ? Y[i][t] : NAN;
                }
            }
        }
        traj_len[sc] = NV[pick];
        for (int t = 0; t < Tm; ++t) {
            double *o = traj + ((size_t)sc * Tm + t) * 3;
            o[0] = t < NV[pick] ? X[pick][t] : NAN;
            o[1] = t < NV[pick] ? Y[pick][t] : NAN;
            o[2] = t < NT[pick] ? (best >= 0 ? SD[pick][t] : 0.0) : NAN;
        }
    }
}
