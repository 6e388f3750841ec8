// Frenet candidate-trajectory planner, batched over scenarios.
//
// Replaces, for n independent scenarios per call, one planning tick of the reference's
// Simulation_testing/Simulation_Data_Collection/Data_From_Carla/Agent/zzz/ (below: JTP =
// JunctionTrajectoryPlanner.py, SPL = cubic_spline_planner.py, PRD = predict.py):
//   calc_frenet_paths JTP:292-340   quintic lateral / quartic longitudinal candidates and their cost
//   calc_global_paths JTP:342-379   Frenet -> global through the route's cubic spline (SPL:47-178)
//   check_paths       JTP:381-394   speed / acceleration / curvature limits
//   predict           PRD:6-110     constant-velocity circles of the nearest obstacles, collision test
//   get_optimal_trajectory JTP:123-130, get_trajectory_by_index JTP:132-156 (stateless branch)
// and the spline set-up of build_frenet_path / generate_target_course (JTP:79-90, 279-290; SPL:18-45).
//
// The reference walks Python lists candidate by candidate (10 candidates x 14 points, two
// np.linalg.solve per candidate).  Here one warp plans one scenario: the polynomial systems are
// solved by the first lanes, the 140 (candidate, point) pairs are spread over the lanes for the
// spline evaluation and the heading/arc-length pass, one lane per candidate accumulates cost,
// feasibility and the collision test, and a shuffle reduction picks the cheapest admissible
// candidate (ties to the lower index, like sorted()).  float64 throughout.  640 bytes per
// scenario against ~40 k float64 instructions: compute-bound, no bandwidth roofline.
#include <math_constants.h>
#include <string.h>

#include "common.cuh"

namespace dcarl {

static_assert(sizeof(dcarl_frenet_params) % 8 == 0, "params are passed by value");

constexpr int kFrWarpsPerCta = 4;

// ---- SPL:18-45, 123-150: natural cubic splines x(s), y(s) through the knots of one route ---------
// coef rows: {s, ax, bx, cx, dx, ay, by, cy, dy}.  The reference assembles the dense system and
// calls np.linalg.solve; the system is tridiagonal and strictly diagonally dominant: Thomas.
// One warp per route.  The route is staged in shared memory (the serial parts — the running arc length
// and the two elimination sweeps — then cost a shared-memory round trip per knot instead of a DRAM one):
// lanes share the row set-up and the final coefficients, lane 0 solves the x system while lane 1 solves
// the y system.  Every value is computed by the same expression, in the same order, as by the serial
// form below (routes longer than the staging area fall back to it): results are bitwise the same.
__device__ void spline_build_serial(const double *q, double *c, int n) {
    double s = 0.0;
    for (int i = 0; i < n; ++i) {
        if (i) {
            const double dx = q[2 * i] - q[2 * i - 2], dy = q[2 * i + 1] - q[2 * i - 1];
            s += sqrt(dx * dx + dy * dy);                                      // SPL:130-137
        }
        c[9 * i] = s;
        c[9 * i + 1] = q[2 * i];
        c[9 * i + 5] = q[2 * i + 1];
    }
    for (int base = 1; base <= 5; base += 4) {   // x then y; slots: a = base, b = +1, c = +2, d = +3
        double *A = c + base, *B = c + base + 1, *Cc = c + base + 2, *D = c + base + 3;
        // interior rows i = 1..n-2 (diag kept in the b slot, rhs in the d slot while solving)
        for (int i = 1; i <= n - 2; ++i) {
            const double h0 = c[9 * i] - c[9 * (i - 1)], h1 = c[9 * (i + 1)] - c[9 * i];
            B[9 * i] = 2.0 * (h0 + h1);
            D[9 * i] = 3.0 * (A[9 * (i + 1)] - A[9 * i]) / h1 - 3.0 * (A[9 * i] - A[9 * (i - 1)]) / h0;   // SPL:118-119
        }
        for (int i = 2; i <= n - 2; ++i) {
            const double sub = c[9 * i] - c[9 * (i - 1)];
            const double l = sub / B[9 * (i - 1)];
            B[9 * i] -= l * sub;
            D[9 * i] -= l * D[9 * (i - 1)];
        }
        Cc[0] = 0.0;
        Cc[9 * (n - 1)] = 0.0;
        for (int i = n - 2; i >= 1; --i) {
            const double sup = c[9 * (i + 1)] - c[9 * i];
            Cc[9 * i] = (D[9 * i] - (i + 1 <= n - 2 ? sup * Cc[9 * (i + 1)] : 0.0)) / B[9 * i];
        }
        for (int i = 0; i < n - 1; ++i) {                                      // SPL:38-43
            const double h = c[9 * (i + 1)] - c[9 * i];
            D[9 * i] = (Cc[9 * (i + 1)] - Cc[9 * i]) / (3.0 * h);
            B[9 * i] = (A[9 * (i + 1)] - A[9 * i]) / h - h * (Cc[9 * (i + 1)] + 2.0 * Cc[9 * i]) / 3.0;
        }
        B[9 * (n - 1)] = 0.0;
        D[9 * (n - 1)] = 0.0;
    }
}

constexpr int kSplineCols = 7;   // staged per knot: s, a_x, a_y, diag_x, rhs_x -> c_x, diag_y, rhs_y -> c_y share slots

__global__ void __launch_bounds__(32) k_spline_build(const double *__restrict__ xy,
                                                     const long long *__restrict__ knot_off, int n_paths,
                                                     double *__restrict__ coef, int cap) {
    extern __shared__ double spl[];
    const int pth = blockIdx.x, lane = threadIdx.x;
    if (pth >= n_paths) return;
    const long long k0 = knot_off[pth];
    const int n = (int)(knot_off[pth + 1] - k0);
    if (n < 2) return;
    const double *q = xy + 2 * k0;
    double *c = coef + 9 * k0;
    if (n > cap) {                                  // longer than the staging area: the serial form in place
        if (lane == 0) spline_build_serial(q, c, n);
        return;
    }
    double *S = spl, *AX = S + cap, *AY = AX + cap, *BX = AY + cap, *DX = BX + cap, *BY = DX + cap, *DY = BY + cap;
    for (int i = lane; i < n; i += 32) { AX[i] = q[2 * i]; AY[i] = q[2 * i + 1]; }
    __syncwarp();
    for (int i = lane + 1; i < n; i += 32) {        // segment lengths (kept in BX until the running sum is taken)
        const double dx = AX[i] - AX[i - 1], dy = AY[i] - AY[i - 1];
        BX[i] = sqrt(dx * dx + dy * dy);
    }
    __syncwarp();
    if (lane == 0) {                                // SPL:130-137: a running sum, in knot order
        double s = 0.0;
        S[0] = 0.0;
        for (int i = 1; i < n; ++i) { s += BX[i]; S[i] = s; }
    }
    __syncwarp();
    for (int i = lane + 1; i <= n - 2; i += 32) {   // interior rows of both systems
        const double h0 = S[i] - S[i - 1], h1 = S[i + 1] - S[i];
        const double diag = 2.0 * (h0 + h1);
        BX[i] = diag; BY[i] = diag;
        DX[i] = 3.0 * (AX[i + 1] - AX[i]) / h1 - 3.0 * (AX[i] - AX[i - 1]) / h0;   // SPL:118-119
        DY[i] = 3.0 * (AY[i + 1] - AY[i]) / h1 - 3.0 * (AY[i] - AY[i - 1]) / h0;
    }
    __syncwarp();
    if (lane < 2) {                                 // Thomas: lane 0 the x system, lane 1 the y system
        double *B = lane ? BY : BX, *D = lane ? DY : DX;
        for (int i = 2; i <= n - 2; ++i) {
            const double sub = S[i] - S[i - 1];
            const double l = sub / B[i - 1];
            B[i] -= l * sub;
            D[i] -= l * D[i - 1];
        }
        double next = 0.0;                          // c of row i+1 (0 at the natural end)
        for (int i = n - 2; i >= 1; --i) {
            const double sup = S[i + 1] - S[i];
            next = (D[i] - (i + 1 <= n - 2 ? sup * next : 0.0)) / B[i];
            D[i] = next;                            // the rhs slot now holds c
        }
        D[0] = 0.0;
        D[n - 1] = 0.0;
    }
    __syncwarp();
    for (int i = lane; i < n; i += 32) {            // SPL:38-43, and the rows of the output
        double bx = 0.0, dx = 0.0, by = 0.0, dy = 0.0;
        if (i < n - 1) {
            const double h = S[i + 1] - S[i];
            dx = (DX[i + 1] - DX[i]) / (3.0 * h);
            bx = (AX[i + 1] - AX[i]) / h - h * (DX[i + 1] + 2.0 * DX[i]) / 3.0;
            dy = (DY[i + 1] - DY[i]) / (3.0 * h);
            by = (AY[i + 1] - AY[i]) / h - h * (DY[i + 1] + 2.0 * DY[i]) / 3.0;
        }
        double *r = c + 9 * i;
        r[0] = S[i]; r[1] = AX[i]; r[2] = bx; r[3] = DX[i]; r[4] = dx;
        r[5] = AY[i]; r[6] = by; r[7] = DY[i]; r[8] = dy;
    }
}

// ---- JTP:397-485: boundary-value polynomials.  Gaussian elimination with partial pivoting on the
// matrices the reference hands to np.linalg.solve. ---------------------------------------------
__device__ __forceinline__ void quintic_coeffs(double xs, double vxs, double axs, double xe, double vxe,
                                               double axe, double T, double *a) {
    a[0] = xs; a[1] = vxs; a[2] = axs * 0.5;
    const double T2 = T * T, T3 = T2 * T, T4 = T3 * T, T5 = T4 * T;
    double M[3][4] = {{T3, T4, T5, xe - a[0] - a[1] * T - a[2] * T2},
                      {3 * T2, 4 * T3, 5 * T4, vxe - a[1] - 2 * a[2] * T},
                      {6 * T, 12 * T2, 20 * T3, axe - 2 * a[2]}};
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        int piv = k;
#pragma unroll
        for (int i = k + 1; i < 3; ++i)
            if (fabs(M[i][k]) > fabs(M[piv][k])) piv = i;
#pragma unroll
        for (int j = 0; j < 4; ++j) {   // row swap (static indexing keeps M in registers)
            const double t0 = M[k][j];
            double tp = t0;
#pragma unroll
            for (int i = k + 1; i < 3; ++i)
                if (i == piv) { tp = M[i][j]; M[i][j] = t0; }
            M[k][j] = tp;
        }
#pragma unroll
        for (int i = k + 1; i < 3; ++i) {
            const double l = M[i][k] / M[k][k];
#pragma unroll
            for (int j = k; j < 4; ++j) M[i][j] -= l * M[k][j];
        }
    }
    a[5] = M[2][3] / M[2][2];
    a[4] = (M[1][3] - M[1][2] * a[5]) / M[1][1];
    a[3] = (M[0][3] - M[0][1] * a[4] - M[0][2] * a[5]) / M[0][0];
}

__device__ __forceinline__ void quartic_coeffs(double xs, double vxs, double axs, double vxe, double axe,
                                               double T, double *a) {
    a[0] = xs; a[1] = vxs; a[2] = axs * 0.5;
    const double T2 = T * T, T3 = T2 * T;
    double r0[3] = {3 * T2, 4 * T3, vxe - a[1] - 2 * a[2] * T};
    double r1[3] = {6 * T, 12 * T2, axe - 2 * a[2]};
    if (fabs(r1[0]) > fabs(r0[0])) {
#pragma unroll
        for (int j = 0; j < 3; ++j) { const double t = r0[j]; r0[j] = r1[j]; r1[j] = t; }
    }
    const double l = r1[0] / r0[0];
    a[4] = (r1[2] - l * r0[2]) / (r1[1] - l * r0[1]);
    a[3] = (r0[2] - r0[1] * a[4]) / r0[0];
}

// (cost, index) minimum over the warp; ties -> the lower index.  idx < 0 = lane does not compete.
__device__ __forceinline__ int warp_argmin(double cost, int idx) {
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double oc = __shfl_xor_sync(FULL_MASK, cost, off);
        const int oi = __shfl_xor_sync(FULL_MASK, idx, off);
        const bool take = oi >= 0 && (idx < 0 || oc < cost || (oc == cost && oi < idx));
        if (take) { cost = oc; idx = oi; }
    }
    return idx;
}

// ---- calculate_start_state (JTP:248-276): the ego projected on its route ------------------------
// tools.get_frenet_state (tools.py:235-270) over dist_from_point_to_polyline2d (tools.py:138-232) on
// the route's knots (columns ax, ay of coef; column s is the arc length the reference re-sums), with
// the all-zero tangent table the planner passes (JTP:84).  One warp per scenario: the lanes scan
// the knots for the closest one (first minimum, like np.argmin), the case analysis on the one or
// two adjacent segments is evaluated redundantly by every lane.
// ego [n,5] = {x, y, vx, vy, v}  ->  start [n,5] = {s0, c_d = -d, c_d_d = vd, 0, c_speed = v}
__device__ __forceinline__ void point_to_segment(double x0, double y0, double x1, double y1, double x2, double y2,
                                                 double &dl, double &d1, double &d2) {   // tools.py:122-135
    const double l = sqrt((y2 - y1) * (y2 - y1) + (x2 - x1) * (x2 - x1));
    if (l == 0.0) { dl = sqrt((y0 - y1) * (y0 - y1) + (x0 - x1) * (x0 - x1)); d1 = 0.0; d2 = 0.0; return; }
    dl = ((y2 - y1) * x0 - (x2 - x1) * y0 + x2 * y1 - x1 * y2) / l;
    d1 = (x1 * x1 + x0 * (x2 - x1) - x1 * x2 + y1 * y1 + y0 * (y2 - y1) - y1 * y2) / l;
    d2 = (x2 * x2 - x0 * (x2 - x1) - x1 * x2 + y2 * y2 - y0 * (y2 - y1) - y1 * y2) / l;
}

__global__ void __launch_bounds__(kFrWarpsPerCta * 32) k_frenet_start(const double *__restrict__ coef,
                                                                      const long long *__restrict__ knot_off,
                                                                      int n_paths, const int *__restrict__ path_id,
                                                                      const double *__restrict__ ego, long long n,
                                                                      double *__restrict__ start) {
    const int lane = lane_id();
    const int warp = __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0);
    const long long sc = (long long)blockIdx.x * kFrWarpsPerCta + warp;
    if (sc >= n) return;
    const int pid = path_id[sc];
    const bool path_ok = pid >= 0 && pid < n_paths;
    const long long k0 = path_ok ? knot_off[pid] : 0;
    const int nk = path_ok ? (int)(knot_off[pid + 1] - k0) : 0;
    double *o = start + 5 * sc;
    if (nk < 2) {   // warp-uniform
        if (lane < 5) o[lane] = CUDART_NAN;
        return;
    }
    const double *c = coef + 9 * k0;
    const double x0 = ego[5 * sc], y0 = ego[5 * sc + 1], vx = ego[5 * sc + 2], vy = ego[5 * sc + 3], v = ego[5 * sc + 4];
    double best = CUDART_INF;
    int bi = -1;
    for (int i = lane; i < nk; i += 32) {
        const double dx = c[9 * (size_t)i + 1] - x0, dy = c[9 * (size_t)i + 5] - y0;
        const double d = sqrt(dx * dx + dy * dy);
        if (d < best) { best = d; bi = i; }   // ascending i per lane: strict compare keeps the first
    }
    const int ci = warp_argmin(best, bi);
    const double cdx = c[9 * (size_t)ci + 1] - x0, cdy = c[9 * (size_t)ci + 5] - y0;
    const double dmin_ = sqrt(cdx * cdx + cdy * cdy);   // dist_line[closest_idx]
    auto X = [&](int i) { return c[9 * (size_t)i + 1]; };
    auto Y = [&](int i) { return c[9 * (size_t)i + 5]; };
    auto S = [&](int i) { return c[9 * (size_t)i]; };
    double dc = 0.0, nh = 0.0, nt = 0.0, ph = 0.0, pt = 0.0, dn = 0.0, dp = 0.0;
    int ctype = 0;
    if (ci == 0) {
        point_to_segment(x0, y0, X(0), Y(0), X(1), Y(1), dn, nh, nt);
        if (nh < 0) dc = dn < 0 ? -dmin_ : dmin_;
        else { dc = dn; ctype = 1; }
    } else if (ci == nk - 1) {
        point_to_segment(x0, y0, X(nk - 2), Y(nk - 2), X(nk - 1), Y(nk - 1), dp, ph, pt);
        if (pt < 0) dc = dp < 0 ? -dmin_ : dmin_;
        else { dc = dp; ctype = -1; }
    } else {
        point_to_segment(x0, y0, X(ci - 1), Y(ci - 1), X(ci), Y(ci), dp, ph, pt);
        point_to_segment(x0, y0, X(ci), Y(ci), X(ci + 1), Y(ci + 1), dn, nh, nt);
        if (pt < 0 && nh < 0) {
            double sg, u1, u2;
            point_to_segment(X(ci + 1), Y(ci + 1), X(ci - 1), Y(ci - 1), X(ci), Y(ci), sg, u1, u2);
            dc = sg > 0 ? -dmin_ : dmin_;
        } else if (pt < 0) { dc = dn; ctype = 1; }
        else if (nh < 0) { dc = dp; ctype = -1; }
        else if (fabs(dn) > fabs(dp)) { dc = dp; ctype = -1; }
        else { dc = dn; ctype = 1; }
    }
    double s, psi;
    if (ctype == 1) { s = nh + S(ci); psi = atan2(Y(ci + 1) - Y(ci), X(ci + 1) - X(ci)); }
    else if (ctype == -1) { s = ph + S(ci - 1); psi = atan2(Y(ci) - Y(ci - 1), X(ci) - X(ci - 1)); }
    else { s = S(ci); psi = 0.0; }                                             // tangents[nearest_idx] == 0 (JTP:84)
    double sn, cs;
    sincos(psi, &sn, &cs);
    if (lane == 0) {
        o[0] = s;
        o[1] = -dc;                   // JTP:269
        o[2] = -sn * vx + cs * vy;    // tools.py:262-268
        o[3] = 0.0;
        o[4] = v;                     // JTP:263
    }
}

struct FrenetArgs {
    const double *coef;
    const long long *knot_off;
    const int *path_id;
    const double *start, *ego_xy, *obs;
    const int *n_obs;
    int *out_index;
    double *out_cost;
    unsigned char *out_flags;
    int *out_n_valid;
    double *out_paths, *out_traj;
    int *out_traj_len;
    long long n;
    int n_paths, max_obs;
};

__host__ __device__ inline size_t frenet_warp_bytes(int n_lat, int n_lon, int C, int Tm, int n_circ) {
    // lat[n_lat][6], lon[n_lon][5], X/Y/YW/DL[C][Tm], circles[n_circ][4]
    return sizeof(double) * ((size_t)n_lat * 6 + (size_t)n_lon * 5 + (size_t)4 * C * Tm + (size_t)n_circ * 4);
}

#ifndef DCARL_FR_MIN_BLOCKS
#define DCARL_FR_MIN_BLOCKS 8   // 64 registers: 1.64 ms per 262 144 scenarios (79 registers: 1.70, 116: 2.04, 48: 1.70, 40: 1.90)
#endif
__global__ void __launch_bounds__(kFrWarpsPerCta * 32, DCARL_FR_MIN_BLOCKS) k_frenet_plan(const FrenetArgs a, const dcarl_frenet_params p,
                                                                     const int Tm, const unsigned warp_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id();
    const int warp = __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0);
    const long long sc = (long long)blockIdx.x * kFrWarpsPerCta + warp;
    if (sc >= a.n) return;

    const int n_lat = p.n_d * p.n_T, n_lon = p.n_T * p.n_v, C = n_lat * p.n_v;
    double *lat = reinterpret_cast<double *>(smem_raw + (size_t)warp * warp_bytes);
    double *lon = lat + n_lat * 6;
    double *X = lon + n_lon * 5;
    double *Y = X + C * Tm;
    double *YW = Y + C * Tm;
    double *DL = YW + C * Tm;
    double *circ = DL + C * Tm;

    const double *st = a.start + 5 * sc;
    const double s0 = st[0], c_d = st[1], c_d_d = st[2], c_d_dd = st[3], c_speed = st[4];
    const int pid = a.path_id[sc];
    const bool path_ok = pid >= 0 && pid < a.n_paths;
    const long long k0 = path_ok ? a.knot_off[pid] : 0;
    const int nk = path_ok ? (int)(a.knot_off[pid + 1] - k0) : 0;
    const double *cf = a.coef + 9 * k0;
    if (!path_ok || nk < 2) {   // warp-uniform: invalid route -> index -1, NaN outputs
        if (lane == 0) { a.out_index[sc] = -1; a.out_traj_len[sc] = 0; }
        for (int i = lane; i < C; i += 32) {
            a.out_cost[sc * C + i] = CUDART_NAN; a.out_flags[sc * C + i] = 0; a.out_n_valid[sc * C + i] = 0;
        }
        for (int i = lane; i < Tm * 3; i += 32) a.out_traj[sc * Tm * 3 + i] = CUDART_NAN;
        if (a.out_paths)
            for (int i = lane; i < C * Tm * 2; i += 32) a.out_paths[sc * C * Tm * 2 + i] = CUDART_NAN;
        return;
    }

    // ---- polynomial coefficients: one lane per lateral (d, T) pair, then per longitudinal (T, v) pair
    for (int i = lane; i < n_lat; i += 32)
        quintic_coeffs(c_d, c_d_d, c_d_dd, p.d_samples[i / p.n_T], 0.0, 0.0, p.T_samples[i % p.n_T], lat + 6 * i);   // JTP:306
    for (int i = lane; i < n_lon; i += 32)
        quartic_coeffs(s0, c_speed, 0.0, p.tv_samples[i % p.n_v], 0.0, p.T_samples[i / p.n_v], lon + 5 * i);        // JTP:318

    // ---- PRD:64-110: the obstacles_considered nearest obstacles (stable order by distance to the
    // ego), a front and a back circle each: {x0, y0, vx, vy}
    const int no = min(min(a.n_obs[sc], a.max_obs), 32);
    const int nuse = min(no, p.obstacles_considered);
    {
        double ox = 0.0, oy = 0.0, ovx = 0.0, ovy = 0.0, oyaw = 0.0;
        if (lane < no) {   // obs may be null when max_obs == 0
            const double *o = a.obs + ((size_t)sc * a.max_obs + lane) * 5;
            ox = o[0]; oy = o[1]; ovx = o[2]; ovy = o[3]; oyaw = o[4];
        }
        const double dx = ox - a.ego_xy[2 * sc], dy = oy - a.ego_xy[2 * sc + 1];
        const double dist = lane < no ? sqrt(dx * dx + dy * dy) : CUDART_INF;
        int rank = 0;
        for (int j = 0; j < no; ++j) {
            const double dj = __shfl_sync(FULL_MASK, dist, j);
            rank += (dj < dist || (dj == dist && j < lane)) ? 1 : 0;
        }
        if (lane < no && rank < nuse) {
            double sn, cs;
            sincos(oyaw, &sn, &cs);
            const double gx = cs * p.move_gap, gy = sn * p.move_gap;
            double *f = circ + 8 * rank;
            f[0] = ox + gx; f[1] = oy + gy; f[2] = ovx; f[3] = ovy;
            f[4] = ox - gx; f[5] = oy - gy; f[6] = ovx; f[7] = ovy;
        }
    }
    __syncwarp();

    // ---- one lane per candidate: jerk sums, terminal terms, speed / acceleration limits ------------
    // candidate ci = (id * n_T + iT) * n_v + iv   (loop nest of JTP:300-338)
    const int ci = lane;
    const bool has = ci < C;
    const int ilat = has ? ci / p.n_v : 0;
    const int iT = ilat % p.n_T;
    const int ilon = has ? iT * p.n_v + ci % p.n_v : 0;
    const int nt = p.n_t[iT];
    const double Ti = p.T_samples[iT];
    double cost = CUDART_INF;
    bool feasible = has;
    {
        const double *la = lat + 6 * ilat, *lo = lon + 5 * ilon;
        const double a3 = la[3], a4 = la[4], a5 = la[5], b1 = lo[1], b2 = lo[2], b3 = lo[3], b4 = lo[4];
        double Jp = 0.0, Js = 0.0, sd_last = 0.0, d_last = 0.0;
        for (int it = 0; it < nt; ++it) {
            const double t = it * p.dt, t2 = t * t, t3 = t2 * t;
            const double j5 = 6 * a3 + 24 * a4 * t + 60 * a5 * t2;                       // JTP:439-442
            const double j4 = 6 * b3 + 24 * b4 * t;                                      // JTP:482-485
            Jp += j5 * j5;                                                               // JTP:325-326
            Js += j4 * j4;
            const double s_d = b1 + 2 * b2 * t + 3 * b3 * t2 + 4 * b4 * t3;
            const double s_dd = 2 * b2 + 6 * b3 * t + 12 * b4 * t2;
            if (s_d > p.max_speed || fabs(s_dd) > p.max_accel) feasible = false;        // JTP:385-388
            sd_last = s_d;
            d_last = la[0] + la[1] * t + la[2] * t2 + a3 * t3 + a4 * t2 * t2 + a5 * t2 * t3;
        }
        const double ds = (p.target_speed - sd_last) * (p.target_speed - sd_last);
        const double cd = p.kj * Jp + p.kt * Ti + p.kd * d_last * d_last;               // JTP:331
        const double cv = p.kj * Js + p.kt * Ti + p.kd * ds;                             // JTP:332
        if (has) cost = p.klat * cd + p.klon * cv;                                       // JTP:333
    }

    // ---- (candidate, point) pairs over the lanes: Frenet -> global (JTP:345-356, SPL:47-178) ---------
    const int n_pairs = C * Tm;
    const double s_first = cf[0], s_end = cf[9 * (size_t)(nk - 1)];
    const double inv_h = (double)(nk - 1) / (s_end - s_first);
    for (int pr = lane; pr < n_pairs; pr += 32) {
        const int c2 = pr / Tm, it = pr - c2 * Tm;
        const int l2 = c2 / p.n_v, T2i = l2 % p.n_T;
        double x = CUDART_NAN, y = CUDART_NAN;
        if (it < p.n_t[T2i]) {
            const double *la = lat + 6 * l2, *lo = lon + 5 * (T2i * p.n_v + c2 % p.n_v);
            const double t = it * p.dt, t2 = t * t, t3 = t2 * t, t4 = t2 * t2;
            const double d = la[0] + la[1] * t + la[2] * t2 + la[3] * t3 + la[4] * t4 + la[5] * t4 * t;
            const double s = lo[0] + lo[1] * t + lo[2] * t2 + lo[3] * t3 + lo[4] * t4;
            if (s >= s_first && s < s_end) {   // SPL:54-57 (s == s_end runs past the last segment there)
                int j = min(max((int)((s - s_first) * inv_h), 0), nk - 2);
                while (j > 0 && s < cf[9 * (size_t)j]) --j;                             // bisect_right - 1
                while (j < nk - 2 && s >= cf[9 * (size_t)(j + 1)]) ++j;
                const double *r = cf + 9 * (size_t)j;
                const double u = s - r[0], u2 = u * u;
                const double ix = r[1] + r[2] * u + r[3] * u2 + r[4] * u2 * u;
                const double iy = r[5] + r[6] * u + r[7] * u2 + r[8] * u2 * u;
                const double ddx = r[2] + 2.0 * r[3] * u + 3.0 * r[4] * u2;
                const double ddy = r[6] + 2.0 * r[7] * u + 3.0 * r[8] * u2;
                // cos(yaw + pi/2) = -sin(yaw), sin(yaw + pi/2) = cos(yaw), yaw = atan2(ddy, ddx)  (JTP:352-353)
                const double inv = 1.0 / sqrt(ddx * ddx + ddy * ddy);
                x = ix - d * ddy * inv;
                y = iy + d * ddx * inv;
            }
        }
        X[pr] = x;
        Y[pr] = y;
    }
    __syncwarp();
    // truncation: the first point past the end of the route ends the path (JTP:347-349)
    int nv = 0;
    if (has) {
        while (nv < nt && X[ci * Tm + nv] == X[ci * Tm + nv]) ++nv;
        for (int it = nv; it < Tm; ++it) { X[ci * Tm + it] = CUDART_NAN; Y[ci * Tm + it] = CUDART_NAN; }
    }
    __syncwarp();
    // heading and arc length of every path segment (JTP:358-361)
    for (int pr = lane; pr < n_pairs; pr += 32) {
        const int c2 = pr / Tm, it = pr - c2 * Tm;
        double yw = 0.0, dl = 0.0;
        if (it + 1 < Tm) {
            const double dx = X[pr + 1] - X[pr], dy = Y[pr + 1] - Y[pr];   // NaN past the truncation: never read
            yw = atan2(dy, dx);
            dl = sqrt(dx * dx + dy * dy);
        }
        YW[pr] = yw;
        DL[pr] = dl;
    }
    __syncwarp();
    bool free_of_collision = true;
    if (has) {
        // curvature limit (JTP:363-377, 389-390): the last heading / length are repeated
        for (int i = 0; i + 1 < nv; ++i) {
            const double y0 = YW[ci * Tm + i], y1 = i + 1 < nv - 1 ? YW[ci * Tm + i + 1] : YW[ci * Tm + nv - 2];
            double dl = DL[ci * Tm + i];
            if (dl < 0.00001) dl = 0.1;
            if (fabs((y1 - y0) / dl) > p.max_curvature) feasible = false;
        }
        // PRD:21-60: sampled steps t = 2, 4, ... < min(len(fp.x) - 1, len(path.t) - 1)
        if (nuse > 0 && nt >= 2) {
            const int lim = min(nv - 1, p.n_pred - 1);
            const double r2 = p.robot_radius * p.robot_radius;
            for (int q = 0; q < 2 * nuse; ++q) {
                const double *f = circ + 4 * q;
                for (int t = 2; t < lim; t += 2) {
                    const double px = f[0] + (t * p.dt) * f[2], py = f[1] + (t * p.dt) * f[3];
                    const double ex = px - X[ci * Tm + t], ey = py - Y[ci * Tm + t];
                    if (ex * ex + ey * ey <= r2) free_of_collision = false;
                }
            }
        }
    }

    // ---- JTP:125-130: ascending cost (stable), first feasible and collision-free; 0 = none ----------
    const int best = warp_argmin(cost, (has && feasible && free_of_collision) ? ci : -1);
    const int lowest = warp_argmin(cost, has ? ci : -1);
    const int pick = best >= 0 ? best : lowest;
    if (has) {
        a.out_cost[sc * C + ci] = cost;
        a.out_flags[sc * C + ci] = (unsigned char)((feasible ? 1 : 0) | (free_of_collision ? 2 : 0));
        a.out_n_valid[sc * C + ci] = nv;
    }
    const int pick_nv = __shfl_sync(FULL_MASK, nv, pick);
    const int pick_nt = __shfl_sync(FULL_MASK, nt, pick);
    if (lane == 0) { a.out_index[sc] = best + 1; a.out_traj_len[sc] = pick_nv; }
    if (a.out_paths)
        for (int i = lane; i < n_pairs; i += 32) {
            a.out_paths[(sc * n_pairs + i) * 2] = X[i];
            a.out_paths[(sc * n_pairs + i) * 2 + 1] = Y[i];
        }
    // the selected trajectory: x, y and the desired speed profile (JTP:154-156); index 0 brakes (JTP:149-150)
    {
        const int l2 = pick / p.n_v, T2i = l2 % p.n_T;
        const double *lo = lon + 5 * (T2i * p.n_v + pick % p.n_v);
        for (int it = lane; it < Tm; it += 32) {
            const double t = it * p.dt, t2 = t * t;
            double *o = a.out_traj + (sc * Tm + it) * 3;
            o[0] = X[pick * Tm + it];
            o[1] = Y[pick * Tm + it];
            o[2] = it < pick_nt ? (best >= 0 ? lo[1] + 2 * lo[2] * t + 3 * lo[3] * t2 + 4 * lo[4] * t2 * t : 0.0) : CUDART_NAN;
        }
    }
}

}  // namespace dcarl

using namespace dcarl;

extern "C" void dcarl_default_frenet_params(dcarl_frenet_params *p) {
    if (!p) return;
    memset(p, 0, sizeof(*p));
    // JTP:13-39.  Grids as np.arange produces them for the shipped constants.
    const double d[5] = {-4, -2, 0, 2, 4};
    for (int i = 0; i < 5; ++i) p->d_samples[i] = d[i];
    p->n_d = 5;
    p->T_samples[0] = 4.0; p->n_T = 1; p->n_t[0] = 14;
    p->target_speed = 30.0 / 3.6;
    p->tv_samples[0] = 30.0 / 3.6 - 15 / 3.6; p->tv_samples[1] = p->tv_samples[0] + 15 / 3.6; p->n_v = 2;
    p->n_pred = 15; p->obstacles_considered = 4;
    p->dt = 0.3;
    p->max_speed = 50.0 / 3.6; p->max_accel = 10.0; p->max_curvature = 500.0;
    p->robot_radius = 1.0; p->move_gap = 1.0;
    p->kj = 0.1; p->kt = 0.1; p->kd = 1.0; p->klat = 1.0; p->klon = 1.0;
}

extern "C" int dcarl_spline_build(const double *knots_xy, const int64_t *knot_off, int32_t n_paths, double *coef,
                                  dcarl_stream_t stream) {
    if (n_paths < 0) return bad_arg(3, "n_paths < 0");
    if (n_paths == 0) return 0;
    if (!knots_xy || !knot_off) return bad_arg(1, "knots_xy/knot_off is null");
    if (!coef) return bad_arg(4, "coef is null");
    const int cap = 1792;                                             // knots staged per route: 98 KB of shared memory
    const size_t smem = (size_t)cap * kSplineCols * sizeof(double);
    DCARL_CUDA(cudaFuncSetAttribute(k_spline_build, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    k_spline_build<<<(unsigned)n_paths, 32, smem, (cudaStream_t)stream>>>(
        knots_xy, reinterpret_cast<const long long *>(knot_off), n_paths, coef, cap);
    DCARL_LAUNCH_CHECK("k_spline_build");
    return 0;
}

extern "C" int dcarl_frenet_start(const double *coef, const int64_t *knot_off, int32_t n_paths, const int32_t *path_id,
                                  const double *ego, int64_t n, double *out_start, dcarl_stream_t stream) {
    if (n < 0) return bad_arg(6, "n < 0");
    if (n == 0) return 0;
    if (n_paths < 1 || !coef || !knot_off) return bad_arg(1, "no routes");
    if (!path_id || !ego) return bad_arg(4, "null scenario input");
    if (!out_start) return bad_arg(7, "out_start is null");
    if ((n + kFrWarpsPerCta - 1) / kFrWarpsPerCta > 0x7fffffffll) return bad_arg(6, "n too large for one launch");
    const long long grid = (n + kFrWarpsPerCta - 1) / kFrWarpsPerCta;
    k_frenet_start<<<(unsigned)grid, kFrWarpsPerCta * 32, 0, (cudaStream_t)stream>>>(
        coef, reinterpret_cast<const long long *>(knot_off), n_paths, path_id, ego, n, out_start);
    DCARL_LAUNCH_CHECK("k_frenet_start");
    return 0;
}

extern "C" int dcarl_frenet_plan(const double *coef, const int64_t *knot_off, int32_t n_paths, const int32_t *path_id,
                                 const double *start, const double *ego_xy, const double *obs, const int32_t *n_obs,
                                 int32_t max_obs, int64_t n, const dcarl_frenet_params *p, int32_t *out_index,
                                 double *out_cost, uint8_t *out_flags, int32_t *out_n_valid, double *out_paths,
                                 double *out_traj, int32_t *out_traj_len, dcarl_stream_t stream) {
    if (n < 0) return bad_arg(10, "n < 0");
    if (!p) return bad_arg(11, "params is null");
    if (p->n_d < 1 || p->n_d > DCARL_FRENET_MAX_D) return bad_arg(11, "n_d outside [1, 16]");
    if (p->n_T < 1 || p->n_T > DCARL_FRENET_MAX_T) return bad_arg(11, "n_T outside [1, 4]");
    if (p->n_v < 1 || p->n_v > DCARL_FRENET_MAX_V) return bad_arg(11, "n_v outside [1, 8]");
    const int C = p->n_d * p->n_T * p->n_v;
    if (C > 32) return bad_arg(11, "more than 32 candidates");
    int Tm = 0;
    for (int i = 0; i < p->n_T; ++i) {
        if (p->n_t[i] < 1 || p->n_t[i] > 32) return bad_arg(11, "n_t outside [1, 32]");
        if (p->n_t[i] > Tm) Tm = p->n_t[i];
    }
    if (p->obstacles_considered < 0 || p->obstacles_considered > 16) return bad_arg(11, "obstacles_considered outside [0, 16]");
    if (max_obs < 0 || max_obs > 32) return bad_arg(9, "max_obs outside [0, 32]");
    if (!(p->dt > 0.0)) return bad_arg(11, "dt <= 0");
    if (n == 0) return 0;
    if (n_paths < 1 || !coef || !knot_off) return bad_arg(1, "no routes");
    if (!path_id || !start || !ego_xy || !n_obs) return bad_arg(4, "null scenario input");
    if (max_obs > 0 && !obs) return bad_arg(7, "obs is null");
    if (!out_index || !out_cost || !out_flags || !out_n_valid || !out_traj || !out_traj_len)
        return bad_arg(12, "null output");
    if ((n + kFrWarpsPerCta - 1) / kFrWarpsPerCta > 0x7fffffffll) return bad_arg(10, "n too large for one launch");
    FrenetArgs a;
    a.coef = coef; a.knot_off = reinterpret_cast<const long long *>(knot_off); a.path_id = path_id;
    a.start = start; a.ego_xy = ego_xy; a.obs = obs; a.n_obs = n_obs;
    a.out_index = out_index; a.out_cost = out_cost; a.out_flags = out_flags; a.out_n_valid = out_n_valid;
    a.out_paths = out_paths; a.out_traj = out_traj; a.out_traj_len = out_traj_len;
    a.n = n; a.n_paths = n_paths; a.max_obs = max_obs;
    const size_t wb = (frenet_warp_bytes(p->n_d * p->n_T, p->n_T * p->n_v, C, Tm, 2 * p->obstacles_considered) + 15) & ~(size_t)15;
    const size_t smem = wb * kFrWarpsPerCta;
    if (smem > 48 * 1024)
        DCARL_CUDA(cudaFuncSetAttribute(k_frenet_plan, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    const long long grid = (n + kFrWarpsPerCta - 1) / kFrWarpsPerCta;
    k_frenet_plan<<<(unsigned)grid, kFrWarpsPerCta * 32, smem, (cudaStream_t)stream>>>(a, *p, Tm, (unsigned)wb);
    DCARL_LAUNCH_CHECK("k_frenet_plan");
    return 0;
}
