// Streaming confidence-bound estimator + argmax selector.
//
// Replaces the hot loops of Simulation_testing/Simulation_1/test_DCARL.py:73-99 and
// Simulation_testing/Simulation_2/test_DCARL.py:72-105 for n_inst independent scenario
// instances.  The reference re-materialises the whole sample list of a cell on every
// step (np.array(list), np.mean, np.std); here a cell is three running float64 moments
// (n, sum, sum of squares) and the bound is a closed form of them.
//
// Two kernels:
//   k_stream_rows  — the product path.  One warp owns one instance; the S x A moment
//                    table, the value table and the per-state (max, argmax, bound on the
//                    runner-up) live in shared memory.  32 consecutive rows are processed
//                    at once, one row per lane:
//                      phase 1  cell moments: peers of the same cell inside the batch are
//                               found with match.any and prefix-summed in time order,
//                               then every lane evaluates its bound in float64;
//                      phase 2  argmax: a row cannot move its state's argmax unless it
//                               updates the argmax cell to below the runner-up bound or
//                               lifts another cell above that bound; rows that cannot are
//                               resolved in parallel (the running max is the value of the
//                               latest argmax-cell row), the rare row that might is
//                               replayed exactly with a warp-wide top-2 scan.
//   k_stream_walk  — any shape: tables in a global-memory workspace, rows replayed one
//                    at a time with lanes over actions.  Fallback + cross-check.
#include <math_constants.h>

#include "common.cuh"

namespace dcarl {

struct StreamArgs {
    const int *idx;
    const int *act;
    const void *R;
    void *out_value;
    int *out_act;
    long long *out_activation;
    void *out_overall;
    unsigned long long *act_hist;
    unsigned long long *bad_rows;
    long long n_inst, n_steps;
    int S, A;
    double c_h;  // scale * sqrt(ln(1/alpha) / 2)
    double cap, prior_rule, prior_other, bonus;
    int n_thres, rule_act;
};

__device__ __forceinline__ double u32_to_double(unsigned n) {
    // exact for every uint32; avoids the slow I2F.F64 conversion path
    return __hiloint2double(0x43300000, (int)n) - 4503599627370496.0;
}

// Value of a cell holding n samples with sum s1 and sum of squares s2.
//   rule action:  min(cap, mean + scale*sqrt(L/2/n))                       ST1:12
//   other:        min(mean - scale*sqrt(L/2/n),                            ST1:16
//                     sum/n/(n+1) - 4*std/(n+1) + sum/(n+1) - scale*sqrt(L/2/(n+1)))   ST1:24
// evaluated in float64 with one rsqrt per radicand (a few ulp from the reference's
// divide/sqrt chain, far inside the 1e-5 value tolerance).
__device__ __forceinline__ double cell_value(unsigned n, double s1, double s2, bool is_rule,
                                             double c_h, double cap) {
    const double dn = u32_to_double(n);
    const double rs = rsqrt(dn);
    const double inv_n = rs * rs;
    const double hn = c_h * rs;
    const double mean = s1 * inv_n;
    const double ub = fmin(cap, mean + hn);
    const double rs1 = rsqrt(dn + 1.0);
    const double inv_n1 = rs1 * rs1;
    const double hn1 = c_h * rs1;
    const double var = fma(s2, inv_n, -mean * mean);
    const double sigma = var > 0.0 ? var * rsqrt(var) : 0.0;
    const double lb = mean - hn;
    const double ci = fma(fma(-4.0, sigma, mean) + s1, inv_n1, -hn1);
    return is_rule ? ub : fmin(lb, ci);
}

// Runner-up bound kept per state: any value in [second, max) keeps both fast-path
// tests sound; the midpoint tolerates drift of either cell.
__device__ __forceinline__ double runner_up_bound(double second, double mx) {
    if (second == -CUDART_INF) return second;
    if (!(second < mx)) return mx;
    return 0.5 * second + 0.5 * mx;
}

// Warp-wide exact (max, first argmax, max of the rest) of row[0..A).
__device__ __forceinline__ void warp_top2(const double *row, int A, double &mx, int &arg,
                                          double &second) {
    const int lane = lane_id();
    double bm = -CUDART_INF, b2 = -CUDART_INF;
    int bg = 0x7fffffff;
    for (int a = lane; a < A; a += 32) {
        const double c = row[a];
        if (c > bm) { b2 = bm; bm = c; bg = a; }
        else if (c > b2) b2 = c;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
        const double om = __shfl_xor_sync(FULL_MASK, bm, off);
        const double o2 = __shfl_xor_sync(FULL_MASK, b2, off);
        const int og = __shfl_xor_sync(FULL_MASK, bg, off);
        const bool better = om > bm || (om == bm && og < bg);
        if (better) { b2 = fmax(bm, o2); bm = om; bg = og; }
        else { b2 = fmax(b2, om); }
    }
    mx = bm; arg = bg; second = b2;
}

__host__ __device__ inline size_t rows_warp_bytes(int S, int A) {
    const size_t ncell = (size_t)S * A;
    size_t b = ncell * 3 * sizeof(double)      // s1, s2, val
               + (size_t)S * 2 * sizeof(double) // smax, sub2
               + ncell * sizeof(unsigned)       // cnt
               + (size_t)S * 3 * sizeof(int)    // sarg, steps, actstep
               + (size_t)A * sizeof(unsigned);  // per-warp histogram of chosen actions
    return (b + 15) & ~(size_t)15;
}

template <typename RT, typename OT, bool OVERALL>
__global__ void __launch_bounds__(128) k_stream_rows(const StreamArgs p, const int warps_per_cta,
                                                      const unsigned warp_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id();
    const int warp = threadIdx.x >> 5;
    const long long inst = (long long)blockIdx.x * warps_per_cta + warp;
    if (inst >= p.n_inst) return;  // whole warp leaves; no block-wide barriers below

    const int S = p.S, A = p.A, ncell = S * A;
    double *s1 = reinterpret_cast<double *>(smem_raw + (size_t)warp * warp_bytes);
    double *s2 = s1 + ncell;
    double *val = s2 + ncell;
    double *smax = val + ncell;
    double *sub2 = smax + S;
    unsigned *cnt = reinterpret_cast<unsigned *>(sub2 + S);
    int *sarg = reinterpret_cast<int *>(cnt + ncell);
    int *steps = sarg + S;
    int *actstep = steps + S;
    unsigned *hist = reinterpret_cast<unsigned *>(actstep + S);

    for (int c = lane; c < ncell; c += 32) {
        s1[c] = 0.0; s2[c] = 0.0; cnt[c] = 0u;
        val[c] = (c % A == p.rule_act) ? p.prior_rule : p.prior_other;   // ST1:51-53
    }
    for (int a = lane; a < A; a += 32) hist[a] = 0u;
    __syncwarp();
    {
        double m0, sec0; int g0;
        warp_top2(val, A, m0, g0, sec0);   // every state starts identical
        const double ub0 = runner_up_bound(sec0, m0);
        for (int s = lane; s < S; s += 32) {
            smax[s] = m0; sub2[s] = ub0; sarg[s] = g0; steps[s] = 0; actstep[s] = -1;   // ST1:58
        }
    }
    __syncwarp();

    const unsigned lt = lanemask_lt(), le = lanemask_le(), gt = lanemask_gt();
    const long long row0 = inst * p.n_steps;
    const int *gidx = p.idx + row0;
    const int *gact = p.act + row0;
    const RT *gR = reinterpret_cast<const RT *>(p.R) + row0;
    OT *oval = reinterpret_cast<OT *>(p.out_value) + row0;
    int *oact = p.out_act + row0;
    OT *oall = OVERALL ? reinterpret_cast<OT *>(p.out_overall) + row0 : nullptr;

    double total = 0.0;          // running overall_value (ST2:99-105), warp-uniform
    unsigned long long bad = 0;  // lane-local count of out-of-range rows

    // software prefetch: rows of batch b+1 are requested while batch b is processed
    int n_idx = 0, n_act = 0;
    RT n_R = RT(0);
    if (lane < p.n_steps) { n_idx = ld_stream(gidx + lane); n_act = ld_stream(gact + lane); n_R = ld_stream(gR + lane); }

    for (long long k0 = 0; k0 < p.n_steps; k0 += 32) {
        const long long k = k0 + lane;
        const bool in_stream = k < p.n_steps;
        const int s_in = n_idx, a_in = n_act;
        const double r = (double)n_R;
        {
            const long long kn = k + 32;
            if (kn < p.n_steps) { n_idx = ld_stream(gidx + kn); n_act = ld_stream(gact + kn); n_R = ld_stream(gR + kn); }
        }
        const bool active = in_stream && (unsigned)s_in < (unsigned)S && (unsigned)a_in < (unsigned)A;
        if (in_stream && !active) bad++;
        // inactive lanes get a private key so they match nobody
        const int s = active ? s_in : (0x40000000 | lane);
        const int a = a_in;
        const int key = active ? s_in * A + a_in : (0x40000000 | lane);

        // ---- phase 1: running moments of every row's cell, in time order -------------------
        const unsigned mk = __match_any_sync(FULL_MASK, key);
        const unsigned ms = __match_any_sync(FULL_MASK, s);
        unsigned n0 = 0;
        double acc1 = 0.0, acc2 = 0.0;
        if (active) { n0 = cnt[key]; acc1 = s1[key]; acc2 = s2[key]; }
        unsigned earlier = mk & lt;
        while (__any_sync(FULL_MASK, earlier != 0)) {
            const int j = earlier ? (__ffs(earlier) - 1) : lane;
            const double rj = shfl_f64(r, j);
            if (earlier) { acc1 += rj; acc2 = fma(rj, rj, acc2); earlier &= earlier - 1; }
        }
        acc1 += r;
        acc2 = fma(r, r, acc2);
        const unsigned n = n0 + __popc(mk & lt) + 1u;
        __syncwarp();
        if (active && (mk & gt) == 0) { cnt[key] = n; s1[key] = acc1; s2[key] = acc2; }
        const bool upd = active && n > (unsigned)p.n_thres;                      // ST1:86
        const double v = cell_value(n, acc1, acc2, a == p.rule_act, p.c_h, p.cap);

        // ---- phase 2: per-row (max, argmax) of the row's state ----------------------------
        unsigned pend = __ballot_sync(FULL_MASK, active);
        bool mine = active;
        double m_out = CUDART_NAN, smax0 = 0.0;
        int g_out = -1;
        bool first_round = true;
        while (pend) {
            double ms0 = 0.0, ub = 0.0;
            int g = -1;
            if (mine) { ms0 = smax[s]; g = sarg[s]; ub = sub2[s]; }
            if (first_round) { smax0 = ms0; first_round = false; }
            const bool T = mine && upd && a == g;        // row refreshes the argmax cell
            const unsigned tm = __ballot_sync(FULL_MASK, T);
            const unsigned cand = tm & ms & le;
            const int jl = cand ? (31 - __clz(cand)) : lane;
            const double vj = shfl_f64(v, jl);
            const double m_k = cand ? vj : ms0;
            const bool fail = mine && upd && (T ? !(v > ub) : !(v <= ub && v < m_k));
            const unsigned fm = __ballot_sync(FULL_MASK, fail);
            const int j = fm ? (__ffs(fm) - 1) : 32;
            const unsigned fin = fm ? (pend & ((1u << j) - 1u)) : pend;
            if (mine && ((fin >> lane) & 1u)) {
                g_out = g; m_out = m_k;
                if (upd && (mk & fin & gt) == 0) val[key] = v;
                if ((ms & fin & gt) == 0) smax[s] = m_k;
            }
            if (fm == 0) break;
            __syncwarp();
            // exact replay of row j: apply its update, rescan its state
            const int sj = __shfl_sync(FULL_MASK, s, j);
            if (lane == j) val[key] = v;
            __syncwarp();
            double mx, second; int arg;
            warp_top2(val + sj * A, A, mx, arg, second);
            if (lane == j) {
                g_out = arg; m_out = mx;
                smax[sj] = mx; sarg[sj] = arg; sub2[sj] = runner_up_bound(second, mx);
            }
            __syncwarp();
            pend &= ~(fin | (1u << j));
            mine = mine && lane > j;
        }
        __syncwarp();

        // ---- activation step (ST1:98-99) and overall_value (ST2:99-105) -------------------
        int t_before = 0, ast = -1;
        if (active) { t_before = steps[s]; ast = actstep[s]; }
        const unsigned nonrule = __ballot_sync(FULL_MASK, active && g_out != p.rule_act) & ms;
        const bool act_before = ast != -1;
        const int rank_s = __popc(ms & lt);
        __syncwarp();
        if (active) {
            if (!act_before && nonrule && lane == __ffs(nonrule) - 1) actstep[s] = t_before + rank_s + 1;
            if ((ms & gt) == 0) steps[s] = t_before + __popc(ms);
        }
        if (OVERALL) {
            const bool act_now = act_before || (nonrule & le) != 0;
            const double c_k = (active && act_now) ? m_out + p.bonus : 0.0;
            const unsigned prev = ms & lt;
            const int pl = prev ? (31 - __clz(prev)) : lane;
            const double c_prev_lane = shfl_f64(c_k, pl);
            const double c_prev = prev ? c_prev_lane : (act_before ? smax0 + p.bonus : 0.0);
            double d = active ? c_k - c_prev : 0.0;
#pragma unroll
            for (int off = 1; off < 32; off <<= 1) {
                const double o = __shfl_up_sync(FULL_MASK, d, off);
                if (lane >= off) d += o;
            }
            const double ov = total + d;
            total += shfl_f64(d, 31);
            if (in_stream) st_stream(oall + k, (OT)ov);
        }
        if (in_stream) {
            st_stream(oval + k, (OT)m_out);
            st_stream(oact + k, g_out);
        }
        if (p.act_hist != nullptr) {
            const int hk = active ? g_out : (0x40000000 | lane);
            const unsigned mh = __match_any_sync(FULL_MASK, hk);
            if (active && (mh & lt) == 0) hist[g_out] += __popc(mh);
            __syncwarp();
        }
    }
    __syncwarp();
    for (int s = lane; s < S; s += 32) p.out_activation[inst * S + s] = (long long)actstep[s];
    if (p.act_hist != nullptr)
        for (int a = lane; a < A; a += 32)
            if (hist[a]) atomicAdd(p.act_hist + a, (unsigned long long)hist[a]);
    if (p.bad_rows != nullptr && bad) atomicAdd(p.bad_rows, bad);
}

// ---- general-shape walker: tables in global memory, one row at a time ----------------------
__host__ __device__ inline size_t walk_inst_bytes(int S, int A) {
    const size_t ncell = (size_t)S * A;
    size_t b = ncell * 3 * sizeof(double) + (size_t)S * sizeof(double) + ncell * sizeof(unsigned) +
               (size_t)S * 2 * sizeof(int);
    return (b + 15) & ~(size_t)15;
}

template <typename RT, typename OT, bool OVERALL>
__global__ void __launch_bounds__(128) k_stream_walk(const StreamArgs p, unsigned char *ws,
                                                      const size_t inst_bytes) {
    const int lane = lane_id();
    const long long inst = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (inst >= p.n_inst) return;
    const int S = p.S, A = p.A, ncell = S * A;
    double *s1 = reinterpret_cast<double *>(ws + (size_t)inst * inst_bytes);
    double *s2 = s1 + ncell;
    double *val = s2 + ncell;
    double *smax = val + ncell;
    unsigned *cnt = reinterpret_cast<unsigned *>(smax + S);
    int *steps = reinterpret_cast<int *>(cnt + ncell);
    int *actstep = steps + S;

    for (int c = lane; c < ncell; c += 32) {
        s1[c] = 0.0; s2[c] = 0.0; cnt[c] = 0u;
        val[c] = (c % A == p.rule_act) ? p.prior_rule : p.prior_other;
    }
    __syncwarp();
    {
        double m0, sec0; int g0;
        warp_top2(val, A, m0, g0, sec0);
        for (int s = lane; s < S; s += 32) { smax[s] = m0; steps[s] = 0; actstep[s] = -1; }
    }
    __syncwarp();

    const long long row0 = inst * p.n_steps;
    const RT *gR = reinterpret_cast<const RT *>(p.R) + row0;
    OT *oval = reinterpret_cast<OT *>(p.out_value) + row0;
    OT *oall = OVERALL ? reinterpret_cast<OT *>(p.out_overall) + row0 : nullptr;
    unsigned long long bad = 0;

    for (long long k0 = 0; k0 < p.n_steps; k0 += 32) {
        const long long k = k0 + lane;
        int b_idx = 0, b_act = 0;
        double b_r = 0.0;
        if (k < p.n_steps) { b_idx = p.idx[row0 + k]; b_act = p.act[row0 + k]; b_r = (double)gR[k]; }
        double o_val = CUDART_NAN, o_all = 0.0;
        int o_act = -1;
        const int nrow = (int)min((long long)32, p.n_steps - k0);
        for (int t = 0; t < nrow; ++t) {
            const int s = __shfl_sync(FULL_MASK, b_idx, t);
            const int a = __shfl_sync(FULL_MASK, b_act, t);
            const double r = shfl_f64(b_r, t);
            double mx = CUDART_NAN;
            int arg = -1;
            if ((unsigned)s < (unsigned)S && (unsigned)a < (unsigned)A) {   // warp-uniform
                const int key = s * A + a;
                const unsigned n = cnt[key] + 1u;
                const double a1 = s1[key] + r;
                const double a2 = fma(r, r, s2[key]);
                __syncwarp();
                if (lane == 0) {
                    cnt[key] = n; s1[key] = a1; s2[key] = a2;
                    if (n > (unsigned)p.n_thres) val[key] = cell_value(n, a1, a2, a == p.rule_act, p.c_h, p.cap);
                }
                __syncwarp();
                double second;
                warp_top2(val + s * A, A, mx, arg, second);
                if (lane == 0) {
                    smax[s] = mx;
                    const int st = steps[s] + 1;
                    steps[s] = st;
                    if (actstep[s] == -1 && arg != p.rule_act) actstep[s] = st;
                }
                __syncwarp();
            } else if (lane == 0) {
                bad++;
            }
            double ov = 0.0;
            if (OVERALL) {
                // state order, like the reference's inner loop
                for (int sb = 0; sb < S; sb += 32) {
                    const int si = sb + lane;
                    const double c = (si < S && actstep[si] != -1) ? smax[si] + p.bonus : 0.0;
                    const int lim = min(32, S - sb);
                    for (int q = 0; q < lim; ++q) ov += shfl_f64(c, q);
                }
            }
            if (lane == t) { o_val = mx; o_act = arg; o_all = ov; }
        }
        if (k < p.n_steps) {
            oval[k] = (OT)o_val;
            p.out_act[row0 + k] = o_act;
            if (OVERALL) oall[k] = (OT)o_all;
            if (p.act_hist != nullptr && o_act >= 0) atomicAdd(p.act_hist + o_act, 1ull);
        }
    }
    __syncwarp();
    for (int s = lane; s < S; s += 32) p.out_activation[inst * S + s] = (long long)actstep[s];
    if (p.bad_rows != nullptr && bad) atomicAdd(p.bad_rows, bad);
}

// ---- summary statistics (fixed-order float64 reduction) -------------------------------------
template <typename OT>
__global__ void __launch_bounds__(256) k_stream_stats(const OT *out_value, const OT *out_overall,
                                                       const long long *activation,
                                                       const unsigned long long *act_hist,
                                                       long long n_inst, long long n_steps, int S, int A,
                                                       double *stats) {
    __shared__ double sh[4][256];
    const int t = threadIdx.x;
    double v = 0.0, o = 0.0, na = 0.0, sa = 0.0;
    if (n_steps > 0)
        for (long long i = t; i < n_inst; i += 256) {
            v += (double)out_value[i * n_steps + n_steps - 1];
            if (out_overall) o += (double)out_overall[i * n_steps + n_steps - 1];
        }
    for (long long i = t; i < n_inst * S; i += 256) {
        const long long st = activation[i];
        if (st != -1) { na += 1.0; sa += (double)st; }
    }
    sh[0][t] = v; sh[1][t] = o; sh[2][t] = na; sh[3][t] = sa;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (t < w)
            for (int q = 0; q < 4; ++q) sh[q][t] += sh[q][t + w];
        __syncthreads();
    }
    if (t < 4) stats[t] = sh[t][0];
    if (t == 4) stats[4] = (double)n_inst * (double)n_steps;
    for (int a = t; a < A; a += 256) stats[DCARL_STATS_HEADER + a] = act_hist ? (double)act_hist[a] : 0.0;
}

// ---- host side -------------------------------------------------------------------------------
static const size_t kMaxSmem = 227 * 1024;

template <typename RT, typename OT, bool OV>
static int launch_rows(const StreamArgs &a, cudaStream_t st) {
    const size_t wb = rows_warp_bytes(a.S, a.A);
    int wpc = (int)(kMaxSmem / wb);
    if (wpc < 1) return -1;
    if (wpc > 4) wpc = 4;
    const size_t smem = wb * wpc;
    auto kern = k_stream_rows<RT, OT, OV>;
    DCARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    DCARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    const long long grid = (a.n_inst + wpc - 1) / wpc;
    kern<<<(unsigned)grid, wpc * 32, smem, st>>>(a, wpc, (unsigned)wb);
    DCARL_LAUNCH_CHECK("k_stream_rows");
    return 0;
}

template <typename RT, typename OT, bool OV>
static int launch_walk(const StreamArgs &a, void *ws, cudaStream_t st) {
    const size_t ib = walk_inst_bytes(a.S, a.A);
    const long long grid = (a.n_inst + 3) / 4;
    k_stream_walk<RT, OT, OV><<<(unsigned)grid, 128, 0, st>>>(a, (unsigned char *)ws, ib);
    DCARL_LAUNCH_CHECK("k_stream_walk");
    return 0;
}

template <typename RT, typename OT>
static int dispatch_ov(const StreamArgs &a, bool rows, void *ws, cudaStream_t st) {
    if (a.out_overall)
        return rows ? launch_rows<RT, OT, true>(a, st) : launch_walk<RT, OT, true>(a, ws, st);
    return rows ? launch_rows<RT, OT, false>(a, st) : launch_walk<RT, OT, false>(a, ws, st);
}

}  // namespace dcarl

using namespace dcarl;

extern "C" size_t dcarl_stream_bounds_ws(int64_t n_inst, int64_t n_steps, int32_t S, int32_t A) {
    (void)n_steps;
    if (n_inst <= 0 || S <= 0 || A <= 0) return 0;
    return (size_t)n_inst * walk_inst_bytes(S, A);
}

extern "C" int dcarl_stream_bounds(const int32_t *idx, const int32_t *act, const void *R, int r_is_f64,
                                   int64_t n_inst, int64_t n_steps, int32_t S, int32_t A,
                                   const dcarl_params *p, void *out_value, int out_is_f64,
                                   int32_t *out_act, int64_t *out_activation_step, void *out_overall,
                                   uint64_t *act_hist, uint64_t *bad_rows, void *ws, size_t ws_bytes,
                                   int variant, dcarl_stream_t stream) {
    if (n_inst < 0) return bad_arg(5, "n_inst < 0");
    if (n_steps < 0) return bad_arg(6, "n_steps < 0");
    if (S <= 0) return bad_arg(7, "S <= 0");
    if (A <= 0) return bad_arg(8, "A <= 0");
    if ((long long)S * A > (1ll << 28)) return bad_arg(7, "S*A too large");
    if (!p) return bad_arg(9, "params is null");
    if (!(p->alpha > 0.0 && p->alpha <= 1.0)) return bad_arg(9, "alpha must be in (0, 1]");
    if (p->rule_act < 0 || p->rule_act >= A) return bad_arg(9, "rule_act outside [0, A)");
    if (p->n_thres < 0) return bad_arg(9, "n_thres < 0");
    if (n_steps > 0xfffffff0ll) return bad_arg(6, "n_steps does not fit the 32-bit cell counters");
    if (n_inst == 0) return 0;
    if (!out_activation_step) return bad_arg(13, "out_activation_step is null");
    if (n_steps > 0 && (!idx || !act || !R)) return bad_arg(1, "idx/act/R is null");
    if (n_steps > 0 && (!out_value || !out_act)) return bad_arg(10, "out_value/out_act is null");
    if (variant < 0 || variant > 2) return bad_arg(19, "variant must be 0, 1 or 2");
    if ((n_inst + 3) / 4 > 0x7fffffffll) return bad_arg(5, "n_inst too large for one launch");

    const bool fits = rows_warp_bytes(S, A) <= kMaxSmem;
    bool rows = fits;
    if (variant == 1) rows = false;
    if (variant == 2 && !fits) return bad_arg(19, "the S*A table does not fit shared memory");
    if (!rows) {
        const size_t need = dcarl_stream_bounds_ws(n_inst, n_steps, S, A);
        if (!ws || ws_bytes < need) return bad_arg(17, "workspace too small (see dcarl_stream_bounds_ws)");
    }
    StreamArgs a;
    a.idx = idx; a.act = act; a.R = R;
    a.out_value = out_value; a.out_act = out_act;
    a.out_activation = reinterpret_cast<long long *>(out_activation_step);
    a.out_overall = out_overall;
    a.act_hist = reinterpret_cast<unsigned long long *>(act_hist);
    a.bad_rows = reinterpret_cast<unsigned long long *>(bad_rows);
    a.n_inst = n_inst; a.n_steps = n_steps; a.S = S; a.A = A;
    a.c_h = p->scale * sqrt(log(1.0 / p->alpha) / 2.0);
    a.cap = p->cap; a.prior_rule = p->prior_rule; a.prior_other = p->prior_other;
    a.bonus = p->activation_bonus; a.n_thres = p->n_thres; a.rule_act = p->rule_act;
    cudaStream_t st = (cudaStream_t)stream;
    if (r_is_f64)
        return out_is_f64 ? dispatch_ov<double, double>(a, rows, ws, st) : dispatch_ov<double, float>(a, rows, ws, st);
    return out_is_f64 ? dispatch_ov<float, double>(a, rows, ws, st) : dispatch_ov<float, float>(a, rows, ws, st);
}

extern "C" int dcarl_stream_stats(const void *out_value, int out_is_f64, const void *out_overall,
                                  const int64_t *out_activation_step, const uint64_t *act_hist,
                                  int64_t n_inst, int64_t n_steps, int32_t S, int32_t A, double *stats,
                                  dcarl_stream_t stream) {
    if (n_inst < 0 || n_steps < 0) return bad_arg(6, "negative size");
    if (S <= 0 || A <= 0) return bad_arg(8, "S/A <= 0");
    if (!stats) return bad_arg(10, "stats is null");
    if (n_inst > 0 && !out_activation_step) return bad_arg(4, "out_activation_step is null");
    if (n_inst > 0 && n_steps > 0 && !out_value) return bad_arg(1, "out_value is null");
    cudaStream_t st = (cudaStream_t)stream;
    const long long *as = reinterpret_cast<const long long *>(out_activation_step);
    const unsigned long long *ah = reinterpret_cast<const unsigned long long *>(act_hist);
    if (out_is_f64)
        k_stream_stats<double><<<1, 256, 0, st>>>((const double *)out_value, (const double *)out_overall, as, ah,
                                                  n_inst, n_steps, S, A, stats);
    else
        k_stream_stats<float><<<1, 256, 0, st>>>((const float *)out_value, (const float *)out_overall, as, ah,
                                                 n_inst, n_steps, S, A, stats);
    DCARL_LAUNCH_CHECK("k_stream_stats");
    return 0;
}
