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
//                    at once, one row per lane, in two stages skewed by one batch:
//                      moments   peers of the same cell inside the batch are found with
//                                match.any and prefix-summed in time order, then every lane
//                                evaluates its bound in float64;
//                      selector  a row cannot move its state's argmax unless it updates the
//                                argmax cell to below the runner-up bound or lifts another
//                                cell above that bound; rows that cannot are resolved in
//                                parallel (the running max is the value of the latest
//                                argmax-cell row), the rare row that might is replayed
//                                exactly with a warp-wide top-2 scan.
//   k_stream_walk  — any shape: tables in a global-memory workspace, rows replayed one
//                    at a time with lanes over actions.  Fallback + cross-check.
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"
#include "peer.cuh"

namespace dcarl {

struct StreamArgs {
    const void *idx;   // int32, or uint8 in the compact wire format
    const void *act;
    const void *R;
    void *out_value;
    void *out_act;
    long long *out_activation;
    void *out_overall;
    unsigned long long *act_hist;
    unsigned long long *bad_rows;
    long long n_inst;
    unsigned n_steps;
    int S, A;
    double c_h;  // scale * sqrt(ln(1/alpha) / 2)
    double cap, prior_rule, prior_other, bonus;
    int n_thres, rule_act;
};


__device__ __forceinline__ double u32_to_double(unsigned n) {
    // exact for every uint32; avoids the slow I2F.F64 conversion path
    return __hiloint2double(0x43300000, (int)n) - 4503599627370496.0;
}

// 1/sqrt(x) for positive, normal, finite x: hardware seed (MUFU.RSQ64H, ~2^-20) and one
// third-order step y(1 + e/2 + 3e^2/8), e = 1 - x y^2; remaining error ~e^3 << 2^-53.
#ifdef DCARL_V_CONTRACT   // A/B variant: let the compiler contract mul+add pairs (fewer instructions, path-dependent bits)
#define DMUL(a, b) ((a) * (b))
#define DADD(a, b) ((a) + (b))
#define DFMA(a, b, c) fma(a, b, c)
#else
#define DMUL(a, b) __dmul_rn(a, b)
#define DADD(a, b) __dadd_rn(a, b)
#define DFMA(a, b, c) __fma_rn(a, b, c)
#endif
__device__ __forceinline__ double rsqrt_pos(double x) {
    double y;
    asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
    const double t = DMUL(x, y);
    const double e = DFMA(-t, y, 1.0);
    const double q = DFMA(e, 0.375, 0.5);
    return DFMA(DMUL(y, e), q, y);
}

__device__ __forceinline__ double dmin(double a, double b) { return b < a ? b : a; }

// Index of the highest set bit (0xffffffff for 0): one FLO, where 31 - __clz(x) compiles to three instructions.
__device__ __forceinline__ unsigned top_bit(unsigned x) {
    unsigned r;
    asm("bfind.u32 %0, %1;" : "=r"(r) : "r"(x));
    return r;
}

// Everything the bound needs from the sample count alone: 1/n and c_h/sqrt(n), c_h = scale*sqrt(ln(1/alpha)/2).
// The row kernel keeps a table of these per CTA (n below the table size: two 16-byte loads replace two
// 1/sqrt evaluations); every other path builds them on the fly with the same roundings.
struct __align__(16) NEntry {
    double inv_n, hn;
};
__device__ __forceinline__ NEntry make_entry(double dn, double c_h) {
    const double rs = rsqrt_pos(dn);
    NEntry e;
    e.inv_n = DMUL(rs, rs);
    e.hn = DMUL(c_h, rs);
    return e;
}

// Value of a cell holding n samples with sum s1 and sum of squares s2; e0 / e1 = entries of n / n+1.
//   rule action:  min(cap, mean + scale*sqrt(L/2/n))                       ST1:12
//   other:        min(mean - scale*sqrt(L/2/n),                            ST1:16
//                     sum/n/(n+1) - 4*std/(n+1) + sum/(n+1) - scale*sqrt(L/2/(n+1)))   ST1:24
// evaluated in float64 with explicit roundings (identical bits on every path; a few ulp from the
// reference's divide/sqrt chain, far inside the 1e-5 value tolerance; the argmax only flips on ties
// closer than ~1e-11).  The variance comes from one-pass moments, s2/n - mean^2, whose rounding error is
// bounded by (3n + 4) 2^-53 mean^2: anything at or below 8n 2^-53 mean^2 is indistinguishable from a
// constant cell and gives sigma = 0 exactly, like the reference's two-pass np.std on constant samples.
__device__ __forceinline__ double cell_value_e(const NEntry e0, const NEntry e1, double dn, double s1, double s2,
                                               bool is_rule, double cap) {
    const double mean = DMUL(s1, e0.inv_n);
    // +hn for the rule action (upper bound), -hn otherwise (lower bound)
    const double hs = __hiloint2double(__double2hiint(e0.hn) ^ (is_rule ? 0 : (int)0x80000000), __double2loint(e0.hn));
    const double x = DADD(mean, hs);
    const double t = DMUL(mean, mean);
    const double var = DFMA(s2, e0.inv_n, -t);
    // var <= t * n * 2^-50  <=>  var - (t * 2^-50) * n <= 0.  The 1/sqrt chain does not wait for the
    // test: it runs on whatever var holds (garbage for var <= 0) and the select discards it.
    const double sd = DMUL(var, rsqrt_pos(var));
#ifdef DCARL_V_ABS_CLAMP
    const double sigma = var > 1e-280 ? sd : 0.0;
#else
    const double sigma = DFMA(-DMUL(t, 8.8817841970012523e-16 /* 2^-50 */), dn, var) > 0.0 ? sd : 0.0;
#endif
    const double ci = DFMA(DADD(DFMA(-4.0, sigma, mean), s1), e1.inv_n, -e1.hn);
    return dmin(x, is_rule ? cap : ci);
}

__device__ __forceinline__ double cell_value(unsigned n, double s1, double s2, bool is_rule, double c_h,
                                             double cap) {
    const double dn = u32_to_double(n);
    return cell_value_e(make_entry(dn, c_h), make_entry(DADD(dn, 1.0), c_h), dn, s1, s2, is_rule, cap);
}

// Runner-up bound kept per state: any value in [second, max) keeps both fast-path tests
// sound.  The midpoint measured best on config 2 (1.21 ms vs 1.23 at 1/4, 1.35 at 1/20 of the gap).
__device__ __forceinline__ double runner_up_bound(double second, double mx) {
    if (second == -CUDART_INF) return second;
    if (!(second < mx)) return mx;
    const double b = fma(0.5, mx - second, second);
    return b < mx ? b : second;
}

// ---- exact warp-wide max of doubles with integer reductions (REDUX) ---------------------------
// Order-preserving 64-bit key of a (non-NaN) double, split in two 32-bit halves.
__device__ __forceinline__ void order_key(double x, unsigned &hi, unsigned &lo) {
    x += 0.0;  // -0.0 -> +0.0 so that equal doubles get equal keys
    const int h = __double2hiint(x);
    const unsigned sgn = (unsigned)(h >> 31);
    hi = (unsigned)h ^ (sgn | 0x80000000u);
    lo = (unsigned)__double2loint(x) ^ sgn;
}
// Mask of the lanes holding the warp maximum of x.
__device__ __forceinline__ unsigned warp_max_lanes(double x) {
    unsigned hi, lo;
    order_key(x, hi, lo);
    const unsigned H = __reduce_max_sync(FULL_MASK, hi);
    const bool top = hi == H;
    const unsigned L = __reduce_max_sync(FULL_MASK, top ? lo : 0u);
    return __ballot_sync(FULL_MASK, top && lo == L);
}

// Exact (max, first argmax, max of the rest) of row[0 .. A) with element stride `stride` doubles.
__device__ __forceinline__ void warp_top2(const double *row, int stride, int A, double &mx, int &arg,
                                          double &second) {
    const int lane = lane_id();
    double bm = -CUDART_INF, b2 = -CUDART_INF;
    int bg = 0x7fffffff;
    for (int a = lane; a < A; a += 32) {   // ascending a: a strict compare keeps the first index
        const double c = row[(size_t)a * stride];
        if (c > bm) { b2 = bm; bm = c; bg = a; }
        else if (c > b2) b2 = c;
    }
    const unsigned win = warp_max_lanes(bm);
    arg = __reduce_min_sync(FULL_MASK, ((win >> lane) & 1u) ? bg : 0x7fffffff);
    const int wl = __ffs(__ballot_sync(FULL_MASK, ((win >> lane) & 1u) && bg == arg)) - 1;
    mx = shfl_f64(bm, wl);
    const double contrib = lane == wl ? b2 : bm;
    const unsigned win2 = warp_max_lanes(contrib);
    second = shfl_f64(contrib, __ffs(win2) - 1);
}

// Inverse of order_key.
__device__ __forceinline__ double key_to_double(unsigned hi, unsigned lo) {
    const unsigned sgn = ~(unsigned)((int)hi >> 31);   // key top bit clear <=> the double was negative
    return __hiloint2double((int)(hi ^ (sgn | 0x80000000u)), (int)(lo ^ sgn));
}

// A <= 32: one action per lane.  Exact (max, first argmax, max of the rest) with four integer
// reductions; lanes past A hold -inf, so a single action gives second = -inf.
__device__ __forceinline__ void warp_top2_small(const double *row, int stride, int A, double &mx, int &arg,
                                                double &second) {
    const int lane = lane_id();
    const double c = lane < A ? row[(size_t)lane * stride] : -CUDART_INF;
    unsigned hi, lo;
    order_key(c, hi, lo);
    const unsigned H = __reduce_max_sync(FULL_MASK, hi);
    const bool top = hi == H;
    const unsigned L = __reduce_max_sync(FULL_MASK, top ? lo : 0u);
    arg = __ffs(__ballot_sync(FULL_MASK, top && lo == L)) - 1;
    mx = key_to_double(H, L);
    const bool rest = lane != arg;
    const unsigned H2 = __reduce_max_sync(FULL_MASK, rest ? hi : 0u);
    const unsigned L2 = __reduce_max_sync(FULL_MASK, (rest && hi == H2) ? lo : 0u);
    second = key_to_double(H2, L2);
}

// One step of an inclusive warp scan: x += (value `off` lanes below), using the in-range predicate
// the shuffle itself returns (no lane compare).
__device__ __forceinline__ float scan_step(float x, int off) {
    float r;
    asm volatile(
        "{\n"
        ".reg .f32 t;\n"
        ".reg .pred p;\n"
        "shfl.sync.up.b32 t|p, %1, %2, 0, 0xffffffff;\n"
        "mov.f32 %0, %1;\n"
        "@p add.f32 %0, %1, t;\n"
        "}\n"
        : "=f"(r) : "f"(x), "r"(off));
    return r;
}
__device__ __forceinline__ double scan_step(double x, int off) {
    double r;
    asm volatile(
        "{\n"
        ".reg .b32 lo, hi, tl, th;\n"
        ".reg .f64 t;\n"
        ".reg .pred p;\n"
        "mov.b64 {lo, hi}, %1;\n"
        "shfl.sync.up.b32 tl|p, lo, %2, 0, 0xffffffff;\n"
        "shfl.sync.up.b32 th|p, hi, %2, 0, 0xffffffff;\n"
        "mov.b64 t, {tl, th};\n"
        "mov.f64 %0, %1;\n"
        "@p add.f64 %0, %1, t;\n"
        "}\n"
        : "=d"(r) : "d"(x), "r"(off));
    return r;
}

// Per-instance tables of the row kernel (shared memory).  The kernel is bound by issue slots first (time
// follows the instruction count, profiles/r02_stream_rows_*), by the shared-memory data pipe second: a
// cell's {s1, s2, val} share one address computation, the counts sit in their own array (a 4-byte field
// inside a wide record costs far more bank conflicts than it saves instructions: measured).
struct Cell {
    double s1, s2;   // running sum and sum of squares of the cell's returns   (moment stage)
    double val;      // the cell's current bound value                          (selector stage)
};
// Per-state record: one 16-byte load per row.  The low 16 mantissa bits of `sub2` carry the state's
// argmax (15 bits) and its "activated" flag (bit 15), see pack_bound().
struct __align__(16) StateHot {
    double smax;   // current max_a value[s][a]
    double sub2;   // runner-up bound: second-best <= sub2 (and < smax unless the top two are closer than ~1e-10)
};

constexpr unsigned kArgMask = 0x7fffu, kActivatedBit = 0x8000u;

// A bound >= b whose low 16 mantissa bits are `payload`.  Any value >= the true second-best keeps
// the selector's fast path sound (a looser bound only sends more rows to the exact path), so the
// bound is first moved up by 2^18 ulp, then its low 16 bits are replaced (a change below 2^16 ulp).
__device__ __forceinline__ double pack_bound(double b, unsigned payload) {
    b = b > -1e300 ? b : -1e300;                                   // -inf (a single action) -> finite
    b = __dadd_rn(__fma_rn(fabs(b), 5.8207660913467407e-11 /* 2^-34 */, b), 1e-300);   // (+1e-300: zeros and denormals)
    return __hiloint2double(__double2hiint(b), (int)(((unsigned)__double2loint(b) & 0xffff0000u) | payload));
}
// Same for a bound known to be finite (every path but the initial one, where a single action gives -inf).
__device__ __forceinline__ double pack_bound_finite(double b, unsigned payload) {
    b = __fma_rn(fabs(b), 5.8207660913467407e-11 /* 2^-34 */, b);   // b == 0 stays 0: the payload alone is a tiny positive
    return __hiloint2double(__double2hiint(b), (int)(((unsigned)__double2loint(b) & 0xffff0000u) | payload));
}
__device__ __forceinline__ unsigned bound_payload(double sub2) { return (unsigned)__double2loint(sub2) & 0xffffu; }

__host__ __device__ inline size_t rows_warp_bytes(int S, int A) {
    const size_t ncell = (size_t)S * A;
    size_t b = (size_t)S * (sizeof(StateHot) + 2 * sizeof(int))                    // hot, actstep, flushpt
               + ncell * (sizeof(Cell) + sizeof(unsigned))                         // cells, cnt
               + (size_t)A * sizeof(unsigned);                                     // per-warp action histogram
    return (b + 15) & ~(size_t)15;
}
// (Tried: a lockstep of the warps of a CTA — every 8 batches a warp that is ahead of the slowest warp of its
// CTA sleeps, so that all instances of an SM finish together.  0.849 -> 0.866-0.883 ms for every period / lag /
// sleep tried: the spread of the finish times is not where the time goes.)
constexpr size_t kRowsCtaFixedBytes = 0;
constexpr int kRowsMaxWarps = 28;   // 72 registers per thread: 28 resident warps per SM, 7 per scheduler

// ---- row kernel: one warp per instance, 32 rows per batch, software-pipelined ----------------
// The loop is skewed by one batch.  The moment stage of batch b+1 (row loads, peer matching,
// running moments, bound) touches only {cnt, s1, s2}; the selector stage of batch b (argmax
// bookkeeping, overall_value scan, output stores) touches only {val, state records}.  Placing
// the two in the same straight-line region gives every warp two independent dependency chains:
// with 7 warps per scheduler (config 2 is one resident wave) the kernel is bound by the length
// of a warp's own chain, not by issue slots.  The first selector round is peeled out of its retry
// loop so that it shares a basic block with the next batch's moment loads; the match.any pair is
// issued before that round and consumed after it.  The warp index is broadcast from lane 0 so
// that the compiler knows every loop bound is warp-uniform (no divergence fallbacks around the
// warp-synchronous intrinsics).  float32 outputs scan the in-batch overall_value deltas in
// float32 (the running total stays float64; activation jumps, rare, are added in float64).
// Lanes of the same state (ms) and of the same action (ma); their AND is the cell mask.
// match.any takes about one round per distinct value in the warp, so two matches (~13 + ~11
// distinct values) beat one match on the cell key (~28); one ballot per key bit was measured
// slower still (more instructions: the kernel is within 20 % of issue-bound).
__device__ __forceinline__ void peer_masks(int s, int a, unsigned &ms, unsigned &ma) {
    ms = __match_any_sync(FULL_MASK, s);
    ma = __match_any_sync(FULL_MASK, a);
}

#ifdef DCARL_DEBUG_TIMING
__device__ unsigned long long g_rows_timing[4 * 65536];   // per instance: start ns, end ns, smid, hw warp id
#endif

template <typename RT, typename OT, typename KT, bool OVERALL, bool FORCE_FIRST>
__global__ void __launch_bounds__(kRowsMaxWarps * 32, 1)
k_stream_rows(const StreamArgs p, const int warps_per_cta, const unsigned warp_bytes) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int lane = lane_id();
    // broadcast from lane 0: tells the compiler the warp index (hence `inst` and every loop bound)
    // is warp-uniform, so the warp-synchronous intrinsics below need no divergence fallbacks
    const int warp = __shfl_sync(FULL_MASK, (int)(threadIdx.x >> 5), 0);

    // One instance per warp.  (A persistent loop over instances was tried: wrapping the batch loop in an
    // outer loop makes ptxas give up its proof that the warp is converged at the warp-synchronous
    // intrinsics, and every one of them gets a divergence fallback again — 64 BRA.DIV instead of 0.)
    const long long inst = (long long)blockIdx.x * warps_per_cta + warp;
    if (inst >= p.n_inst) return;

    const int S = p.S, A = p.A, ncell = S * A;
    StateHot *hot = reinterpret_cast<StateHot *>(smem_raw + kRowsCtaFixedBytes + (size_t)warp * warp_bytes);
    Cell *cells = reinterpret_cast<Cell *>(hot + S);
    unsigned *cnt = reinterpret_cast<unsigned *>(cells + ncell);
    int *actstep = reinterpret_cast<int *>(cnt + ncell);       // activation step per state, -1 = not yet (ST1:58)
    int *flushpt = actstep + S;
    unsigned *hist = reinterpret_cast<unsigned *>(flushpt + S);

    for (int c = lane; c < ncell; c += 32) {
        cells[c].s1 = 0.0; cells[c].s2 = 0.0; cnt[c] = 0u;
        cells[c].val = (c % A == p.rule_act) ? p.prior_rule : p.prior_other;   // ST1:51-53
    }
    for (int a = lane; a < A; a += 32) hist[a] = 0u;
    __syncwarp();
    {
        double m0, sec0; int g0;
        warp_top2(&cells[0].val, 3, A, m0, g0, sec0);
        StateHot h0;
        h0.smax = m0; h0.sub2 = pack_bound(runner_up_bound(sec0, m0), (unsigned)g0);
        for (int s = lane; s < S; s += 32) { hot[s] = h0; actstep[s] = -1; flushpt[s] = 0; }
    }
    __syncwarp();

#ifdef DCARL_DEBUG_TIMING
    if (lane == 0 && inst < 65536) {
        unsigned long long t; unsigned smid, wid;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
        asm volatile("mov.u32 %0, %%warpid;" : "=r"(wid));
        g_rows_timing[inst * 4 + 0] = t; g_rows_timing[inst * 4 + 2] = smid; g_rows_timing[inst * 4 + 3] = wid;
    }
#endif
    // the per-state records are addressed by 32-bit shared-space address in the loop: the generic
    // pointer costs four instructions per batch to rebuild its shared window base
    const unsigned hot_sa = (unsigned)__cvta_generic_to_shared(hot);
    const unsigned bit = 1u << lane;
    const unsigned lt = bit - 1u, le = lt | bit, gt = ~le;
    auto gt_of = [](int l) { return ~((2u << l) - 1u); };   // lanes strictly above lane l
    const unsigned n_steps = p.n_steps;
    const double c_h = p.c_h, cap = p.cap, bonus = p.bonus;
    const int rule = p.rule_act;
    const unsigned n_thres = (unsigned)p.n_thres;
    constexpr bool SCAN_F32 = std::is_same<OT, float>::value;

    const KT *__restrict__ gidx = reinterpret_cast<const KT *>(p.idx);
    const KT *__restrict__ gact = reinterpret_cast<const KT *>(p.act);
    const RT *__restrict__ gR = reinterpret_cast<const RT *>(p.R);
    OT *__restrict__ oval = reinterpret_cast<OT *>(p.out_value);
    KT *__restrict__ oact = reinterpret_cast<KT *>(p.out_act);
    OT *__restrict__ oall = reinterpret_cast<OT *>(p.out_overall);
    const unsigned ebase = (unsigned)inst * n_steps + lane;

    double total = 0.0;   // running overall_value (ST2:99-105), warp-uniform
    unsigned clamped = 0;   // lane-local OR of (raw key ^ clamped key): non-zero once a row was out of range

    int n_idx = 0, n_act = 0;
    RT n_R = RT(0);
    if ((unsigned)lane < n_steps) { n_idx = (int)ld_stream(gidx + ebase); n_act = (int)ld_stream(gact + ebase); n_R = ld_stream(gR + ebase); }

    struct Rows {      // a batch after its moment stage
        int s, sx, a, kx;
        unsigned ms, mk, n;
        double v;
        bool active;
    };
    struct Mom {       // moment stage in flight
        double acc1, acc2;
        unsigned earlier;   // earlier lanes of the same cell, oldest first
        RT rin, r1;         // own return and that of the oldest earlier peer
    };
    struct Sel {       // selector stage results
        double m_out, mprev_out;   // the state's max just after / just before the row
        int g_out;                 // the state's argmax after the row
        bool act_prev;             // the row's state was activated before the row
        bool jump;                 // the row activates its state (ST1:98-99)
        bool any_jump;             // warp-uniform: some row of the batch does
    };

    // ---- moment stage, issue part: consume the prefetched rows, request the batch after, start
    // the two peer matches and the cell loads.  Nothing here waits on a result: match.any takes
    // ~100 cycles per call, so its consumers sit in p1a_peers, after the selector stage of the
    // previous batch.
    auto p1a_issue = [&](auto tail_tag, const unsigned k, Rows &r, Mom &m) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const bool in_stream = TAIL ? (k < n_steps) : true;
        // out-of-range keys are clamped (memory safety) and remembered: such rows are counted at the end
        // (lanes past the end of the stream re-read keys that were already checked)
        const int s_in = (int)min((unsigned)n_idx, (unsigned)(S - 1));
        const int a = (int)min((unsigned)n_act, (unsigned)(A - 1));
        clamped |= (unsigned)(n_idx ^ s_in) | (unsigned)(n_act ^ a);
        m.rin = n_R;
        const unsigned e = ebase + (k - lane);
        if (k + 32 < n_steps) { n_idx = (int)ld_stream(gidx + e + 32); n_act = (int)ld_stream(gact + e + 32); n_R = ld_stream(gR + e + 32); }
        r.active = in_stream;
        r.s = in_stream ? s_in : (0x40000000 | lane);
        r.sx = in_stream ? s_in : 0;
        r.a = a;
        r.kx = in_stream ? s_in * A + a : 0;
        peer_masks(r.s, a, r.ms, r.mk);   // mk = action peers until p1a_peers
        const Cell *cell = cells + r.kx;
        m.acc1 = cell->s1; m.acc2 = cell->s2;
        r.n = cnt[r.kx];
    };
    // peers part: the cell mask, the row's sample count, the return of the oldest in-batch peer
    auto p1a_peers = [&](Rows &r, Mom &m) {
        r.mk &= r.ms;
        m.earlier = r.mk & lt;
        r.n += __popc(m.earlier) + 1u;
        m.r1 = __shfl_sync(FULL_MASK, m.rin, __ffs(m.earlier) - 1);   // (no peer: lane -1 & 31, the value is not used)
    };
    // the oldest peer (11 % of the rows have one), then a loop for the rows with more (0.7 %)
    auto p1_peers = [&](Mom &m) {
        if (m.earlier) {
            const double x = (double)m.r1; m.acc1 += x; m.acc2 = fma(x, x, m.acc2);
            m.earlier &= m.earlier - 1;
        }
        while (__any_sync(FULL_MASK, m.earlier != 0)) {
            const RT rj = __shfl_sync(FULL_MASK, m.rin, m.earlier ? (__ffs(m.earlier) - 1) : lane);
            if (m.earlier) { const double x = (double)rj; m.acc1 += x; m.acc2 = fma(x, x, m.acc2); m.earlier &= m.earlier - 1; }
        }
    };
    // part b: own return, write-back by the last lane of each cell group, the bound in float64
    auto p1b = [&](Rows &r, Mom &m) {
        const double x = (double)m.rin;
        m.acc1 += x;
        m.acc2 = fma(x, x, m.acc2);
        __syncwarp();   // every lane has read its cell
        if (r.active && (r.mk & gt) == 0) {
            Cell *cell = cells + r.kx;
            cnt[r.kx] = r.n; cell->s1 = m.acc1; cell->s2 = m.acc2;
        }
        r.v = cell_value(r.n, m.acc1, m.acc2, r.a == rule, c_h, cap);
    };

    // ---- selector stage: per-row (max, argmax) of the row's state
    auto p2 = [&](const Rows &r, Sel &o) {
        const int lane_ = lane;
        const unsigned sh_a = hot_sa + (unsigned)r.sx * (unsigned)sizeof(StateHot);
        StateHot h;
        lds_v2f64(sh_a, h.smax, h.sub2);
        const bool upd = r.active && r.n > n_thres;                               // ST1:86
        const double v = r.v;
        const int a = r.a;
        const unsigned ms = r.ms, mk = r.mk;
        o.jump = false; o.any_jump = false;
        unsigned pend = __ballot_sync(FULL_MASK, r.active);
        bool mine = r.active;
        unsigned fm, fin;
        double m_prev;
        auto round = [&](auto first_tag) {
            constexpr bool FIRST = decltype(first_tag)::value;   // first round of the batch: every active row is pending
            const double c_ms0 = h.smax, c_ub = h.sub2;
            const unsigned pl = bound_payload(c_ub);
            const int c_g = (int)(pl & kArgMask);
            const bool activated = (pl & kActivatedBit) != 0;
            const bool T = mine && upd && a == c_g;
            const unsigned tm = __ballot_sync(FULL_MASK, T);
            const unsigned cand = tm & ms & lt;
            // (no candidate: the source lane is 0xffffffff & 31 = 31, a valid lane whose value the select below drops)
            const double vp = shfl_f64(v, (int)top_bit(cand));
            m_prev = cand ? vp : c_ms0;
            const double m_k = T ? v : m_prev;
            // (bitwise on purpose: the compares are evaluated unconditionally, no short-circuit branches)
            const bool gt_ub = v > c_ub, lt_mk = v < m_k;
            // A row of another cell that rises above the runner-up bound but stays below the max keeps
            // the argmax; it only lifts the bound that later argmax-cell rows of its state must clear.
            const bool forced = FORCE_FIRST ? (mine & !activated) : false;
            const bool lifter = mine & upd & !T & gt_ub & lt_mk & !forced;
            const unsigned lm = __ballot_sync(FULL_MASK, lifter);
            double B = c_ub;   // bound in force at this row: c_ub and the earlier lifters of its state
            if (lm != 0) {     // rare
                for (unsigned w = lm; w; w &= w - 1) {
                    const int l = __ffs(w) - 1;
                    const double vl = shfl_f64(v, l);
                    const unsigned msl = __shfl_sync(FULL_MASK, ms, l);
                    if ((msl & bit & gt_of(l)) && vl > B) B = vl;
                }
            }
            // doubtful: the argmax cell drops to the bound, or another cell reaches the max.
            // FORCE_FIRST (the priors' argmax is not the rule action): every row of a state that is
            // not activated yet goes through the exact path, which is where activation is recorded.
            const bool doubt = mine & upd & (T ? !(v > B) : !lt_mk);
            fm = __ballot_sync(FULL_MASK, doubt | forced);
            fin = fm ? (pend & ((fm & (0u - fm)) - 1u)) : pend;
            __syncwarp();
            const bool done = mine & ((fin & bit) != 0);
            if (FIRST) {   // rows that are not done yet are overwritten by a later round or by their replay
                o.g_out = c_g; o.m_out = m_k; o.mprev_out = m_prev; o.act_prev = activated;
            } else {
                o.g_out = done ? c_g : o.g_out;
                o.m_out = done ? m_k : o.m_out;
                o.mprev_out = done ? m_prev : o.mprev_out;
                o.act_prev = done ? activated : o.act_prev;
            }
            if (done & upd & ((mk & fin & gt) == 0)) cells[r.kx].val = v;
            const bool last_of_state = done & ((ms & fin & gt) == 0);
            if (last_of_state) sts_f64(sh_a, m_k);
            if (lm != 0) {     // rare: the last finalized row of a lifted state re-centres its bound
                if (last_of_state & ((lm & ms & fin) != 0))
                    sts_f64(sh_a + 8u, pack_bound_finite(runner_up_bound((lifter & (v > B)) ? v : B, m_k), pl));
            }
        };
        round(std::true_type{});
        if (fm != 0) {
            for (;;) {
                // ---- exact replay of the first doubtful row j
                const int j = __ffs(fm) - 1;
                __syncwarp();
                const int sj = __shfl_sync(FULL_MASK, r.s, j);
                const unsigned msj = __shfl_sync(FULL_MASK, ms, j);
                if (lane_ == j && upd) cells[r.kx].val = v;
                __syncwarp();
                const unsigned opl = bound_payload(lds_f64(hot_sa + (unsigned)sj * (unsigned)sizeof(StateHot) + 8u));
                const int o_arg = (int)(opl & kArgMask);
                const bool o_activated = (opl & kActivatedBit) != 0;
                double mx, second; int arg;
                if (A <= 32) warp_top2_small(&cells[sj * A].val, 3, A, mx, arg, second);
                else warp_top2(&cells[sj * A].val, 3, A, mx, arg, second);
                __syncwarp();   // every lane has read the state record and the values before lane j rewrites them
                // The argmax only ever changes here, so a state can only activate here (ST1:98-99):
                // between two exact rows "argmax != rule" implies "already activated".
                const bool activates = !o_activated && arg != rule;
                if (arg != o_arg || activates) {
                    unsigned csum = 0;
                    for (int q = lane_; q < A; q += 32) csum += cnt[sj * A + q];
                    // 1-based count of this state's rows up to and including row j
                    const int t_j = (int)__reduce_add_sync(FULL_MASK, csum) - __popc(msj & ~((2u << j) - 1u));
                    if (lane_ == j) {
                        if (arg != o_arg) {
                            // rows (flushpt, t_j) of this state were emitted with the old argmax: account for them
                            hist[o_arg] += (unsigned)(t_j - 1 - flushpt[sj]); flushpt[sj] = t_j - 1;
                        }
                        if (activates) { actstep[sj] = t_j; o.jump = true; }
                    }
                    o.any_jump |= activates;
                }
                if (lane_ == j) {
                    o.g_out = arg; o.m_out = mx; o.mprev_out = m_prev; o.act_prev = o_activated;
                    sts_v2f64(sh_a, mx, pack_bound(runner_up_bound(second, mx),
                                                   (unsigned)arg | ((o_activated | activates) ? kActivatedBit : 0u)));
                }
                __syncwarp();
                pend &= ~(fin | (1u << j));
                mine = mine && lane_ > j;
                if (pend == 0) break;
                lds_v2f64(sh_a, h.smax, h.sub2);
                round(std::false_type{});
                if (fm == 0) break;
            }
        }
    };

    // ---- outputs of a batch: overall_value prefix over the batch (ST2:99-105), stores
    auto fin_out = [&](auto tail_tag, const unsigned k, const Rows &r, const Sel &o) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const bool in_stream = TAIL ? r.active : true;
        const unsigned e = ebase + (k - lane);
        if (in_stream) {
            st_stream(oval + e, (OT)o.m_out);
            st_stream(oact + e, (KT)o.g_out);
        }
        if (OVERALL) {
            // rows of states that were active before the row: the bonus cancels in after - before
            const double d = (r.active && o.act_prev) ? o.m_out - o.mprev_out : 0.0;
            double pre, batch_sum;   // inclusive prefix of the deltas at this row, and over the whole batch
            if constexpr (SCAN_F32) {
                float f = (float)d;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) f = scan_step(f, off);
                pre = (double)f;
                batch_sum = (double)__shfl_sync(FULL_MASK, f, 31);
            } else {
                pre = d;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) pre = scan_step(pre, off);
                batch_sum = shfl_f64(pre, 31);
            }
            if (o.any_jump) {   // rare: a state activates at this row, its (max + bonus) enters the sum
                double jmp = o.jump ? o.m_out + bonus : 0.0;
#pragma unroll
                for (int off = 1; off < 32; off <<= 1) jmp = scan_step(jmp, off);
                pre += jmp;
                batch_sum += shfl_f64(jmp, 31);
            }
            const double ov = total + pre;
            total += batch_sum;
            if (in_stream) st_stream(oall + e, (OT)ov);
        }
    };

    const unsigned nfull = n_steps >> 5;
    unsigned k = lane;
    Rows cur;
    Mom mom;
    Sel sel;
    if (nfull) {
        p1a_issue(std::false_type{}, k, cur, mom);
        p1a_peers(cur, mom);
        p1_peers(mom);
        p1b(cur, mom);
        __syncwarp();
        for (unsigned b = 1; b < nfull; ++b) {
            Rows nxt;
            p1a_issue(std::false_type{}, k + 32, nxt, mom);   // reads {cnt, s1, s2} only
            p2(cur, sel);                               // reads/writes {val, state records}; cnt read-only
            p1a_peers(nxt, mom);
            p1_peers(mom);
            p1b(nxt, mom);                              // writes {cnt, s1, s2}
            fin_out(std::false_type{}, k, cur, sel);
            __syncwarp();   // table writes of this iteration are ordered before the next one's reads
            cur = nxt;
            k += 32;
        }
        p2(cur, sel);
        fin_out(std::false_type{}, k, cur, sel);
        __syncwarp();
        k += 32;
    }
    if (k - lane < n_steps) {   // ragged last batch: lanes past the end of the stream are inactive
        p1a_issue(std::true_type{}, k, cur, mom);
        p1a_peers(cur, mom);
        p1_peers(mom);
        p1b(cur, mom);
        __syncwarp();
        p2(cur, sel);
        fin_out(std::true_type{}, k, cur, sel);
        __syncwarp();
    }
#ifdef DCARL_DEBUG_TIMING
    if (lane == 0 && inst < 65536) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g_rows_timing[inst * 4 + 1] = t;
    }
#endif
    for (int s = lane; s < S; s += 32) p.out_activation[inst * S + s] = (long long)actstep[s];
    if (p.act_hist != nullptr) {
        for (int s = lane; s < S; s += 32) {
            unsigned csum = 0;
            for (int q = 0; q < A; ++q) csum += cnt[s * A + q];
            atomicAdd(&hist[bound_payload(hot[s].sub2) & kArgMask], csum - (unsigned)flushpt[s]);
        }
        __syncwarp();
        for (int a = lane; a < A; a += 32)
            if (hist[a]) atomicAdd(p.act_hist + a, (unsigned long long)hist[a]);
    }
    if (p.bad_rows != nullptr && __any_sync(FULL_MASK, clamped != 0)) {
        // cold path: some key of this instance was out of range — count the offending rows exactly
        unsigned nbad = 0;
        for (unsigned q = lane; q < n_steps; q += 32) {
            const unsigned e = (unsigned)inst * n_steps + q;
            nbad += ((unsigned)gidx[e] >= (unsigned)S) | ((unsigned)gact[e] >= (unsigned)A);
        }
        nbad = __reduce_add_sync(FULL_MASK, nbad);
        if (lane == 0 && nbad) atomicAdd(p.bad_rows, (unsigned long long)nbad);
    }
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
        warp_top2(val, 1, A, m0, g0, sec0);
        for (int s = lane; s < S; s += 32) { smax[s] = m0; steps[s] = 0; actstep[s] = -1; }
    }
    __syncwarp();

    const long long n_steps = p.n_steps;
    const long long row0 = inst * n_steps;
    const RT *gR = reinterpret_cast<const RT *>(p.R) + row0;
    OT *oval = reinterpret_cast<OT *>(p.out_value) + row0;
    OT *oall = OVERALL ? reinterpret_cast<OT *>(p.out_overall) + row0 : nullptr;
    unsigned long long bad = 0;

    for (long long k0 = 0; k0 < n_steps; k0 += 32) {
        const long long k = k0 + lane;
        int b_idx = 0, b_act = 0;
        double b_r = 0.0;
        if (k < n_steps) {
            b_idx = reinterpret_cast<const int *>(p.idx)[row0 + k];
            b_act = reinterpret_cast<const int *>(p.act)[row0 + k];
            b_r = (double)gR[k];
        }
        double o_val = CUDART_NAN, o_all = 0.0;
        int o_act = -1;
        const int nrow = (int)min((long long)32, n_steps - k0);
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
                warp_top2(val + s * A, 1, A, mx, arg, second);
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
        if (k < n_steps) {
            oval[k] = (OT)o_val;
            reinterpret_cast<int *>(p.out_act)[row0 + k] = o_act;
            if (OVERALL) oall[k] = (OT)o_all;
            if (p.act_hist != nullptr && o_act >= 0) atomicAdd(p.act_hist + o_act, 1ull);
        }
    }
    __syncwarp();
    for (int s = lane; s < S; s += 32) p.out_activation[inst * S + s] = (long long)actstep[s];
    if (p.bad_rows != nullptr && bad) atomicAdd(p.bad_rows, bad);
}

// ---- summary statistics (fixed-order float64 reduction, one thread-block cluster) ---------------
// One CTA was fine for the 4096 instances of config 2 (20 us) and the bottleneck of config 5: 131 072 instances
// x 20 activation steps are 21 MB, ~0.3 ms through one SM — a quarter of the batch.  Now the kStatsCtas CTAs of
// one cluster take contiguous instance ranges, hand their four partial sums to CTA 0 through distributed shared
// memory, and CTA 0 adds them in rank order (then the sum over ranks, as before): still one launch, no
// workspace, and a fixed order of additions for a given n_inst.
constexpr int kStatsCtas = 8;   // portable cluster size

template <typename OT>
__global__ void __launch_bounds__(1024) k_stream_stats(const OT *__restrict__ out_value,
                                                        const OT *__restrict__ out_overall,
                                                        const long long *__restrict__ activation,
                                                        const unsigned long long *__restrict__ act_hist,
                                                        long long n_inst, long long n_steps, int S, int A,
                                                        double *stats, int width, int accumulate,
                                                        const PeerCtx peers) {
    __shared__ double sh[4][1024];
    __shared__ double part[kStatsCtas][4];          // CTA 0's copy collects the partial sums of the cluster
    const int t = threadIdx.x;
    const unsigned cr = cluster_ctarank(), nc = cluster_nctarank();
    const long long per = (n_inst + nc - 1) / nc;
    const long long i_lo = min(n_inst, (long long)cr * per), i_hi = min(n_inst, i_lo + per);
    double v = 0.0, o = 0.0, na = 0.0, sa = 0.0;
    if (n_steps > 0) {
        const OT *lv = out_value + n_steps - 1;
        const OT *lo = out_overall ? out_overall + n_steps - 1 : nullptr;
        for (long long i = i_lo + t; i < i_hi; i += 4096) {   // four independent loads in flight per array
            OT x[4], y[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const long long ii = i + 1024ll * u;
                x[u] = ii < i_hi ? lv[ii * n_steps] : OT(0);
                y[u] = (lo && ii < i_hi) ? lo[ii * n_steps] : OT(0);
            }
#pragma unroll
            for (int u = 0; u < 4; ++u) { v += (double)x[u]; o += (double)y[u]; }
        }
    }
    const long long a_lo = i_lo * S, a_hi = i_hi * S;
    for (long long i = a_lo + t; i < a_hi; i += 8192) {
        long long st[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) { const long long ii = i + 1024ll * u; st[u] = ii < a_hi ? activation[ii] : -1; }
#pragma unroll
        for (int u = 0; u < 8; ++u)
            if (st[u] != -1) { na += 1.0; sa += (double)st[u]; }
    }
    sh[0][t] = v; sh[1][t] = o; sh[2][t] = na; sh[3][t] = sa;
    __syncthreads();
    for (int w = 512; w > 0; w >>= 1) {
        if (t < w)
            for (int q = 0; q < 4; ++q) sh[q][t] += sh[q][t + w];
        __syncthreads();
    }
    if (nc > 1) cluster_arrive_wait_entered();   // (distributed shared memory may only be touched once its CTA runs)
    if (t < 4) st_shared_cluster(mapa_shared(smem_u32(&part[cr][t]), 0u), sh[t][0]);
    cluster_sync_all();
    if (cr != 0) return;
    // this call's vector, zero-padded to `width`, optionally on top of what `stats` holds (mixed batches)
    for (int i = t; i < width; i += 1024) {
        double x = 0.0;
        if (i < 4) { for (unsigned r = 0; r < nc; ++r) x += part[r][i]; }
        else if (i == 4) x = (double)n_inst * (double)n_steps;
        else if (i - DCARL_STATS_HEADER < A) x = act_hist ? (double)act_hist[i - DCARL_STATS_HEADER] : 0.0;
        stats[i] = accumulate ? stats[i] + x : x;
    }
    if (peers.world > 1) {          // one kernel: local reduction, then the sum over ranks through peer memory
        __syncthreads();
        peer_allreduce_cta(stats, stats, width, peers);
    }
}

// ---- the closed form on caller-held moments (also the precision probe of cell_value) ---------
__global__ void __launch_bounds__(256) k_cell_values(const int *__restrict__ n, const double *__restrict__ s1,
                                                      const double *__restrict__ s2,
                                                      const int *__restrict__ act, long long count, double c_h,
                                                      double cap, int rule_act, double *__restrict__ out) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    out[i] = n[i] > 0 ? cell_value((unsigned)n[i], s1[i], s2[i], act[i] == rule_act, c_h, cap) : CUDART_NAN;
}

// ---- host side -------------------------------------------------------------------------------
static const size_t kMaxSmemPerCta = 227 * 1024;

// Warps (= instances) per CTA.  Measured on config 2 (4096 x 20000, one resident wave): one 28-warp
// CTA per SM 0.96 ms, four 7-warp CTAs 1.07 ms — the arbiter favours the warps of some CTAs, they
// finish early and leave their scheduler under-occupied.  With many waves (config 5, 32768 x 500)
// small CTAs refill an SM sooner: 4 warps 0.180 ms, 28 warps 0.188 ms.
static int pick_warps_per_cta(long long n_inst, size_t wb, int sms) {
    long long fit = (long long)((kMaxSmemPerCta - kRowsCtaFixedBytes) / wb);   // warps whose tables fit one CTA
    if (fit > kRowsMaxWarps) fit = kRowsMaxWarps;
    if (fit < 1) fit = 1;
    if (n_inst <= fit * sms) {                                 // one wave: one CTA per SM
        const long long w = (n_inst + sms - 1) / sms;
        return (int)(w < 1 ? 1 : w);
    }
    return (int)(fit < 4 ? fit : 4);
}

// The argmax of the priors (ST1:51-53): when it is not the rule action, every state activates on its
// first row without any change of argmax, which the kernel handles by forcing those rows through its
// exact path (FORCE_FIRST).
static bool priors_argmax_is_rule(const StreamArgs &a) {
    if (a.prior_rule > a.prior_other) return true;
    if (a.prior_rule == a.prior_other) return a.rule_act == 0;
    return a.A == 1;
}

template <typename RT, typename OT, typename KT, bool OV, bool FF>
static int launch_rows_ff(const StreamArgs &a, cudaStream_t st) {
    const size_t wb = rows_warp_bytes(a.S, a.A);
    const int sms = sm_count();
    const int wpc = pick_warps_per_cta(a.n_inst, wb, sms);
    const size_t smem = kRowsCtaFixedBytes + wb * wpc;
    auto kern = k_stream_rows<RT, OT, KT, OV, FF>;
    // per launch, not cached: the attribute belongs to the current device's context
    DCARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemPerCta));
    DCARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, 100));
    // the kernel indexes rows with 32 bits: launch in instance chunks of < 2^32 rows each
    const long long max_inst = a.n_steps ? (long long)(0xffffff00u / a.n_steps) : a.n_inst;
    for (long long lo = 0; lo < a.n_inst; lo += max_inst) {
        StreamArgs c = a;
        c.n_inst = a.n_inst - lo < max_inst ? a.n_inst - lo : max_inst;
        const long long off = lo * (long long)a.n_steps;
        c.idx = reinterpret_cast<const KT *>(a.idx) + off;
        c.act = reinterpret_cast<const KT *>(a.act) + off;
        c.out_act = reinterpret_cast<KT *>(a.out_act) + off;
        c.R = reinterpret_cast<const RT *>(a.R) + off;
        c.out_value = reinterpret_cast<OT *>(a.out_value) + off;
        if (OV) c.out_overall = reinterpret_cast<OT *>(a.out_overall) + off;
        c.out_activation = a.out_activation + lo * a.S;
        const long long grid = (c.n_inst + wpc - 1) / wpc;
        kern<<<(unsigned)grid, wpc * 32, smem, st>>>(c, wpc, (unsigned)wb);
        DCARL_LAUNCH_CHECK("k_stream_rows");
    }
    return 0;
}

template <typename RT, typename OT, typename KT, bool OV>
static int launch_rows(const StreamArgs &a, cudaStream_t st) {
    return priors_argmax_is_rule(a) ? launch_rows_ff<RT, OT, KT, OV, false>(a, st)
                                    : launch_rows_ff<RT, OT, KT, OV, true>(a, st);
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
        return rows ? launch_rows<RT, OT, int, true>(a, st) : launch_walk<RT, OT, true>(a, ws, st);
    return rows ? launch_rows<RT, OT, int, false>(a, st) : launch_walk<RT, OT, false>(a, ws, st);
}

}  // namespace dcarl

using namespace dcarl;

extern "C" size_t dcarl_stream_bounds_ws(int64_t n_inst, int64_t n_steps, int32_t S, int32_t A) {
    (void)n_steps;
    if (n_inst <= 0 || S <= 0 || A <= 0) return 0;
    return (size_t)n_inst * walk_inst_bytes(S, A);
}

extern "C" int dcarl_stream_rows_fit(int32_t S, int32_t A) {
    if (S <= 0 || A <= 0 || (long long)S * A > (1ll << 28)) return 0;
    return kRowsCtaFixedBytes + rows_warp_bytes(S, A) <= kMaxSmemPerCta ? 1 : 0;
}

// Shared argument checks of the two stream entry points; fills `a` (everything but the row pointers).
static int stream_args(int64_t n_inst, int64_t n_steps, int32_t S, int32_t A, const dcarl_params *p,
                       int64_t *out_activation_step, void *out_overall, uint64_t *act_hist,
                       uint64_t *bad_rows, StreamArgs &a) {
    if (n_inst < 0) return bad_arg(5, "n_inst < 0");
    if (n_steps < 0) return bad_arg(6, "n_steps < 0");
    if (S <= 0) return bad_arg(7, "S <= 0");
    if (A <= 0) return bad_arg(8, "A <= 0");
    if ((long long)S * A > (1ll << 28)) return bad_arg(7, "S*A too large");
    if (!p) return bad_arg(9, "params is null");
    if (!(p->alpha > 0.0 && p->alpha <= 1.0)) return bad_arg(9, "alpha must be in (0, 1]");
    if (p->rule_act < 0 || p->rule_act >= A) return bad_arg(9, "rule_act outside [0, A)");
    if (p->n_thres < 0) return bad_arg(9, "n_thres < 0");
    if (n_steps > 0x7fffffc0ll) return bad_arg(6, "n_steps does not fit the 32-bit row counters");
    if ((n_inst + 3) / 4 > 0x7fffffffll) return bad_arg(5, "n_inst too large for one launch");
    if (n_inst > 0 && !out_activation_step) return bad_arg(13, "out_activation_step is null");
    a.out_activation = reinterpret_cast<long long *>(out_activation_step);
    a.out_overall = out_overall;
    a.act_hist = reinterpret_cast<unsigned long long *>(act_hist);
    a.bad_rows = reinterpret_cast<unsigned long long *>(bad_rows);
    a.n_inst = n_inst; a.n_steps = (unsigned)n_steps; a.S = S; a.A = A;
    a.c_h = p->scale * sqrt(log(1.0 / p->alpha) / 2.0);
    a.cap = p->cap; a.prior_rule = p->prior_rule; a.prior_other = p->prior_other;
    a.bonus = p->activation_bonus; a.n_thres = p->n_thres; a.rule_act = p->rule_act;
    return 0;
}

extern "C" int dcarl_stream_bounds(const int32_t *idx, const int32_t *act, const void *R, int r_is_f64,
                                   int64_t n_inst, int64_t n_steps, int32_t S, int32_t A,
                                   const dcarl_params *p, void *out_value, int out_is_f64,
                                   int32_t *out_act, int64_t *out_activation_step, void *out_overall,
                                   uint64_t *act_hist, uint64_t *bad_rows, void *ws, size_t ws_bytes,
                                   int variant, dcarl_stream_t stream) {
    StreamArgs a;
    if (int rc = stream_args(n_inst, n_steps, S, A, p, out_activation_step, out_overall, act_hist, bad_rows, a))
        return rc;
    if (n_inst == 0) return 0;
    if (n_steps > 0 && (!idx || !act || !R)) return bad_arg(1, "idx/act/R is null");
    if (n_steps > 0 && (!out_value || !out_act)) return bad_arg(10, "out_value/out_act is null");
    if (variant < 0 || variant > 2) return bad_arg(19, "variant must be 0, 1 or 2");

    const bool fits = kRowsCtaFixedBytes + rows_warp_bytes(S, A) <= kMaxSmemPerCta;
    bool rows = fits;
    if (variant == 1) rows = false;
    if (variant == 2 && !fits) return bad_arg(19, "the S*A table does not fit shared memory");
    if (!rows) {
        const size_t need = dcarl_stream_bounds_ws(n_inst, n_steps, S, A);
        if (!ws || ws_bytes < need) return bad_arg(17, "workspace too small (see dcarl_stream_bounds_ws)");
    }
    a.idx = idx; a.act = act; a.R = R;
    a.out_value = out_value; a.out_act = out_act;
    cudaStream_t st = (cudaStream_t)stream;
    if (r_is_f64)
        return out_is_f64 ? dispatch_ov<double, double>(a, rows, ws, st) : dispatch_ov<double, float>(a, rows, ws, st);
    return out_is_f64 ? dispatch_ov<float, double>(a, rows, ws, st) : dispatch_ov<float, float>(a, rows, ws, st);
}

// Compact wire format: one byte per key.  Same kernel, keys widened at the load and the argmax
// narrowed at the store; float32 returns in, float32 values out.
extern "C" int dcarl_stream_bounds_u8(const uint8_t *idx, const uint8_t *act, const float *R,
                                      int64_t n_inst, int64_t n_steps, int32_t S, int32_t A,
                                      const dcarl_params *p, float *out_value, uint8_t *out_act,
                                      int64_t *out_activation_step, float *out_overall,
                                      uint64_t *act_hist, uint64_t *bad_rows, dcarl_stream_t stream) {
    StreamArgs a;
    if (int rc = stream_args(n_inst, n_steps, S, A, p, out_activation_step, out_overall, act_hist, bad_rows, a))
        return rc;
    if (S > 256) return bad_arg(6, "S > 256 does not fit one-byte keys");
    if (A > 256) return bad_arg(7, "A > 256 does not fit one-byte keys");
    if (kRowsCtaFixedBytes + rows_warp_bytes(S, A) > kMaxSmemPerCta)
        return bad_arg(6, "the S*A table does not fit shared memory (use dcarl_stream_bounds)");
    if (n_inst == 0) return 0;
    if (n_steps > 0 && (!idx || !act || !R)) return bad_arg(1, "idx/act/R is null");
    if (n_steps > 0 && (!out_value || !out_act)) return bad_arg(9, "out_value/out_act is null");
    a.idx = idx; a.act = act; a.R = R;
    a.out_value = out_value; a.out_act = out_act;
    cudaStream_t st = (cudaStream_t)stream;
    return out_overall ? launch_rows<float, float, uint8_t, true>(a, st)
                       : launch_rows<float, float, uint8_t, false>(a, st);
}

#ifdef DCARL_DEBUG_TIMING
extern "C" int dcarl_debug_rows_timing(unsigned long long *host_out, int64_t n_inst) {
    DCARL_CUDA(cudaDeviceSynchronize());
    DCARL_CUDA(cudaMemcpyFromSymbol(host_out, g_rows_timing, sizeof(unsigned long long) * 4 * (size_t)n_inst));
    return 0;
}
#endif

extern "C" int dcarl_cell_values(const int32_t *n, const double *sum, const double *sumsq, const int32_t *act,
                                 int64_t count, const dcarl_params *p, double *out, dcarl_stream_t stream) {
    if (count < 0) return bad_arg(5, "count < 0");
    if (!p) return bad_arg(6, "params is null");
    if (!(p->alpha > 0.0 && p->alpha <= 1.0)) return bad_arg(6, "alpha must be in (0, 1]");
    if (count == 0) return 0;
    if (!n || !sum || !sumsq || !act) return bad_arg(1, "null input");
    if (!out) return bad_arg(7, "out is null");
    const double c_h = p->scale * sqrt(log(1.0 / p->alpha) / 2.0);
    k_cell_values<<<(unsigned)((count + 255) / 256), 256, 0, (cudaStream_t)stream>>>(n, sum, sumsq, act, count, c_h,
                                                                                     p->cap, p->rule_act, out);
    DCARL_LAUNCH_CHECK("k_cell_values");
    return 0;
}

extern "C" int dcarl_stream_stats_reduce(const void *out_value, int out_is_f64, const void *out_overall,
                                         const int64_t *out_activation_step, const uint64_t *act_hist,
                                         int64_t n_inst, int64_t n_steps, int32_t S, int32_t A, double *stats,
                                         int32_t stats_width, int accumulate, const dcarl_peer_ctx *peers,
                                         dcarl_stream_t stream) {
    if (n_inst < 0 || n_steps < 0) return bad_arg(6, "negative size");
    if (S <= 0 || A <= 0) return bad_arg(8, "S/A <= 0");
    if (!stats) return bad_arg(10, "stats is null");
    if (stats_width < DCARL_STATS_HEADER + A) return bad_arg(11, "stats_width < DCARL_STATS_HEADER + A");
    if (n_inst > 0 && !out_activation_step) return bad_arg(4, "out_activation_step is null");
    if (n_inst > 0 && n_steps > 0 && !out_value) return bad_arg(1, "out_value is null");
    PeerCtx pc;
    if (int rc = make_peer_ctx(peers, stats_width, 13, pc)) return rc;
    cudaStream_t st = (cudaStream_t)stream;
    const long long *as = reinterpret_cast<const long long *>(out_activation_step);
    const unsigned long long *ah = reinterpret_cast<const unsigned long long *>(act_hist);
    cudaLaunchConfig_t cfg = {};
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    // a small batch (config 2: 4096 instances, 0.7 MB) stays ONE CTA: it then fits beside the next batch's stream
    // kernel (147 CTAs on 148 SMs) when the caller pipelines batches over two streams
    const unsigned ctas = n_inst >= 16384 ? kStatsCtas : 1;
    attr[0].val.clusterDim.x = ctas; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.gridDim = dim3(ctas); cfg.blockDim = dim3(1024); cfg.dynamicSmemBytes = 0; cfg.stream = st;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int s_ = S, a_ = A, w_ = stats_width, acc_ = accumulate;
    const long long ni = n_inst, ns = n_steps;
    if (out_is_f64)
        DCARL_CUDA(cudaLaunchKernelEx(&cfg, k_stream_stats<double>, (const double *)out_value, (const double *)out_overall,
                                      as, ah, ni, ns, s_, a_, stats, w_, acc_, pc));
    else
        DCARL_CUDA(cudaLaunchKernelEx(&cfg, k_stream_stats<float>, (const float *)out_value, (const float *)out_overall,
                                      as, ah, ni, ns, s_, a_, stats, w_, acc_, pc));
    DCARL_LAUNCH_CHECK("k_stream_stats");
    return 0;
}

extern "C" int dcarl_stream_stats(const void *out_value, int out_is_f64, const void *out_overall,
                                  const int64_t *out_activation_step, const uint64_t *act_hist,
                                  int64_t n_inst, int64_t n_steps, int32_t S, int32_t A, double *stats,
                                  dcarl_stream_t stream) {
    return dcarl_stream_stats_reduce(out_value, out_is_f64, out_overall, out_activation_step, act_hist, n_inst,
                                     n_steps, S, A, stats, DCARL_STATS_HEADER + (A > 0 ? A : 0), 0, nullptr, stream);
}
