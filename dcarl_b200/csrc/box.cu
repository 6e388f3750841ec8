// Box-kernel confidence over the collected {state (+) action, return} buffer.
//
// Replaces RLS._calculate_visited_times / _calculate_statistics_index
// (Field_testing/.../tools/DCARL/stable_baselines/deepq/RLS.py:161-181), which walk an
// rtree of boxes [s_i - dist, s_i + dist], and the candidate loop of RLS.act_test
// (RLS.py:120-157), which issues two tree queries per candidate action.  Here one
// streaming pass over the buffer answers every (query, candidate action) pair at once:
// the observation dims are tested once per row, the row's own action picks the bin.
//
// Exactness: the reference tests  fl64(s - dist) <= q <= fl64(s + dist)  in float64.
// For a float32 s both sides are monotone in s, so per (query, dim) there is a largest
// float32 s_hi and a smallest float32 s_lo with  match <=> s_lo <= s <= s_hi.
// k_box_prepare finds them (exact float64 predicate, +-1 ulp walk); the scan then needs
// two float32 compares per dim and still reproduces the float64 decision bit for bit.
//
// Data movement: the AoS buffer [M, D+1] (float32, or float64 in the reference-parity storage
// mode) is staged tile by tile into shared memory with 1-D TMA bulk copies (cp.async.bulk +
// mbarrier), so global reads are whole 16-byte-aligned contiguous tiles regardless of the odd
// row pitch, and a thread reads "its" row from shared memory (pitch D+1 = 21 words: conflict free).
//
// Kernels: k_box_moments (one or a few queries: CTA-wide 256-row tiles, 4-stage ring),
// k_box_filter (eight or more queries: per-dimension bitmap pre-filter, per-warp 32-row rings),
// both also as the whole act_test query in one launch (FUSED: thresholds in the prologue, two
// alternating accumulators, totals + sum over ranks + select in the last CTA's epilogue);
// k_box_reduce_select (many queries: sum over ranks + select, one CTA per query);
// k_box_prepare / k_box_select (the unfused thresholds / select).
#include <math_constants.h>

#include <type_traits>

#include "common.cuh"
#include "peer.cuh"

namespace dcarl {

constexpr int kTileRows = 256;
constexpr int kStages = 4;       // plain scan (HBM bound: deep ring)
constexpr int kMaxSmemBins = 1024;  // (query, action) bins kept in shared memory: 24 KB

// Largest stored value s_hi and smallest stored value s_lo (in the buffer's element type BT) with
//   fl64(s - w) <= q <= fl64(s + w)  <=>  s_lo <= s <= s_hi
// (exact float64 predicate, +-1 ulp walk around the rounded candidates); an empty interval for NaN inputs.
__device__ __forceinline__ float next_up(float x) { return nextafterf(x, CUDART_INF_F); }
__device__ __forceinline__ float next_dn(float x) { return nextafterf(x, -CUDART_INF_F); }
__device__ __forceinline__ double next_up(double x) { return nextafter(x, CUDART_INF); }
__device__ __forceinline__ double next_dn(double x) { return nextafter(x, -CUDART_INF); }
__device__ __forceinline__ float narrow_dn(float v) { return v; }
__device__ __forceinline__ float narrow_up(float v) { return v; }
__device__ __forceinline__ float narrow_dn(double v) { return __double2float_rd(v); }
__device__ __forceinline__ float narrow_up(double v) { return __double2float_ru(v); }
__device__ __forceinline__ void round_dn_up(double v, float &dn, float &up) { dn = __double2float_rd(v); up = __double2float_ru(v); }
__device__ __forceinline__ void round_dn_up(double v, double &dn, double &up) { dn = v; up = v; }

template <typename BT>
__device__ __forceinline__ void box_thresholds(double q, double w, BT &l, BT &h) {
    h = (BT)-CUDART_INF; l = (BT)CUDART_INF;
    if (q == q && w == w) {
        BT c, unused;
        // hi: largest s with (double)s - w <= q
        round_dn_up(q + w, c, unused);
        for (int it = 0; it < 4; ++it) {
            const BT up = next_up(c);
            if (up != c && (double)up - w <= q) c = up; else break;
        }
        for (int it = 0; it < 4 && !((double)c - w <= q); ++it) c = next_dn(c);
        if ((double)c - w <= q) h = c;
        // lo: smallest s with q <= (double)s + w
        round_dn_up(q - w, unused, c);
        for (int it = 0; it < 4; ++it) {
            const BT dn = next_dn(c);
            if (dn != c && q <= (double)dn + w) c = dn; else break;
        }
        for (int it = 0; it < 4 && !(q <= (double)c + w); ++it) c = next_up(c);
        if (q <= (double)c + w) l = c;
    }
}

// thresholds layout in the workspace: lo[Q][D], hi[Q][D] (element type BT)
template <typename BT>
__global__ void k_box_prepare(const double *__restrict__ queries, const double *__restrict__ dist, int Q,
                              int D, BT *__restrict__ lo, BT *__restrict__ hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * D) return;
    BT l, h;
    box_thresholds<BT>(queries[i], dist[i % D], l, h);
    lo[i] = l;
    hi[i] = h;
}

struct BoxArgs {
    const void *buf_state;     // [M, D+1] float32, or float64 in the reference-parity storage mode
    const void *buf_ret;       // [M]
    long long M;
    int D, Q, n_actions;
    const void *lo, *hi;       // [Q, D] thresholds in the buffer's element type (unfused path)
    const double *dist;  // [D+1]; dist[D] is the action tolerance
    double *out;  // [Q, n_actions, 3], zero on entry: the totals of the unfused scan
    int bins_in_smem;
    // fused query (FUSED kernels): thresholds in the prologue, reduction + select in the epilogue.  Two
    // accumulators alternate between calls (hdr[1] counts the calls): a call adds into one and zeroes the other
    // for the next call on its way in, spread over all CTAs, so nothing is memset or copied per query.
    const double *queries;     // [Q, D]
    unsigned *hdr;             // {ticket: CTAs that have flushed their bins, calls, extent: doubles the last call used,
                               //  pad, two tickets of k_box_reduce_select, pad, pad}
    double *acc[2];            // the two accumulators (whole workspace halves)
    unsigned q0_off;           // this launch's first accumulator element (queries are chunked over launches)
    unsigned n3_total;         // accumulator elements of the whole call
    int last_launch;           // the last launch of a call advances hdr[1]
    int epilogue;              // 0: the scan only accumulates (k_box_reduce_select follows)
    // unfused k_box_filter on a small (L2-resident) buffer with more queries than one pass takes: n_chunks > 1 —
    // every CTA takes whole chunks of Q queries (lo / hi / out advance by chunk; the last chunk holds
    // q_total - (n_chunks - 1) * Q queries) and scans the WHOLE buffer for each, in ONE launch
    int n_chunks, q_total;
    double *out_moments;       // [Q, n_actions, 3] totals (summed over ranks when peers.world > 1)
    double *out_mv;            // [Q, n_actions, 2], nullable
    int *out_action;           // [Q], nullable
    dcarl_rls_params rp;
    PeerCtx peers;
};

// Adds (1, r, r*r) of every lane with `hit` into bins[bin*3 ..]; lanes hitting the same bin
// are combined with shuffles first, so a dense match costs one atomic triple per bin per warp.
template <bool SMEM>
__device__ __forceinline__ void warp_accumulate(bool hit, int bin, double r, double *bins) {
    const int lane = lane_id();
    const unsigned mh = __match_any_sync(FULL_MASK, hit ? bin : (0x40000000 | lane));
    const bool leader = hit && (mh & lanemask_lt()) == 0;
    unsigned rest = leader ? (mh & lanemask_gt()) : 0u;
    double s1 = r, s2 = r * r;
    while (__any_sync(FULL_MASK, rest != 0)) {
        const int j = rest ? (__ffs(rest) - 1) : lane;
        const double rj = shfl_f64(r, j);
        if (rest) { s1 += rj; s2 = fma(rj, rj, s2); rest &= rest - 1; }
    }
    if (leader) {
        double *b = bins + (size_t)bin * 3;
        atomicAdd(b + 0, (double)__popc(mh));
        atomicAdd(b + 1, s1);
        atomicAdd(b + 2, s2);
    }
    (void)SMEM;
}

// moments -> mean/var (+ sentinel) and the act_test pick of one query
__device__ __forceinline__ void box_select_one(const double *m, int n_actions, const dcarl_rls_params &rp,
                                               double *out_mv, int *out_action) {
    auto mean_var = [&](int a, double &n, double &mean, double &var) {
        n = m[a * 3];
        if (n > 0.0) {
            mean = m[a * 3 + 1] / n;
            var = fmax(m[a * 3 + 2] / n - mean * mean, 0.0);
        } else { mean = -1.0; var = -1.0; }                                    // RLS.py:169-170
    };
    double n0, mean0, var0;
    mean_var(0, n0, mean0, var0);
    int pick = 0;
    for (int a = 0; a < n_actions; ++a) {
        double n, mean, var;
        mean_var(a, n, mean, var);
        if (out_mv) { out_mv[a * 2] = mean; out_mv[a * 2 + 1] = var; }
        if (a == 0 || pick != 0) continue;
        if (n0 < (double)rp.visited_times_thres || n < (double)rp.min_candidate_visits || mean0 > rp.rule_mean_gate)
            continue;                                                          // RLS.py:141
        const double z = (mean - mean0) / sqrt(var0 / n0 + var / n);         // RLS.py:144-148
        const double cdf = 0.5 * erfc(-z * 0.70710678118654752440);
        if (cdf > rp.confidence_thres) pick = a;                               // RLS.py:150-154
    }
    if (out_action) *out_action = pick;
}

// Fused prologue (every thread of every CTA): this call's accumulator, and the other one zeroed for the next call.
__device__ __forceinline__ double *box_fused_accumulator(const BoxArgs &p, const int tid, const int nthr) {
    const unsigned calls = p.hdr[1], ext = p.hdr[2];
    double *oth = p.acc[(calls & 1u) ^ 1u];
    const unsigned nz = ext > p.n3_total ? ext : p.n3_total;
    for (unsigned i = blockIdx.x * (unsigned)nthr + tid; i < nz; i += gridDim.x * (unsigned)nthr) oth[i] = 0.0;
    return p.acc[calls & 1u] + p.q0_off;
}

// Fused epilogue of a scan launch (all threads of every CTA call it after their bins reached the accumulator
// `acc`): the last CTA to arrive owns the totals — it adds the other ranks' totals through peer memory and
// writes out_moments, (mean, var) and the act_test pick in one pass over the (query, action) pairs.
__device__ __forceinline__ void box_epilogue(const BoxArgs &p, const double *acc, double *scratch /* shared, 3 Q doubles */,
                                             const int tid, const int nthr) {
    const int nbins = p.Q * p.n_actions;
    __shared__ unsigned s_ticket;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(p.hdr, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1u) return;
    __threadfence();
    const bool solo = p.peers.world <= 1;
    const double *mom = acc;                 // totals: the accumulator itself, or out_moments after the sum over ranks
    if (!solo) {
        peer_allreduce_cta(acc, p.out_moments, nbins * 3, p.peers);
        __syncthreads();
        mom = p.out_moments;
    }
    // one (query, action) pair per thread and round, the loads of four pairs in flight (this is one CTA: L2
    // latency, not bandwidth).  The pick is the smallest action index that passes the gate and the z-test
    // (RLS.py:135-154): an atomic min.
    // (the rule action's statistics once per query, not once per pair: double divisions are what this CTA spends
    // its time on)
    for (int q = tid; q < p.Q; q += nthr) {
        if (p.out_action) p.out_action[q] = 0x7fffffff;
        const double *m0 = mom + (size_t)q * p.n_actions * 3;
        const double n0 = __ldcg(m0);
        const double mean0 = n0 > 0.0 ? __ldcg(m0 + 1) / n0 : -1.0;
        scratch[3 * q] = n0;
        scratch[3 * q + 1] = mean0;
        scratch[3 * q + 2] = n0 > 0.0 ? fmax(__ldcg(m0 + 2) / n0 - mean0 * mean0, 0.0) : -1.0;
    }
    __syncthreads();
    for (int i0 = 0; i0 < nbins; i0 += nthr * 4) {
        double n_[4], s1_[4], s2_[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = min(i0 + u * nthr + tid, nbins - 1);
            n_[u] = __ldcg(mom + (size_t)i * 3); s1_[u] = __ldcg(mom + (size_t)i * 3 + 1); s2_[u] = __ldcg(mom + (size_t)i * 3 + 2);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int i = i0 + u * nthr + tid;
            if (i >= nbins) continue;
            const int q = i / p.n_actions, a = i - q * p.n_actions;
            if (solo) { p.out_moments[(size_t)i * 3] = n_[u]; p.out_moments[(size_t)i * 3 + 1] = s1_[u]; p.out_moments[(size_t)i * 3 + 2] = s2_[u]; }
            const double n0 = scratch[3 * q], mean0 = scratch[3 * q + 1], var0 = scratch[3 * q + 2];
            const double n = n_[u];
            double mean = -1.0, var = -1.0;                                    // RLS.py:169-170
            if (n > 0.0) { mean = s1_[u] / n; var = fmax(s2_[u] / n - mean * mean, 0.0); }
            if (p.out_mv) { p.out_mv[(size_t)i * 2] = mean; p.out_mv[(size_t)i * 2 + 1] = var; }
            if (!p.out_action || a == 0 || n0 < (double)p.rp.visited_times_thres ||
                n < (double)p.rp.min_candidate_visits || mean0 > p.rp.rule_mean_gate)
                continue;                                                      // RLS.py:141
            const double z = (mean - mean0) / sqrt(var0 / n0 + var / n);     // RLS.py:144-148
            if (0.5 * erfc(-z * 0.70710678118654752440) > p.rp.confidence_thres) atomicMin(p.out_action + q, a);   // RLS.py:150-154
        }
    }
    __syncthreads();
    if (p.out_action)
        for (int q = tid; q < p.Q; q += nthr)
            if (__ldcg(p.out_action + q) == 0x7fffffff) p.out_action[q] = 0;
    if (tid == 0) {
        p.hdr[0] = 0u;
        if (p.last_launch) { p.hdr[2] = p.n3_total; p.hdr[1] = p.hdr[1] + 1u; }
    }
}

// ---- many queries: a per-dimension bitmap pre-filter ---------------------------------------------------
// With Q queries the plain loop spends ~60 instructions per (warp, query) on the 20-dim test and is issue
// bound long before HBM (Q = 64: 6 % of the HBM roofline).  The filter makes the cost per row independent
// of Q: every observation dim is cut into kSigBins bins over the range the queries' boxes span, and a table
// holds, per (dim, bin), the 64-bit set of queries whose interval [s_lo, s_hi] overlaps the bin.  The bin
// function is monotone in the stored value, so a row inside a query's box always falls into a bin that has
// the query's bit: the AND of the row's 20 table entries is a superset of its matching queries (a handful of
// instructions per dim for all queries at once, the warp leaves as soon as no lane has a candidate left), and
// the few (row, query) pairs that survive take the exact test as before.  Rows outside the queries' range
// clamp into the edge bins (conservative); NaN coordinates land in bin 0 and fail the exact test.
#ifndef DCARL_SIG_BINS
#define DCARL_SIG_BINS 32
#endif
constexpr int kSigBins = DCARL_SIG_BINS, kSigMaxQ = 64;
__host__ __device__ inline size_t filter_smem_bytes(int Q, int D) {
    return (size_t)D * kSigBins * sizeof(unsigned long long) + (size_t)D * sizeof(float2) + (size_t)2 * Q * D;
}
__device__ __forceinline__ int sig_bin(float x, float2 g) {
    const int b = __float2int_rd(fmaf(x, g.x, g.y));       // NaN -> 0, saturating
    return max(0, min(b, kSigBins - 1));
}

template <typename BT, int DT, bool FUSED>
__global__ void __launch_bounds__(kTileRows) k_box_moments(const BoxArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int D = DT > 0 ? DT : p.D;
    const int W = D + 1;
    const int tid = threadIdx.x;
    constexpr int NS = kStages;
    const size_t tile_bytes = (size_t)kTileRows * W * sizeof(BT);
    const size_t ret_bytes = (size_t)kTileRows * sizeof(BT);
    // layout: [stages x state tile][stages x ret tile][lo Q*D][hi Q*D][head][bins][mbarriers]
    BT *st_tile = reinterpret_cast<BT *>(smem_raw);
    BT *rt_tile = reinterpret_cast<BT *>(smem_raw + NS * tile_bytes);
    BT *slo = rt_tile + NS * kTileRows;
    BT *shi = slo + p.Q * D;
    size_t off = NS * (tile_bytes + ret_bytes) + (size_t)2 * p.Q * D * sizeof(BT);
    off = (off + 15) & ~(size_t)15;
    // head thresholds {lo0, hi0, lo1, hi1 | lo2, hi2, lo3, hi3} per query: two 16-byte broadcast loads
    // per (warp, query) instead of eight 4-byte ones (float32 buffers)
    constexpr bool HEAD = DT >= 4 && sizeof(BT) == 4;
    float4 *shead = reinterpret_cast<float4 *>(smem_raw + off);
    off += (size_t)p.Q * 2 * sizeof(float4);
    off = (off + 15) & ~(size_t)15;
    double *sbins = reinterpret_cast<double *>(smem_raw + off);
    const int nbins = p.Q * p.n_actions;
    off += p.bins_in_smem ? (size_t)nbins * 3 * sizeof(double) : 0;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + off);

    if (FUSED) {
        // every CTA derives the thresholds of its queries itself (Q*D of them: trivial next to a tile)
        for (int i = tid; i < p.Q * D; i += kTileRows) box_thresholds<BT>(p.queries[i], p.dist[i % D], slo[i], shi[i]);
    } else {
        const BT *glo = reinterpret_cast<const BT *>(p.lo), *ghi = reinterpret_cast<const BT *>(p.hi);
        for (int i = tid; i < p.Q * D; i += kTileRows) { slo[i] = glo[i]; shi[i] = ghi[i]; }
    }
    __syncthreads();
    if (HEAD)
        for (int q = tid; q < p.Q; q += kTileRows) {
            const BT *l = slo + (size_t)q * D, *h = shi + (size_t)q * D;
            shead[2 * q] = make_float4((float)l[0], (float)h[0], (float)l[1], (float)h[1]);
            shead[2 * q + 1] = make_float4((float)l[2], (float)h[2], (float)l[3], (float)h[3]);
        }
    if (p.bins_in_smem)
        for (int i = tid; i < nbins * 3; i += kTileRows) sbins[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < NS; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const BT *gstate = reinterpret_cast<const BT *>(p.buf_state);
    const BT *gret = reinterpret_cast<const BT *>(p.buf_ret);
    const long long n_tiles = (p.M + kTileRows - 1) / kTileRows;
    const long long n_full = p.M / kTileRows;  // tiles that are complete (TMA-loadable)
    auto issue = [&](long long tile, int stage) {
        // thread 0 only; complete tiles only
        mbar_arrive_expect_tx(&full[stage], (unsigned)(tile_bytes + ret_bytes));
        tma_load_1d(st_tile + (size_t)stage * kTileRows * W, gstate + tile * kTileRows * W,
                    (unsigned)tile_bytes, &full[stage]);
        tma_load_1d(rt_tile + (size_t)stage * kTileRows, gret + tile * kTileRows, (unsigned)ret_bytes,
                    &full[stage]);
    };
    if (tid == 0)
        for (int s = 0; s < NS; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < n_full) issue(t, s);
        }

    double *gacc = FUSED ? box_fused_accumulator(p, tid, kTileRows) : p.out;
    double *bins = p.bins_in_smem ? sbins : gacc;
    const double dist_act = p.dist[D];
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = (int)(it % NS);
        BT *ts = st_tile + (size_t)stage * kTileRows * W;
        BT *tr = rt_tile + (size_t)stage * kTileRows;
        if (tile < n_full) {
            mbar_wait(&full[stage], (unsigned)((it / NS) & 1));
        } else {
            // ragged last tile: plain cooperative loads
            const long long base = tile * kTileRows;
            const int rows = (int)(p.M - base);
            for (int i = tid; i < rows * W; i += kTileRows) ts[i] = gstate[base * W + i];
            if (tid < rows) tr[tid] = gret[base + tid];
            __syncthreads();
        }
        const long long row = tile * kTileRows + tid;
        const bool in_buf = row < p.M;
        const BT *x = ts + (size_t)tid * W;
        BT xr[DT > 0 ? DT : 1];
        if (DT > 0) {
#pragma unroll
            for (int d = 0; d < DT; ++d) xr[d] = x[d];
        }
        // the exact test of query q for this thread's row, and the accumulation of the warp's hits
        auto exact_query = [&](int q, bool pre) {
            const BT *l = slo + q * D, *h = shi + q * D;
            bool ok = pre;
            if (DT > 0) {
                constexpr int kHead = HEAD ? 4 : 0;
                if (HEAD) {
                    const float4 t0 = shead[2 * q], t1 = shead[2 * q + 1];
                    ok = ok & ((float)xr[0] >= t0.x) & ((float)xr[0] <= t0.y) & ((float)xr[1] >= t0.z) & ((float)xr[1] <= t0.w) &
                         ((float)xr[2] >= t1.x) & ((float)xr[2] <= t1.y) & ((float)xr[3] >= t1.z) & ((float)xr[3] <= t1.w);
                    // with many queries most (row, query) pairs die on the first dims: leave before the rest
                    if (p.Q > 1 && !__any_sync(FULL_MASK, ok)) return;
                }
#pragma unroll
                for (int d = kHead; d < DT; ++d) ok = ok && (xr[d] >= l[d]) && (xr[d] <= h[d]);
            } else {
                for (int d = 0; d < D; ++d) { const BT v = x[d]; ok = ok && (v >= l[d]) && (v <= h[d]); }
            }
            if (!__any_sync(FULL_MASK, ok)) return;
            // the stored action picks the candidate(s): s_act - dist <= a <= s_act + dist in float64
            const double sa = ok ? (double)x[D] : 0.0;
            const double alo = sa - dist_act, ahi = sa + dist_act;
            int a0 = 0, a1 = -1;
            if (ok && alo == alo && ahi == ahi) {
                a0 = (int)fmax(ceil(alo), 0.0);
                a1 = (int)fmin(floor(ahi), (double)(p.n_actions - 1));
            }
            const int span = __reduce_max_sync(FULL_MASK, a1 - a0 + 1);
            const double r = ok ? (double)tr[tid] : 0.0;
            for (int e = 0; e < span; ++e) {
                const int a = a0 + e;
                const bool hit = ok && a <= a1 && alo <= (double)a && (double)a <= ahi;
                if (p.bins_in_smem) warp_accumulate<true>(hit, q * p.n_actions + a, r, bins);
                else warp_accumulate<false>(hit, q * p.n_actions + a, r, bins);
            }
        };
        for (int q = 0; q < p.Q; ++q) exact_query(q, in_buf);
        const long long nt = tile + (long long)NS * gridDim.x;
        __syncthreads();  // everyone is done with this stage
        if (tid == 0 && nt < n_full) issue(nt, stage);
    }
    if (p.bins_in_smem) {
        __syncthreads();
        for (int i = tid; i < nbins * 3; i += kTileRows) {
            const double v = sbins[i];
            if (v != 0.0) atomicAdd(gacc + i, v);
        }
    }
    if (FUSED && p.epilogue) box_epilogue(p, gacc, reinterpret_cast<double *>(smem_raw), tid, kTileRows);   // (the tile ring is free now)
}

// ---- many queries (float32 buffer, 20 observation dims): the bitmap-filtered scan ----------------------------
// Work per row is uneven here (a row near a query takes all 20 table lookups and the exact test, most rows leave
// after 8-12 dims), so there is no CTA-wide tile and no CTA barrier in the loop: every WARP streams its own
// sub-tiles of 32 rows through a private ring (1-D TMA bulk copies of 32 x 84 + 32 x 4 bytes, one mbarrier per
// stage, lane 0 issues, the 32 lanes wait), warp w of the grid taking sub-tiles w, w + n_warps, ...
#ifndef DCARL_FILT_WARPS
#define DCARL_FILT_WARPS 16
#endif
#ifndef DCARL_FILT_STAGES
#define DCARL_FILT_STAGES 2
#endif
constexpr int kFiltWarps = DCARL_FILT_WARPS, kFiltStages = DCARL_FILT_STAGES, kFiltD = 20, kFiltW = kFiltD + 1;
template <typename BT> constexpr int filt_stage_bytes() { return 32 * kFiltW * (int)sizeof(BT) + 32 * (int)sizeof(BT); }   // float32: 2816 (state rows, then the returns)

__host__ __device__ inline size_t box_filter_smem_bytes(int Q, int n_actions, bool bins_in_smem, size_t elem) {
    size_t off = (size_t)kFiltWarps * kFiltStages * 32 * (kFiltW + 1) * elem + (size_t)2 * Q * kFiltD * elem;
    off = (off + 15) & ~(size_t)15;
    off += filter_smem_bytes(Q, kFiltD);
    off = (off + 15) & ~(size_t)15;
    if (bins_in_smem) off += (size_t)Q * n_actions * 3 * sizeof(double);
    off += (size_t)kFiltWarps * kFiltStages * sizeof(uint64_t);
    return off;
}

template <typename BT, bool FUSED>
__global__ void __launch_bounds__(kFiltWarps * 32) k_box_filter(const BoxArgs p) {
    constexpr int kFiltStageBytes = filt_stage_bytes<BT>();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int D = kFiltD, W = kFiltW, NT = kFiltWarps * 32;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // layout: [warps x stages x {32 rows, 32 returns}][lo Q*D][hi Q*D][tables][bins][mbarriers]
    unsigned char *ring = smem_raw + (size_t)warp * kFiltStages * kFiltStageBytes;
    BT *slo = reinterpret_cast<BT *>(smem_raw + (size_t)kFiltWarps * kFiltStages * kFiltStageBytes);
    BT *shi = slo + p.Q * D;
    size_t off = (size_t)kFiltWarps * kFiltStages * kFiltStageBytes + (size_t)2 * p.Q * D * sizeof(BT);
    off = (off + 15) & ~(size_t)15;
    // per (dim, bin) the set of queries whose interval overlaps the bin; per dim {scale, offset} of the bin
    // function; scratch: the first / last bin of every (query, dim) interval
    unsigned long long *stab = reinterpret_cast<unsigned long long *>(smem_raw + off);
    float2 *sgeo = reinterpret_cast<float2 *>(stab + D * kSigBins);
    unsigned char *sb01 = reinterpret_cast<unsigned char *>(sgeo + D);
    off += filter_smem_bytes(p.Q, D);
    off = (off + 15) & ~(size_t)15;
    double *sbins = reinterpret_cast<double *>(smem_raw + off);
    off += p.bins_in_smem ? (size_t)p.Q * p.n_actions * 3 * sizeof(double) : 0;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + off) + warp * kFiltStages;

    if (lane == 0) {
        for (int s = 0; s < kFiltStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    const BT *gstate = reinterpret_cast<const BT *>(p.buf_state);
    const BT *gret = reinterpret_cast<const BT *>(p.buf_ret);
    const long long n_sub = p.M / 32;                       // complete sub-tiles (TMA-loadable)
    // One chunk of queries for the whole grid (the warps of all CTAs share the buffer's sub-tiles), or — many
    // chunks on a small buffer — whole chunks per CTA (its 16 warps share all sub-tiles, the buffer comes from L2).
    const int n_chunks = FUSED ? 1 : p.n_chunks;
    const bool per_cta = n_chunks > 1;
    const long long n_warps = per_cta ? kFiltWarps : (long long)gridDim.x * kFiltWarps;
    const long long w0 = per_cta ? warp : (long long)blockIdx.x * kFiltWarps + warp;
    auto issue = [&](long long sub, int stage) {            // lane 0 only
        unsigned char *dst = ring + (size_t)stage * kFiltStageBytes;
        mbar_arrive_expect_tx(&full[stage], (unsigned)kFiltStageBytes);
        tma_load_1d(dst, gstate + sub * 32 * W, 32 * W * (int)sizeof(BT), &full[stage]);
        tma_load_1d(dst + 32 * W * sizeof(BT), gret + sub * 32, 32 * (int)sizeof(BT), &full[stage]);
    };
    double *gacc = nullptr;
    long long it = 0;                                       // sub-tiles this warp has taken so far (all chunks): ring position
    for (int chunk = per_cta ? (int)blockIdx.x : 0; chunk < n_chunks; chunk += per_cta ? (int)gridDim.x : 1) {
    const int Qn = per_cta ? min(p.Q, p.q_total - chunk * p.Q) : p.Q;
    const int nbins = Qn * p.n_actions;
    if (FUSED) {
        for (int i = tid; i < Qn * D; i += NT) box_thresholds<BT>(p.queries[i], p.dist[i % D], slo[i], shi[i]);
    } else {
        const BT *glo = reinterpret_cast<const BT *>(p.lo) + (size_t)chunk * p.Q * D;
        const BT *ghi = reinterpret_cast<const BT *>(p.hi) + (size_t)chunk * p.Q * D;
        for (int i = tid; i < Qn * D; i += NT) { slo[i] = glo[i]; shi[i] = ghi[i]; }
    }
    __syncthreads();
    // this warp's first sub-tiles are on their way while the CTA builds the tables
    if (lane == 0)
        for (int s = 0; s < kFiltStages; ++s) {
            const long long t = w0 + (long long)s * n_warps;
            if (t < n_sub) issue(t, (int)((it + s) % kFiltStages));
        }
    // bin geometry: the range of the finite thresholds of all queries, per dim
    if (tid < D) {
        float gmin = CUDART_INF_F, gmax = -CUDART_INF_F;
        for (int q = 0; q < Qn; ++q) {
            const float l = narrow_dn(slo[q * D + tid]), h = narrow_up(shi[q * D + tid]);
            if (l <= h) { gmin = fminf(gmin, l); gmax = fmaxf(gmax, h); }
        }
        float scale = 0.f;
        if (gmin < gmax && gmax - gmin < CUDART_INF_F) scale = (float)kSigBins / (gmax - gmin);
        if (!(gmin <= gmax)) { gmin = 0.f; scale = 0.f; }                    // no query has an interval in this dim
        sgeo[tid] = make_float2(scale, -gmin * scale);                        // bin coordinate t = fma(x, scale, offset)
    }
    __syncthreads();
    // first / last bin of every interval (the same bin function as the rows': monotone in the stored value,
    // so a value inside [l, h] lands in [bin(l), bin(h)]); an empty interval gets first > last
    for (int i = tid; i < Qn * D; i += NT) {
        const float l = narrow_dn(slo[i]), h = narrow_up(shi[i]);   // (float64 thresholds: rounded outwards)
        const float2 g = sgeo[i % D];
        const bool some = slo[i] <= shi[i];
        sb01[2 * i] = (unsigned char)(some ? sig_bin(l, g) : 1);
        sb01[2 * i + 1] = (unsigned char)(some ? sig_bin(h, g) : 0);
    }
    __syncthreads();
    for (int e = tid; e < D * kSigBins; e += NT) {
        const int d = e / kSigBins, b = e % kSigBins;
        unsigned long long m = 0ull;
        for (int q = 0; q < Qn; ++q) {
            const int i = q * D + d;
            m |= (unsigned long long)((int)sb01[2 * i] <= b && b <= (int)sb01[2 * i + 1]) << q;
        }
        stab[e] = m;
    }
    if (p.bins_in_smem)
        for (int i = tid; i < nbins * 3; i += NT) sbins[i] = 0.0;
    __syncthreads();

    gacc = FUSED ? box_fused_accumulator(p, tid, NT) : p.out + (size_t)chunk * p.Q * p.n_actions * 3;
    double *bins = p.bins_in_smem ? sbins : gacc;
    const double dist_act = p.dist[D];
    // One row per lane: x = its 20 observation values + the action (shared or global memory), r = its return.
    // 32 rows, one per lane: rows0 = the first row's 20 observation values + action (shared or global memory, pitch W),
    // rets0 = its return.  TAIL: lanes past the end of the buffer are not valid and must not load.
    auto scan_rows = [&](auto tail_tag, const BT *rows0, const BT *rets0, bool valid) {
        constexpr bool TAIL = decltype(tail_tag)::value;
        const BT *x = rows0 + lane * W;
        // candidates: the AND of the row's table entries, all queries at once, four dims per round; the warp
        // leaves as soon as no lane has a candidate left (most rows are out after two or three rounds)
        unsigned long long surv = valid ? ~0ull : 0ull;
        bool alive = true;                                  // warp-uniform
#pragma unroll
        for (int g = 0; g < D / 4; ++g) {
            if (alive) {
                unsigned long long m[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    const int d = 4 * g + u;
                    // (a float64 value is rounded to nearest for its bin: monotone, and inside the outward-rounded interval
                    //  of every query whose exact interval holds it)
                    const float xv = (float)(TAIL ? (valid ? x[d] : BT(0)) : x[d]);
                    m[u] = stab[d * kSigBins + sig_bin(xv, sgeo[d])];
                }
                surv &= (m[0] & m[1]) & (m[2] & m[3]);
                alive = __any_sync(FULL_MASK, surv != 0ull);
            }
        }
        if (!alive) return;
        // The few surviving (row, query) pairs take the exact test, the WARP on one pair at a time with one
        // observation dim per lane (a lane walking its own pairs through 20 dims kept the other 31 lanes idle:
        // that loop was 30 % of the kernel's instructions).  Hits are rare with many queries: lane 0 adds them
        // to the bins one atomic triple at a time.
        unsigned have = __ballot_sync(FULL_MASK, surv != 0ull);
        while (have) {                                      // warp-uniform loops
            const int src = __ffs(have) - 1;
            have &= have - 1u;
            unsigned long long sv = __shfl_sync(FULL_MASK, surv, src);
            const BT *xs = rows0 + src * W;
            const BT xd = lane < D ? xs[lane] : BT(0);
            while (sv) {
                const int q = __ffsll((long long)sv) - 1;
                sv &= sv - 1ull;
                bool in = true;
                if (lane < D) in = (xd >= slo[q * D + lane]) && (xd <= shi[q * D + lane]);
                if (!__all_sync(FULL_MASK, in) || lane != 0) continue;
                // the stored action picks the candidate(s): s_act - dist <= a <= s_act + dist in float64
                const double sa = (double)xs[D];
                const double alo = sa - dist_act, ahi = sa + dist_act;
                if (!(alo == alo) || !(ahi == ahi)) continue;
                const int a0 = (int)fmax(ceil(alo), 0.0), a1 = (int)fmin(floor(ahi), (double)(p.n_actions - 1));
                const double rr = (double)rets0[src];
                for (int a = a0; a <= a1; ++a)
                    if (alo <= (double)a && (double)a <= ahi) {
                        double *b = bins + ((size_t)q * p.n_actions + a) * 3;
                        atomicAdd(b + 0, 1.0);
                        atomicAdd(b + 1, rr);
                        atomicAdd(b + 2, rr * rr);
                    }
            }
            __syncwarp();
        }
    };

    for (long long sub = w0; sub < n_sub; sub += n_warps, ++it) {
        const int stage = (int)(it % kFiltStages);
        const BT *ts = reinterpret_cast<const BT *>(ring + (size_t)stage * kFiltStageBytes);
        mbar_wait(&full[stage], (unsigned)((it / kFiltStages) & 1));
        scan_rows(std::false_type{}, ts, ts + 32 * W, true);
        __syncwarp();                                       // every lane is done with the stage
        const long long nt = sub + (long long)kFiltStages * n_warps;
        if (lane == 0 && nt < n_sub) issue(nt, stage);
    }
    // the last M % 32 rows, straight from global memory, by the warp whose turn it would be
    if ((p.M & 31) != 0 && w0 == n_sub % n_warps)
        scan_rows(std::true_type{}, gstate + n_sub * 32 * W, gret + n_sub * 32, n_sub * 32 + lane < p.M);
    if (p.bins_in_smem) {
        __syncthreads();
        for (int i = tid; i < nbins * 3; i += NT) {
            const double v = sbins[i];
            if (v != 0.0) atomicAdd(gacc + i, v);
        }
    }
    __syncthreads();                                        // (many chunks: the tables are rebuilt for the next one)
    }   // chunk
    if (FUSED && p.epilogue) box_epilogue(p, gacc, reinterpret_cast<double *>(smem_raw), tid, NT);   // (the rings are free now)
}

// ---- many queries: the sum over ranks and the select as a second launch -----------------------------------
// The fused epilogue is ONE CTA: fine for the 6 KB of a single query, not for 64 x 256 x 3 doubles — alone it
// spent 40 us on the (mean, var, pick) of 16 384 pairs (a quarter of the fused Q = 64 pass), and on row shards it
// would push all of them to every peer (measured at 2 ranks: the pass took as long as on one GPU).
// Here one CTA per query runs the query's select, after — on row shards — exchanging and summing that query's
// na x 3 totals through the peer workspace of
// peer.cuh — same slots, one flag per rank: the last CTA of a rank to have written its slice raises the
// rank's flag on every peer, every CTA waits for all flags in its own workspace — and runs the query's select.
// At most 64 CTAs: all resident, so the CTAs of a rank never wait for one that cannot start.
__global__ void __launch_bounds__(256) k_box_reduce_select(const BoxArgs p) {
    const int tid = threadIdx.x, q = blockIdx.x, na = p.n_actions, n3q = na * 3;
    const PeerCtx &c = p.peers;
    const bool solo = c.world <= 1;
    __shared__ unsigned s_e, s_last, s_lost;
    __shared__ double s_rule[3];
    const unsigned calls = p.hdr[1];
    const double *acc = p.acc[calls & 1u] + p.q0_off + (size_t)q * n3q;
    unsigned e = 0;
    size_t par = 0;
    if (tid == 0) { s_lost = 0u; s_e = solo ? 0u : *c.epoch + 1u; }   // (the epoch is bumped by the last CTA to finish, below)
    __syncthreads();
    if (!solo) {
        e = s_e;
        par = e & 1u;
        // my query's totals into slot [par][my rank] of every rank
        const size_t slot_off = (par * c.world + c.rank) * (size_t)c.n_max + (size_t)q * n3q;
        for (int r = 0; r < c.world; ++r) {
            double *d = c.slots[r] + slot_off;
            for (int i = tid; i < n3q; i += blockDim.x) d[i] = __ldcg(acc + i);
        }
        __threadfence_system();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(p.hdr + 4, 1u) == gridDim.x - 1u;
        __syncthreads();
        if (s_last) {                                     // every slice of this rank is written: tell the peers
            __threadfence_system();
            if (tid < c.world) st_release_sys(c.flags[tid] + par * c.world + c.rank, e);
        }
        if (tid < c.world && !peer_wait_flag(c.flags[c.rank] + par * c.world + tid, e)) s_lost = 1u;
        __syncthreads();
    }
    const bool lost = !solo && s_lost != 0u;              // a peer never arrived: NaN totals (see peer_wait_flag)
    // the totals (on row shards the rank-ordered sum: bit-identical on every rank) -> out_moments
    const double *mine = solo ? nullptr : c.slots[c.rank] + par * c.world * (size_t)c.n_max + (size_t)q * n3q;
    double *tot = p.out_moments + (size_t)q * n3q;
    for (int i = tid; i < n3q; i += blockDim.x) {
        double sum = 0.0;
        if (solo) sum = __ldcg(acc + i);
        else for (int r = 0; r < c.world; ++r) sum += ld_relaxed_sys(mine + (size_t)r * c.n_max + i);
        if (lost) sum = __longlong_as_double(0x7ff8000000000000ll);
        tot[i] = sum;
        if (i < 3) s_rule[i] = sum;
    }
    int *pick = p.out_action ? p.out_action + q : nullptr;
    if (tid == 0 && pick) *pick = 0x7fffffff;
    __syncthreads();
    const double n0 = s_rule[0];
    const double mean0 = n0 > 0.0 ? s_rule[1] / n0 : -1.0;
    const double var0 = n0 > 0.0 ? fmax(s_rule[2] / n0 - mean0 * mean0, 0.0) : -1.0;
    for (int a = tid; a < na; a += blockDim.x) {
        const double n = tot[a * 3];
        double mean = -1.0, var = -1.0;                                        // RLS.py:169-170
        if (n > 0.0) { mean = tot[a * 3 + 1] / n; var = fmax(tot[a * 3 + 2] / n - mean * mean, 0.0); }
        if (p.out_mv) { p.out_mv[((size_t)q * na + a) * 2] = mean; p.out_mv[((size_t)q * na + a) * 2 + 1] = var; }
        if (!pick || a == 0 || n0 < (double)p.rp.visited_times_thres || n < (double)p.rp.min_candidate_visits ||
            mean0 > p.rp.rule_mean_gate)
            continue;                                                          // RLS.py:141
        const double z = (mean - mean0) / sqrt(var0 / n0 + var / n);         // RLS.py:144-148
        if (0.5 * erfc(-z * 0.70710678118654752440) > p.rp.confidence_thres) atomicMin(pick, a);   // RLS.py:150-154
    }
    __syncthreads();
    if (tid == 0) {
        if (pick && __ldcg(pick) == 0x7fffffff) *pick = 0;
        if (atomicAdd(p.hdr + 5, 1u) == gridDim.x - 1u) {   // the last CTA of the launch
            if (!solo) *c.epoch = e;
            p.hdr[4] = 0u; p.hdr[5] = 0u;
            if (p.last_launch) { p.hdr[2] = p.n3_total; p.hdr[1] = calls + 1u; }
        }
    }
}

// ---- per-query row ranges (the sorted index of SURVEY row f1): one warp per query ------------------------------
// With an index every query only has to look at the run of rows that can match it in the sort dimension (plus
// the unsorted tail of later appends).  One launch answers all queries: a warp takes one query, keeps the query's
// thresholds and its own n_actions x 3 bins in shared memory, stages 32 rows at a time (coalesced loads of the
// contiguous AoS rows, then one row per lane at pitch D + 1) and applies the exact test; its bins go to
// out_moments with plain stores (nothing to zero, no atomics between queries).  Small buffers stay in L2.
struct RangeArgs {
    const void *buf_state, *buf_ret;
    const double *dist, *queries;
    const long long *row_lo, *row_hi;      // [Q]
    long long tail_lo, tail_hi;            // a second range every query scans (the unsorted tail), may be empty
    long long M;
    double *out;                           // [Q, n_actions, 3]
    int D, Q, n_actions;
};
constexpr int kRangeWarps = 8;

template <typename BT>
__global__ void __launch_bounds__(kRangeWarps * 32) k_box_ranges(const RangeArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int D = p.D, W = D + 1, na = p.n_actions;
    // per warp: bins [na * 3] doubles, thresholds lo / hi [D], tile [32 * W]
    const size_t per_warp = (((size_t)na * 3 * sizeof(double) + (size_t)(2 * D + 32 * W) * sizeof(BT)) + 15) & ~(size_t)15;
    unsigned char *base = smem_raw + (size_t)warp * per_warp;
    double *bins = reinterpret_cast<double *>(base);
    BT *lo = reinterpret_cast<BT *>(bins + (size_t)na * 3);
    BT *hi = lo + D;
    BT *tile = hi + D;
    const BT *gstate = reinterpret_cast<const BT *>(p.buf_state);
    const BT *gret = reinterpret_cast<const BT *>(p.buf_ret);
    const double dist_act = p.dist[D];
    for (int q = blockIdx.x * kRangeWarps + warp; q < p.Q; q += gridDim.x * kRangeWarps) {
        for (int i = lane; i < na * 3; i += 32) bins[i] = 0.0;
        for (int d = lane; d < D; d += 32) box_thresholds<BT>(p.queries[(size_t)q * D + d], p.dist[d], lo[d], hi[d]);
        __syncwarp();
        for (int part = 0; part < 2; ++part) {
            // (ranges are clamped to the buffer: a wrong range may cost matches, never an out-of-bounds read)
            const long long r_lo = max(0ll, part ? p.tail_lo : p.row_lo[q]), r_hi = min(p.M, part ? p.tail_hi : p.row_hi[q]);
            for (long long r0 = r_lo; r0 < r_hi; r0 += 32) {
                const int rows = (int)min((long long)32, r_hi - r0);
                const BT *src = gstate + r0 * W;
                for (int i = lane; i < rows * W; i += 32) tile[i] = src[i];
                __syncwarp();
                bool ok = lane < rows;
                const BT *x = tile + lane * W;
                if (ok)
                    for (int d = 0; d < D; ++d) { const BT v = x[d]; ok = ok && (v >= lo[d]) && (v <= hi[d]); }
                if (__any_sync(FULL_MASK, ok)) {
                    // the stored action picks the candidate(s): s_act - dist <= a <= s_act + dist in float64
                    const double sa = ok ? (double)x[D] : 0.0;
                    const double alo = sa - dist_act, ahi = sa + dist_act;
                    int a0 = 0, a1 = -1;
                    if (ok && alo == alo && ahi == ahi) {
                        a0 = (int)fmax(ceil(alo), 0.0);
                        a1 = (int)fmin(floor(ahi), (double)(na - 1));
                    }
                    const int span = __reduce_max_sync(FULL_MASK, a1 - a0 + 1);
                    const double rr = ok ? (double)gret[r0 + lane] : 0.0;
                    for (int e = 0; e < span; ++e) {
                        const int a = a0 + e;
                        const bool hit = ok && a <= a1 && alo <= (double)a && (double)a <= ahi;
                        warp_accumulate<true>(hit, a, rr, bins);
                    }
                }
                __syncwarp();                                   // the tile is free again
            }
        }
        __syncwarp();
        double *o = p.out + (size_t)q * na * 3;
        for (int i = lane; i < na * 3; i += 32) o[i] = bins[i];
        __syncwarp();
    }
}

// moments -> mean/var (+ sentinel) and the act_test pick, one thread per query
__global__ void k_box_select(const double *__restrict__ mom, int Q, int n_actions, dcarl_rls_params rp,
                             double *__restrict__ out_mv, int *__restrict__ out_action) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    box_select_one(mom + (size_t)q * n_actions * 3, n_actions, rp,
                   out_mv ? out_mv + (size_t)q * n_actions * 2 : nullptr, out_action ? out_action + q : nullptr);
}

static size_t box_smem_bytes(int Q, int D, int n_actions, bool bins_in_smem, size_t elem = sizeof(float)) {
    const int W = D + 1;
    size_t off = (size_t)kStages * kTileRows * (W + 1) * elem + (size_t)2 * Q * D * elem;
    off = (off + 15) & ~(size_t)15;
    off += (size_t)Q * 32;   // head thresholds
    off = (off + 15) & ~(size_t)15;
    if (bins_in_smem) off += (size_t)Q * n_actions * 3 * sizeof(double);
    off += kStages * sizeof(uint64_t);
    return off;
}

constexpr size_t kSmemBudget = 200 * 1024;
constexpr size_t kSoloEpilogueMax = 4096;   // doubles one CTA sums over the ranks by itself (fused epilogue)

struct ScanPlan {
    int qc;             // queries per launch
    bool filt;          // launches of >= 8 queries take the bitmap-filtered kernel
    size_t elem;
};
// Queries are processed in chunks whose thresholds (+ bins) fit beside the tile ring; many queries on a
// float32 buffer of 20 observation dims go through the bitmap filter, at most kSigMaxQ per launch.
static int plan_scan(int Q, int D, int n_actions, int buf_is_f64, ScanPlan &sp) {
    sp.elem = buf_is_f64 ? sizeof(double) : sizeof(float);
    sp.filt = D == kFiltD && Q >= 8;
    sp.qc = Q;
    if (sp.filt) {
        if (sp.qc > kSigMaxQ) sp.qc = kSigMaxQ;
        return 0;                                           // 64 queries always fit (box_filter_smem_bytes: 108 KB float32, 209 KB float64)
    }
    while (sp.qc > 1 && box_smem_bytes(sp.qc, D, n_actions, false, sp.elem) > kSmemBudget) sp.qc = (sp.qc + 1) / 2;
    if (box_smem_bytes(sp.qc, D, n_actions, false, sp.elem) > kSmemBudget) return bad_arg(4, "D too large for the tile ring");
    return 0;
}

// One scan launch of qn queries (a.Q etc. filled by the caller; unfused: thresholds from the workspace; fused:
// computed in the kernel, epilogue on).  Persistent grid: CTAs per SM by shared memory.
template <typename BT, bool FUSED>
static int launch_scan(BoxArgs &a, const ScanPlan &sp, int sms, cudaStream_t st) {
    const int qn = a.Q, D = a.D, na = a.n_actions;
    auto grid_of = [&](size_t smem, long long units) {
        int per_sm = (int)((size_t)(228 * 1024) / (smem + 1024));   // 228 KB per SM, 1 KB of it reserved per resident CTA
        per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
        long long grid = (long long)sms * per_sm;
        if (grid > units) grid = units;
        return (unsigned)(grid < 1 ? 1 : grid);             // (an empty shard still takes part in the reduction)
    };
    if (sp.filt && qn >= 8) {
        a.bins_in_smem = (qn * na <= kMaxSmemBins) && box_filter_smem_bytes(qn, na, true, sizeof(BT)) <= kSmemBudget;
        const size_t smem = box_filter_smem_bytes(qn, na, a.bins_in_smem, sizeof(BT));
        const unsigned grid = grid_of(smem, (a.M / 32 + kFiltWarps - 1) / kFiltWarps);
        DCARL_CUDA(cudaFuncSetAttribute(k_box_filter<BT, FUSED>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_box_filter<BT, FUSED><<<grid, kFiltWarps * 32, smem, st>>>(a);
        DCARL_LAUNCH_CHECK("k_box_filter");
        return 0;
    }
    a.bins_in_smem = (qn * na <= kMaxSmemBins) && box_smem_bytes(qn, D, na, true, sp.elem) <= kSmemBudget;
    const size_t smem = box_smem_bytes(qn, D, na, a.bins_in_smem, sp.elem);
    const unsigned grid = grid_of(smem, (a.M + kTileRows - 1) / kTileRows);
    auto go = [&](auto kern) -> int {
        DCARL_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kern<<<grid, kTileRows, smem, st>>>(a);
        DCARL_LAUNCH_CHECK("k_box_moments");
        return 0;
    };
    if (D == 20) return go(k_box_moments<BT, 20, FUSED>);
    return go(k_box_moments<BT, 0, FUSED>);
}

}  // namespace dcarl

using namespace dcarl;

extern "C" size_t dcarl_box_moments_ws(int32_t Q, int32_t D, int32_t n_actions) {
    (void)n_actions;
    if (Q <= 0 || D <= 0) return 0;
    return (size_t)2 * Q * D * sizeof(double);   // thresholds in the widest element type
}

extern "C" int dcarl_box_moments(const void *buf_state, const void *buf_ret, int buf_is_f64, int64_t M, int32_t D,
                                 const double *dist, const double *queries, int32_t Q, int32_t n_actions,
                                 double *out_moments, void *ws, size_t ws_bytes, dcarl_stream_t stream) {
    if (M < 0) return bad_arg(4, "M < 0");
    if (D <= 0 || D > 255) return bad_arg(5, "D must be in [1, 255]");
    if (Q < 0) return bad_arg(8, "Q < 0");
    if (n_actions <= 0) return bad_arg(9, "n_actions <= 0");
    if (Q == 0) return 0;
    if (!dist || !queries) return bad_arg(6, "dist/queries is null");
    if (!out_moments) return bad_arg(10, "out_moments is null");
    if (M > 0 && (!buf_state || !buf_ret)) return bad_arg(1, "buffer is null");
    if (((uintptr_t)buf_state & 15) || ((uintptr_t)buf_ret & 15)) return bad_arg(1, "buffer must be 16-byte aligned");
    if (!ws || ws_bytes < dcarl_box_moments_ws(Q, D, n_actions) || ((uintptr_t)ws & 7))
        return bad_arg(11, "workspace too small or misaligned (see dcarl_box_moments_ws)");
    cudaStream_t st = (cudaStream_t)stream;
    DCARL_CUDA(cudaMemsetAsync(out_moments, 0, (size_t)Q * n_actions * 3 * sizeof(double), st));
    if (M == 0) return 0;
    ScanPlan sp;
    if (int rc = plan_scan(Q, D, n_actions, buf_is_f64, sp)) return rc;

    const size_t nthr = (size_t)Q * D;
    void *lo = ws, *hi = reinterpret_cast<unsigned char *>(ws) + nthr * sp.elem;
    if (buf_is_f64) k_box_prepare<double><<<(Q * D + 127) / 128, 128, 0, st>>>(queries, dist, Q, D, (double *)lo, (double *)hi);
    else k_box_prepare<float><<<(Q * D + 127) / 128, 128, 0, st>>>(queries, dist, Q, D, (float *)lo, (float *)hi);
    DCARL_LAUNCH_CHECK("k_box_prepare");

    const int sms = sm_count();
    // Many chunks of queries on a buffer that stays in L2: one launch, whole chunks per CTA (a launch per chunk costs
    // ~40 us of prologue and ramp for a scan of a few microseconds: 512 chunks on 3.6e5 rows took 20 ms)
    const int n_chunks = (Q + sp.qc - 1) / sp.qc;
    if (sp.filt && n_chunks > 1 && (size_t)M * (D + 2) * sp.elem <= ((size_t)96 << 20)) {
        BoxArgs a;
        a.buf_state = buf_state; a.buf_ret = buf_ret; a.M = M; a.D = D; a.Q = sp.qc; a.n_actions = n_actions;
        a.lo = lo; a.hi = hi; a.dist = dist; a.out = out_moments;
        a.queries = nullptr; a.hdr = nullptr; a.acc[0] = a.acc[1] = nullptr; a.q0_off = 0; a.n3_total = 0; a.last_launch = 0;
        a.epilogue = 0; a.n_chunks = n_chunks; a.q_total = Q;
        a.out_moments = nullptr; a.out_mv = nullptr; a.out_action = nullptr;
        a.rp = dcarl_rls_params{}; a.peers = PeerCtx{};
        a.bins_in_smem = (sp.qc * n_actions <= kMaxSmemBins) && box_filter_smem_bytes(sp.qc, n_actions, true, sp.elem) <= kSmemBudget;
        const size_t smem = box_filter_smem_bytes(sp.qc, n_actions, a.bins_in_smem, sp.elem);
        int per_sm = (int)((size_t)(228 * 1024) / (smem + 1024));
        per_sm = per_sm < 1 ? 1 : (per_sm > 4 ? 4 : per_sm);
        const unsigned grid = (unsigned)(n_chunks < sms * per_sm ? n_chunks : sms * per_sm);
        if (buf_is_f64) {
            DCARL_CUDA(cudaFuncSetAttribute(k_box_filter<double, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_box_filter<double, false><<<grid, kFiltWarps * 32, smem, st>>>(a);
        } else {
            DCARL_CUDA(cudaFuncSetAttribute(k_box_filter<float, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_box_filter<float, false><<<grid, kFiltWarps * 32, smem, st>>>(a);
        }
        DCARL_LAUNCH_CHECK("k_box_filter");
        return 0;
    }
    for (int q0 = 0; q0 < Q; q0 += sp.qc) {
        const int qn = (Q - q0 < sp.qc) ? Q - q0 : sp.qc;
        BoxArgs a;
        a.buf_state = buf_state; a.buf_ret = buf_ret; a.M = M; a.D = D; a.Q = qn; a.n_actions = n_actions;
        a.n_chunks = 1; a.q_total = qn;
        a.lo = reinterpret_cast<unsigned char *>(lo) + (size_t)q0 * D * sp.elem;
        a.hi = reinterpret_cast<unsigned char *>(hi) + (size_t)q0 * D * sp.elem;
        a.dist = dist;
        a.out = out_moments + (size_t)q0 * n_actions * 3;
        a.queries = nullptr; a.hdr = nullptr; a.acc[0] = a.acc[1] = nullptr; a.q0_off = 0; a.n3_total = 0; a.last_launch = 0;
        a.epilogue = 0;
        a.out_moments = nullptr; a.out_mv = nullptr; a.out_action = nullptr;
        a.rp = dcarl_rls_params{}; a.peers = PeerCtx{};
        const int rc = buf_is_f64 ? launch_scan<double, false>(a, sp, sms, st) : launch_scan<float, false>(a, sp, sms, st);
        if (rc) return rc;
    }
    return 0;
}

extern "C" size_t dcarl_box_query_ws(int32_t Q, int32_t D, int32_t n_actions) {
    (void)D;
    if (Q <= 0 || n_actions <= 0) return 0;
    const size_t half = ((size_t)Q * n_actions * 3 * sizeof(double) + 15) & ~(size_t)15;
    return 2 * half + 48;   // 32-byte header + two accumulators (+ slack for the split of any ws_bytes)
}

extern "C" int dcarl_box_query(const void *buf_state, const void *buf_ret, int buf_is_f64, int64_t M, int32_t D,
                               const double *dist, const double *queries, int32_t Q, int32_t n_actions,
                               const dcarl_rls_params *rp, double *out_moments, double *out_mean_var,
                               int32_t *out_action, void *ws, size_t ws_bytes, const dcarl_peer_ctx *peers,
                               dcarl_stream_t stream) {
    if (M < 0) return bad_arg(4, "M < 0");
    if (D <= 0 || D > 255) return bad_arg(5, "D must be in [1, 255]");
    if (Q < 0) return bad_arg(8, "Q < 0");
    if (n_actions <= 0) return bad_arg(9, "n_actions <= 0");
    if (!rp) return bad_arg(10, "params is null");
    if (Q == 0) return 0;
    if (!dist || !queries) return bad_arg(6, "dist/queries is null");
    if (!out_moments) return bad_arg(11, "out_moments is null");
    if (M > 0 && (!buf_state || !buf_ret)) return bad_arg(1, "buffer is null");
    if (((uintptr_t)buf_state & 15) || ((uintptr_t)buf_ret & 15)) return bad_arg(1, "buffer must be 16-byte aligned");
    if (!ws || ws_bytes < dcarl_box_query_ws(Q, D, n_actions) || ((uintptr_t)ws & 15))
        return bad_arg(14, "workspace too small or misaligned (see dcarl_box_query_ws)");
    cudaStream_t st = (cudaStream_t)stream;
    if ((size_t)Q * n_actions * 3 > 0x7fffffffu) return bad_arg(8, "Q * n_actions too large");
    // (the halves depend on ws_bytes only: one workspace may serve calls of different shapes)
    unsigned *hdr = reinterpret_cast<unsigned *>(ws);
    const size_t half = ((ws_bytes - 32) / 2) & ~(size_t)15;
    double *acc0 = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(ws) + 32);
    double *acc1 = reinterpret_cast<double *>(reinterpret_cast<unsigned char *>(ws) + 32 + half);
    ScanPlan sp;
    if (int rc = plan_scan(Q, D, n_actions, buf_is_f64, sp)) return rc;
    PeerCtx pc;
    {
        const int nmsg = ((Q < sp.qc) ? Q : sp.qc) * n_actions * 3;
        if (int rc = make_peer_ctx(peers, nmsg, 16, pc)) return rc;
    }
    const int sms = sm_count();
    for (int q0 = 0; q0 < Q; q0 += sp.qc) {
        const int qn = (Q - q0 < sp.qc) ? Q - q0 : sp.qc;
        BoxArgs a;
        a.buf_state = buf_state; a.buf_ret = buf_ret; a.M = M; a.D = D; a.Q = qn; a.n_actions = n_actions;
        a.lo = nullptr; a.hi = nullptr; a.dist = dist;
        a.out = nullptr;
        a.queries = queries + (size_t)q0 * D;
        a.hdr = hdr; a.acc[0] = acc0; a.acc[1] = acc1;
        a.q0_off = (unsigned)((size_t)q0 * n_actions * 3);
        a.n3_total = (unsigned)((size_t)Q * n_actions * 3);
        a.last_launch = q0 + sp.qc >= Q;
        a.n_chunks = 1; a.q_total = qn;
        // many totals: the scan only accumulates, a second launch (one CTA per query) sums over the ranks and selects
        const bool split = (size_t)qn * n_actions * 3 > kSoloEpilogueMax && qn <= 64;
        a.epilogue = split ? 0 : 1;
        a.out_moments = out_moments + (size_t)q0 * n_actions * 3;
        a.out_mv = out_mean_var ? out_mean_var + (size_t)q0 * n_actions * 2 : nullptr;
        a.out_action = out_action ? out_action + q0 : nullptr;
        a.rp = *rp;
        a.peers = pc;
        const int rc = buf_is_f64 ? launch_scan<double, true>(a, sp, sms, st) : launch_scan<float, true>(a, sp, sms, st);
        if (rc) return rc;
        if (split) {
            k_box_reduce_select<<<qn, 256, 0, st>>>(a);
            DCARL_LAUNCH_CHECK("k_box_reduce_select");
        }
    }
    return 0;
}

extern "C" int dcarl_box_select(const double *moments, int32_t Q, int32_t n_actions, const dcarl_rls_params *p,
                                double *out_mean_var, int32_t *out_action, dcarl_stream_t stream) {
    if (Q < 0) return bad_arg(2, "Q < 0");
    if (n_actions <= 0) return bad_arg(3, "n_actions <= 0");
    if (!p) return bad_arg(4, "params is null");
    if (Q == 0) return 0;
    if (!moments) return bad_arg(1, "moments is null");
    k_box_select<<<(Q + 127) / 128, 128, 0, (cudaStream_t)stream>>>(moments, Q, n_actions, *p, out_mean_var, out_action);
    DCARL_LAUNCH_CHECK("k_box_select");
    return 0;
}

extern "C" int dcarl_box_moments_ranges(const void *buf_state, const void *buf_ret, int buf_is_f64, int64_t M, int32_t D,
                                        const double *dist, const double *queries, const int64_t *row_lo,
                                        const int64_t *row_hi, int64_t tail_lo, int64_t tail_hi, int32_t Q,
                                        int32_t n_actions, double *out_moments, dcarl_stream_t stream) {
    if (M < 0) return bad_arg(4, "M < 0");
    if (D <= 0 || D > 255) return bad_arg(5, "D must be in [1, 255]");
    if (Q < 0) return bad_arg(12, "Q < 0");
    if (n_actions <= 0) return bad_arg(13, "n_actions <= 0");
    if (Q == 0) return 0;
    if (!dist || !queries || !row_lo || !row_hi) return bad_arg(6, "dist/queries/row ranges is null");
    if (!out_moments) return bad_arg(14, "out_moments is null");
    if (M > 0 && (!buf_state || !buf_ret)) return bad_arg(1, "buffer is null");
    if (tail_lo < 0 || tail_hi > M || (tail_lo > tail_hi && tail_hi != 0)) return bad_arg(10, "tail range outside [0, M]");
    const size_t elem = buf_is_f64 ? sizeof(double) : sizeof(float);
    const size_t per_warp = (((size_t)n_actions * 3 * sizeof(double) + (size_t)(2 * D + 32 * (D + 1)) * elem) + 15) & ~(size_t)15;
    const size_t smem = per_warp * kRangeWarps;
    if (smem > 227 * 1024) return bad_arg(13, "n_actions / D too large for the per-warp bins (use dcarl_box_moments)");
    RangeArgs a;
    a.buf_state = buf_state; a.buf_ret = buf_ret; a.dist = dist; a.queries = queries;
    a.row_lo = reinterpret_cast<const long long *>(row_lo); a.row_hi = reinterpret_cast<const long long *>(row_hi);
    a.tail_lo = tail_lo; a.tail_hi = tail_hi > tail_lo ? tail_hi : tail_lo;
    a.out = out_moments; a.D = D; a.Q = Q; a.n_actions = n_actions; a.M = M;
    cudaStream_t st = (cudaStream_t)stream;
    long long grid = ((long long)Q + kRangeWarps - 1) / kRangeWarps;
    const long long cap = (long long)sm_count() * 8;
    if (grid > cap) grid = cap;
    if (buf_is_f64) {
        DCARL_CUDA(cudaFuncSetAttribute(k_box_ranges<double>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_box_ranges<double><<<(unsigned)grid, kRangeWarps * 32, smem, st>>>(a);
    } else {
        DCARL_CUDA(cudaFuncSetAttribute(k_box_ranges<float>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        k_box_ranges<float><<<(unsigned)grid, kRangeWarps * 32, smem, st>>>(a);
    }
    DCARL_LAUNCH_CHECK("k_box_ranges");
    return 0;
}
