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
// Data movement: the AoS buffer [M, D+1] float32 is staged tile by tile into shared
// memory with 1-D TMA bulk copies (cp.async.bulk + mbarrier, 4-stage ring), so global
// reads are whole 16-byte-aligned contiguous tiles regardless of the odd row pitch, and
// a thread reads "its" row from shared memory (pitch D+1 = 21 words: conflict free).
#include <math_constants.h>

#include "common.cuh"
#include "peer.cuh"

namespace dcarl {

constexpr int kTileRows = 256;
constexpr int kStages = 4;
constexpr int kMaxSmemBins = 1024;  // (query, action) bins kept in shared memory: 24 KB

// Largest float32 s_hi and smallest float32 s_lo with  fl64(s - w) <= q <= fl64(s + w)  <=>  s_lo <= s <= s_hi
// (exact float64 predicate, +-1 ulp walk around the rounded candidates); an empty interval for NaN inputs.
__device__ __forceinline__ void box_thresholds(double q, double w, float &l, float &h) {
    h = -CUDART_INF_F; l = CUDART_INF_F;
    if (q == q && w == w) {
        // hi: largest float s with (double)s - w <= q
        float c = __double2float_rd(q + w);
        for (int it = 0; it < 4; ++it) {
            const float up = nextafterf(c, CUDART_INF_F);
            if (up != c && (double)up - w <= q) c = up; else break;
        }
        for (int it = 0; it < 4 && !((double)c - w <= q); ++it) c = nextafterf(c, -CUDART_INF_F);
        if ((double)c - w <= q) h = c;
        // lo: smallest float s with q <= (double)s + w
        c = __double2float_ru(q - w);
        for (int it = 0; it < 4; ++it) {
            const float dn = nextafterf(c, -CUDART_INF_F);
            if (dn != c && q <= (double)dn + w) c = dn; else break;
        }
        for (int it = 0; it < 4 && !(q <= (double)c + w); ++it) c = nextafterf(c, CUDART_INF_F);
        if (q <= (double)c + w) l = c;
    }
}

// thresholds layout in the workspace: lo[Q][D], hi[Q][D] (float32)
__global__ void k_box_prepare(const double *__restrict__ queries, const double *__restrict__ dist, int Q,
                              int D, float *__restrict__ lo, float *__restrict__ hi) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= Q * D) return;
    float l, h;
    box_thresholds(queries[i], dist[i % D], l, h);
    lo[i] = l;
    hi[i] = h;
}

struct BoxArgs {
    const float *buf_state;
    const float *buf_ret;
    long long M;
    int D, Q, n_actions;
    const float *lo, *hi;
    const double *dist;  // [D+1]; dist[D] is the action tolerance
    double *out;  // [Q, n_actions, 3]; the accumulator (zero on entry) in the fused query
    int bins_in_smem;
    // fused query (k_box_moments<.., true>): thresholds in the prologue, reduction + select in the epilogue
    const double *queries;     // [Q, D]
    unsigned *ticket;          // CTAs that have flushed their bins; the last one runs the epilogue and resets it
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

template <int DT, bool FUSED>
__global__ void __launch_bounds__(kTileRows) k_box_moments(const BoxArgs p) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int D = DT > 0 ? DT : p.D;
    const int W = D + 1;
    const int tid = threadIdx.x;
    const size_t tile_bytes = (size_t)kTileRows * W * sizeof(float);
    const size_t ret_bytes = (size_t)kTileRows * sizeof(float);
    // layout: [stages x state tile][stages x ret tile][lo Q*D][hi Q*D][bins][mbarriers]
    float *st_tile = reinterpret_cast<float *>(smem_raw);
    float *rt_tile = reinterpret_cast<float *>(smem_raw + kStages * tile_bytes);
    float *slo = rt_tile + kStages * kTileRows;
    float *shi = slo + p.Q * D;
    size_t off = kStages * (tile_bytes + ret_bytes) + (size_t)2 * p.Q * D * sizeof(float);
    off = (off + 15) & ~(size_t)15;
    // head thresholds {lo0, hi0, lo1, hi1 | lo2, hi2, lo3, hi3} per query: two 16-byte broadcast loads
    // per (warp, query) instead of eight 4-byte ones (with many queries the scan is bound by them)
    float4 *shead = reinterpret_cast<float4 *>(smem_raw + off);
    off += (size_t)p.Q * 2 * sizeof(float4);
    double *sbins = reinterpret_cast<double *>(smem_raw + off);
    const int nbins = p.Q * p.n_actions;
    off += p.bins_in_smem ? (size_t)nbins * 3 * sizeof(double) : 0;
    uint64_t *full = reinterpret_cast<uint64_t *>(smem_raw + off);

    if (FUSED) {
        // every CTA derives the thresholds of its queries itself (Q*D of them: trivial next to a tile)
        for (int i = tid; i < p.Q * D; i += kTileRows) box_thresholds(p.queries[i], p.dist[i % D], slo[i], shi[i]);
        __syncthreads();
    } else {
        for (int i = tid; i < p.Q * D; i += kTileRows) { slo[i] = p.lo[i]; shi[i] = p.hi[i]; }
    }
    if (DT >= 4) {
        if (!FUSED) __syncthreads();
        for (int q = tid; q < p.Q; q += kTileRows) {
            const float *l = slo + (size_t)q * D, *h = shi + (size_t)q * D;
            shead[2 * q] = make_float4(l[0], h[0], l[1], h[1]);
            shead[2 * q + 1] = make_float4(l[2], h[2], l[3], h[3]);
        }
    }
    if (p.bins_in_smem)
        for (int i = tid; i < nbins * 3; i += kTileRows) sbins[i] = 0.0;
    if (tid == 0) {
        for (int s = 0; s < kStages; ++s) mbar_init(&full[s], 1);
        mbar_fence_init();
    }
    __syncthreads();

    const long long n_tiles = (p.M + kTileRows - 1) / kTileRows;
    const long long n_full = p.M / kTileRows;  // tiles that are complete (TMA-loadable)
    auto issue = [&](long long tile, int stage) {
        // thread 0 only; complete tiles only
        mbar_arrive_expect_tx(&full[stage], (unsigned)(tile_bytes + ret_bytes));
        tma_load_1d(st_tile + (size_t)stage * kTileRows * W, p.buf_state + tile * kTileRows * W,
                    (unsigned)tile_bytes, &full[stage]);
        tma_load_1d(rt_tile + (size_t)stage * kTileRows, p.buf_ret + tile * kTileRows, (unsigned)ret_bytes,
                    &full[stage]);
    };
    if (tid == 0)
        for (int s = 0; s < kStages; ++s) {
            const long long t = (long long)blockIdx.x + (long long)s * gridDim.x;
            if (t < n_full) issue(t, s);
        }

    double *bins = p.bins_in_smem ? sbins : p.out;
    const double dist_act = p.dist[D];
    long long it = 0;
    for (long long tile = blockIdx.x; tile < n_tiles; tile += gridDim.x, ++it) {
        const int stage = (int)(it % kStages);
        float *ts = st_tile + (size_t)stage * kTileRows * W;
        float *tr = rt_tile + (size_t)stage * kTileRows;
        if (tile < n_full) {
            mbar_wait(&full[stage], (unsigned)((it / kStages) & 1));
        } else {
            // ragged last tile: plain cooperative loads
            const long long base = tile * kTileRows;
            const int rows = (int)(p.M - base);
            for (int i = tid; i < rows * W; i += kTileRows) ts[i] = p.buf_state[base * W + i];
            if (tid < rows) tr[tid] = p.buf_ret[base + tid];
            __syncthreads();
        }
        const long long row = tile * kTileRows + tid;
        const bool in_buf = row < p.M;
        const float *x = ts + (size_t)tid * W;
        float xr[DT > 0 ? DT : 1];
        if (DT > 0) {
#pragma unroll
            for (int d = 0; d < DT; ++d) xr[d] = x[d];
        }
        for (int q = 0; q < p.Q; ++q) {
            const float *l = slo + q * D, *h = shi + q * D;
            bool ok = in_buf;
            if (DT > 0) {
                constexpr int kHead = DT < 4 ? DT : 4;
                if (DT >= 4) {
                    const float4 t0 = shead[2 * q], t1 = shead[2 * q + 1];
                    ok = ok & (xr[0] >= t0.x) & (xr[0] <= t0.y) & (xr[1] >= t0.z) & (xr[1] <= t0.w) &
                         (xr[2] >= t1.x) & (xr[2] <= t1.y) & (xr[3] >= t1.z) & (xr[3] <= t1.w);
                } else {
#pragma unroll
                    for (int d = 0; d < kHead; ++d) ok = ok && (xr[d] >= l[d]) && (xr[d] <= h[d]);
                }
                // with many queries most (row, query) pairs die on the first dims: leave before the rest
                if (p.Q > 1 && !__any_sync(FULL_MASK, ok)) continue;
#pragma unroll
                for (int d = kHead; d < DT; ++d) ok = ok && (xr[d] >= l[d]) && (xr[d] <= h[d]);
            } else {
                for (int d = 0; d < D; ++d) { const float v = x[d]; ok = ok && (v >= l[d]) && (v <= h[d]); }
            }
            if (!__any_sync(FULL_MASK, ok)) continue;
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
        }
        __syncthreads();  // everyone is done with this stage
        if (tid == 0) {
            const long long nt = tile + (long long)kStages * gridDim.x;
            if (nt < n_full) issue(nt, stage);
        }
    }
    if (p.bins_in_smem) {
        __syncthreads();
        for (int i = tid; i < nbins * 3; i += kTileRows) {
            const double v = sbins[i];
            if (v != 0.0) atomicAdd(p.out + i, v);
        }
    }
    if (!FUSED) return;
    // ---- fused epilogue: the last CTA to arrive owns the totals
    __shared__ unsigned s_ticket;
    __threadfence();
    __syncthreads();
    if (tid == 0) s_ticket = atomicAdd(p.ticket, 1u);
    __syncthreads();
    if (s_ticket != gridDim.x - 1) return;
    __threadfence();
    // totals of this GPU -> out_moments; the accumulator and the ticket are left zeroed for the next call
    for (int i = tid; i < nbins * 3; i += kTileRows) {
        const double v = __ldcg(p.out + i);
        p.out[i] = 0.0;
        p.out_moments[i] = v;
    }
    if (tid == 0) *p.ticket = 0u;
    __syncthreads();
    if (p.peers.world > 1) peer_allreduce_cta(p.out_moments, p.out_moments, nbins * 3, p.peers);
    __syncthreads();
    for (int q = tid; q < p.Q; q += kTileRows)
        box_select_one(p.out_moments + (size_t)q * p.n_actions * 3, p.n_actions, p.rp,
                       p.out_mv ? p.out_mv + (size_t)q * p.n_actions * 2 : nullptr,
                       p.out_action ? p.out_action + q : nullptr);
}

// moments -> mean/var (+ sentinel) and the act_test pick, one thread per query
__global__ void k_box_select(const double *__restrict__ mom, int Q, int n_actions, dcarl_rls_params rp,
                             double *__restrict__ out_mv, int *__restrict__ out_action) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;
    if (q >= Q) return;
    box_select_one(mom + (size_t)q * n_actions * 3, n_actions, rp,
                   out_mv ? out_mv + (size_t)q * n_actions * 2 : nullptr, out_action ? out_action + q : nullptr);
}

static size_t box_smem_bytes(int Q, int D, int n_actions, bool bins_in_smem) {
    const int W = D + 1;
    size_t off = (size_t)kStages * kTileRows * (W + 1) * sizeof(float) + (size_t)2 * Q * D * sizeof(float);
    off = (off + 15) & ~(size_t)15;
    off += (size_t)Q * 32;   // head thresholds
    if (bins_in_smem) off += (size_t)Q * n_actions * 3 * sizeof(double);
    off += kStages * sizeof(uint64_t);
    return off;
}

}  // namespace dcarl

using namespace dcarl;

extern "C" size_t dcarl_box_moments_ws(int32_t Q, int32_t D, int32_t n_actions) {
    (void)n_actions;
    if (Q <= 0 || D <= 0) return 0;
    return (size_t)2 * Q * D * sizeof(float);
}

extern "C" int dcarl_box_moments(const float *buf_state, const float *buf_ret, int64_t M, int32_t D,
                                 const double *dist, const double *queries, int32_t Q, int32_t n_actions,
                                 double *out_moments, void *ws, size_t ws_bytes, dcarl_stream_t stream) {
    if (M < 0) return bad_arg(3, "M < 0");
    if (D <= 0 || D > 255) return bad_arg(4, "D must be in [1, 255]");
    if (Q < 0) return bad_arg(7, "Q < 0");
    if (n_actions <= 0) return bad_arg(8, "n_actions <= 0");
    if (Q == 0) return 0;
    if (!dist || !queries) return bad_arg(5, "dist/queries is null");
    if (!out_moments) return bad_arg(9, "out_moments is null");
    if (M > 0 && (!buf_state || !buf_ret)) return bad_arg(1, "buffer is null");
    if (((uintptr_t)buf_state & 15) || ((uintptr_t)buf_ret & 15)) return bad_arg(1, "buffer must be 16-byte aligned");
    if (!ws || ws_bytes < dcarl_box_moments_ws(Q, D, n_actions))
        return bad_arg(10, "workspace too small (see dcarl_box_moments_ws)");
    cudaStream_t st = (cudaStream_t)stream;
    DCARL_CUDA(cudaMemsetAsync(out_moments, 0, (size_t)Q * n_actions * 3 * sizeof(double), st));
    if (M == 0) return 0;

    float *lo = reinterpret_cast<float *>(ws);
    float *hi = lo + (size_t)Q * D;
    k_box_prepare<<<(Q * D + 127) / 128, 128, 0, st>>>(queries, dist, Q, D, lo, hi);
    DCARL_LAUNCH_CHECK("k_box_prepare");

    // queries are processed in chunks whose thresholds (+ bins) fit beside the tile ring
    const size_t budget = 200 * 1024;
    int qc = Q;
    while (qc > 1 && box_smem_bytes(qc, D, n_actions, false) > budget) qc = (qc + 1) / 2;
    if (box_smem_bytes(qc, D, n_actions, false) > budget) return bad_arg(4, "D too large for the tile ring");
    const int sms = sm_count();
    const long long n_tiles = (M + kTileRows - 1) / kTileRows;
    for (int q0 = 0; q0 < Q; q0 += qc) {
        const int qn = (Q - q0 < qc) ? Q - q0 : qc;
        BoxArgs a;
        a.buf_state = buf_state; a.buf_ret = buf_ret; a.M = M; a.D = D; a.Q = qn; a.n_actions = n_actions;
        a.lo = lo + (size_t)q0 * D; a.hi = hi + (size_t)q0 * D; a.dist = dist;
        a.out = out_moments + (size_t)q0 * n_actions * 3;
        a.bins_in_smem = (qn * n_actions <= kMaxSmemBins) &&
                         box_smem_bytes(qn, D, n_actions, true) <= budget;
        const size_t smem = box_smem_bytes(qn, D, n_actions, a.bins_in_smem);
        int per_sm = (int)((size_t)(220 * 1024) / smem);
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        long long grid = (long long)sms * per_sm;
        if (grid > n_tiles) grid = n_tiles;
        if (D == 20) {
            DCARL_CUDA(cudaFuncSetAttribute(k_box_moments<20, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_box_moments<20, false><<<(unsigned)grid, kTileRows, smem, st>>>(a);
        } else {
            DCARL_CUDA(cudaFuncSetAttribute(k_box_moments<0, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_box_moments<0, false><<<(unsigned)grid, kTileRows, smem, st>>>(a);
        }
        DCARL_LAUNCH_CHECK("k_box_moments");
    }
    return 0;
}

extern "C" size_t dcarl_box_query_ws(int32_t Q, int32_t D, int32_t n_actions) {
    (void)D;
    if (Q <= 0 || n_actions <= 0) return 0;
    return (size_t)Q * n_actions * 3 * sizeof(double) + 16;   // accumulator + ticket
}

extern "C" int dcarl_box_query(const float *buf_state, const float *buf_ret, int64_t M, int32_t D,
                               const double *dist, const double *queries, int32_t Q, int32_t n_actions,
                               const dcarl_rls_params *rp, double *out_moments, double *out_mean_var,
                               int32_t *out_action, void *ws, size_t ws_bytes, const dcarl_peer_ctx *peers,
                               dcarl_stream_t stream) {
    if (M < 0) return bad_arg(3, "M < 0");
    if (D <= 0 || D > 255) return bad_arg(4, "D must be in [1, 255]");
    if (Q < 0) return bad_arg(7, "Q < 0");
    if (n_actions <= 0) return bad_arg(8, "n_actions <= 0");
    if (!rp) return bad_arg(9, "params is null");
    if (Q == 0) return 0;
    if (!dist || !queries) return bad_arg(5, "dist/queries is null");
    if (!out_moments) return bad_arg(10, "out_moments is null");
    if (M > 0 && (!buf_state || !buf_ret)) return bad_arg(1, "buffer is null");
    if (((uintptr_t)buf_state & 15) || ((uintptr_t)buf_ret & 15)) return bad_arg(1, "buffer must be 16-byte aligned");
    if (!ws || ws_bytes < dcarl_box_query_ws(Q, D, n_actions) || ((uintptr_t)ws & 15))
        return bad_arg(13, "workspace too small or misaligned (see dcarl_box_query_ws)");
    cudaStream_t st = (cudaStream_t)stream;
    double *acc = reinterpret_cast<double *>(ws);
    unsigned *ticket = reinterpret_cast<unsigned *>(acc + (size_t)Q * n_actions * 3);

    const size_t budget = 200 * 1024;
    int qc = Q;
    while (qc > 1 && box_smem_bytes(qc, D, n_actions, false) > budget) qc = (qc + 1) / 2;
    if (box_smem_bytes(qc, D, n_actions, false) > budget) return bad_arg(4, "D too large for the tile ring");
    PeerCtx pc;
    {
        const int nmsg = ((Q < qc) ? Q : qc) * n_actions * 3;
        if (int rc = make_peer_ctx(peers, nmsg, 15, pc)) return rc;
    }
    const int sms = sm_count();
    const long long n_tiles = (M + kTileRows - 1) / kTileRows;
    for (int q0 = 0; q0 < Q; q0 += qc) {
        const int qn = (Q - q0 < qc) ? Q - q0 : qc;
        BoxArgs a;
        a.buf_state = buf_state; a.buf_ret = buf_ret; a.M = M; a.D = D; a.Q = qn; a.n_actions = n_actions;
        a.lo = nullptr; a.hi = nullptr; a.dist = dist;
        a.out = acc + (size_t)q0 * n_actions * 3;
        a.bins_in_smem = (qn * n_actions <= kMaxSmemBins) &&
                         box_smem_bytes(qn, D, n_actions, true) <= budget;
        a.queries = queries + (size_t)q0 * D;
        a.ticket = ticket;
        a.out_moments = out_moments + (size_t)q0 * n_actions * 3;
        a.out_mv = out_mean_var ? out_mean_var + (size_t)q0 * n_actions * 2 : nullptr;
        a.out_action = out_action ? out_action + q0 : nullptr;
        a.rp = *rp;
        a.peers = pc;
        const size_t smem = box_smem_bytes(qn, D, n_actions, a.bins_in_smem);
        int per_sm = (int)((size_t)(220 * 1024) / smem);
        if (per_sm < 1) per_sm = 1;
        if (per_sm > 4) per_sm = 4;
        long long grid = (long long)sms * per_sm;
        if (grid > n_tiles) grid = n_tiles;
        if (grid < 1) grid = 1;                                   // an empty shard still takes part in the reduction
        if (D == 20) {
            DCARL_CUDA(cudaFuncSetAttribute(k_box_moments<20, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_box_moments<20, true><<<(unsigned)grid, kTileRows, smem, st>>>(a);
        } else {
            DCARL_CUDA(cudaFuncSetAttribute(k_box_moments<0, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
            k_box_moments<0, true><<<(unsigned)grid, kTileRows, smem, st>>>(a);
        }
        DCARL_LAUNCH_CHECK("k_box_moments (fused query)");
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
