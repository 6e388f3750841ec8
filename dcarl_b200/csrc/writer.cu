// Batched RLS.add_data return construction (SURVEY "next" row f2).
//
// Replaces the per-step deque logic of RLS.add_data
// (Field_testing/.../tools/DCARL/stable_baselines/deepq/RLS.py:185-215) for many concurrent
// episode streams.  The reference keeps the last transitions in a deque: a transition is
// recorded with its OWN reward once 10 newer ones have arrived (:188-199), and when a
// transition with done=True arrives everything still queued is recorded with the terminal
// reward discounted by gamma per step of distance (:202-215).  Transitions still queued when
// the stream stops are never recorded.  Closed form for transition t of a stream of length T:
//     u = first done index in [t, t+lag-1]      ->  R = rew[u] * gamma^(u - t)
//     else if t + lag <= T - 1                  ->  R = rew[t]
//     else                                      ->  not recorded
// Recorded transitions of a stream are a prefix [0, T_rec), so the compacted output of stream
// e starts at row_offsets[e] (exclusive scan of the counts, done by the caller).
#include "common.cuh"

namespace dcarl {

__device__ __forceinline__ bool recorded_return(const float *rew, const unsigned char *done, long long b, long long t,
                                                long long T, int lag, double gamma, float &R) {
    const long long lim = t + lag < T ? t + lag : T;
    for (long long u = t; u < lim; ++u)
        if (done[b + u]) {                                    // flushed by the done at u, RLS.py:207
            R = (float)((double)rew[b + u] * pow(gamma, (double)(u - t)));   // rew_right * gamma ** len(buffer)
            return true;
        }
    if (t + lag <= T - 1) { R = rew[b + t]; return true; }   // popped with its own reward, RLS.py:189-192
    return false;
}

// one thread per stream: number of recorded transitions.  Recorded rows are a prefix:
// T_rec = max(last done index + 1, T - lag, 0), and only a done among the last `lag`
// transitions can push it past T - lag.
__global__ void __launch_bounds__(256) k_add_data_counts(const unsigned char *__restrict__ done,
                                                          const long long *__restrict__ ep_offsets, long long E,
                                                          int lag, long long *__restrict__ counts) {
    const long long e = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= E) return;
    const long long b = ep_offsets[e], T = ep_offsets[e + 1] - b;
    long long c = T - lag > 0 ? T - lag : 0;
    for (long long t = T - 1; t >= c; --t)
        if (done[b + t]) { c = t + 1; break; }
    counts[e] = c;
}

__global__ void __launch_bounds__(256) k_add_data_rows(const float *__restrict__ obs, const int *__restrict__ action,
                                                        const float *__restrict__ rew,
                                                        const unsigned char *__restrict__ done,
                                                        const long long *__restrict__ ep_offsets,
                                                        const long long *__restrict__ row_offsets,
                                                        const int *__restrict__ ep_of_row, long long N, int D, int lag,
                                                        double gamma, float *__restrict__ out_state,
                                                        float *__restrict__ out_ret) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;   // global transition index
    if (i >= N) return;
    const int e = ep_of_row[i];
    const long long b = ep_offsets[e], T = ep_offsets[e + 1] - b, t = i - b;
    float R;
    if (!recorded_return(rew, done, b, t, T, lag, gamma, R)) return;
    const long long o = row_offsets[e] + t;
    const float *src = obs + i * D;
    float *dst = out_state + o * (D + 1);
    for (int d = 0; d < D; ++d) dst[d] = src[d];
    dst[D] = (float)action[i];                                              // state_with_action, RLS.py:91-93
    out_ret[o] = R;
}

}  // namespace dcarl

using namespace dcarl;

extern "C" int dcarl_add_data_counts(const float *rew, const uint8_t *done, const int64_t *ep_offsets, int64_t E,
                                     int32_t lag, int64_t *out_counts, dcarl_stream_t stream) {
    if (E < 0) return bad_arg(4, "E < 0");
    if (lag <= 0) return bad_arg(5, "lag <= 0");
    if (E == 0) return 0;
    (void)rew;
    if (!done || !ep_offsets) return bad_arg(2, "null input");
    if (!out_counts) return bad_arg(6, "out_counts is null");
    k_add_data_counts<<<(unsigned)((E + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        done, reinterpret_cast<const long long *>(ep_offsets), E, lag, reinterpret_cast<long long *>(out_counts));
    DCARL_LAUNCH_CHECK("k_add_data_counts");
    return 0;
}

extern "C" int dcarl_add_data_rows(const float *obs, const int32_t *action, const float *rew, const uint8_t *done,
                                   const int64_t *ep_offsets, const int64_t *row_offsets, const int32_t *ep_of_row,
                                   int64_t N, int32_t D, int32_t lag, double gamma, float *out_state, float *out_ret,
                                   dcarl_stream_t stream) {
    if (N < 0) return bad_arg(8, "N < 0");
    if (D <= 0) return bad_arg(9, "D <= 0");
    if (lag <= 0) return bad_arg(10, "lag <= 0");
    if (N == 0) return 0;
    if (!obs || !action || !rew || !done || !ep_offsets || !row_offsets || !ep_of_row) return bad_arg(1, "null input");
    if (!out_state || !out_ret) return bad_arg(12, "null output");
    k_add_data_rows<<<(unsigned)((N + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        obs, action, rew, done, reinterpret_cast<const long long *>(ep_offsets),
        reinterpret_cast<const long long *>(row_offsets), ep_of_row, N, D, lag, gamma, out_state, out_ret);
    DCARL_LAUNCH_CHECK("k_add_data_rows");
    return 0;
}
