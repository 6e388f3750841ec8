// upper_bound / lower_bound / CI_lower_bound / mean_value over sample arrays.
//
// Replaces Simulation_testing/Simulation_1/test_DCARL.py:10-28 (identical copy in
// Simulation_2/test_DCARL.py:9-27).  One CTA per sample array, two passes like
// np.mean / np.std: sum -> mean, then sum of squared deviations -> population std.
// The closed forms are evaluated with IEEE divide/sqrt in the reference's operator order.
#include <math_constants.h>

#include "common.cuh"

namespace dcarl {

__device__ __forceinline__ double block_sum(double v, double *sh) {
    // fixed-order: warp butterfly, then warp 0 over the per-warp partials
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(FULL_MASK, v, off);
    const int w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    __syncthreads();
    if (lane_id() == 0) sh[w] = v;
    __syncthreads();
    double t = (int)lane_id() < nw ? sh[lane_id()] : 0.0;
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) t += __shfl_xor_sync(FULL_MASK, t, off);
    return t;  // identical in every thread
}

template <typename XT>
__global__ void __launch_bounds__(256) k_bounds(const XT *__restrict__ x, const long long *__restrict__ seg,
                                                 long long n_seg, double logterm, double scale, double cap,
                                                 double *__restrict__ out) {
    __shared__ double sh[8];
    for (long long sidx = blockIdx.x; sidx < n_seg; sidx += gridDim.x) {
        const long long b = seg[sidx], e = seg[sidx + 1];
        const long long n = e - b;
        double s = 0.0;
        for (long long i = b + threadIdx.x; i < e; i += blockDim.x) s += (double)x[i];
        const double dsum = block_sum(s, sh);
        const double dn = (double)n, dn1 = (double)(n + 1);
        const double mean = dsum / dn;
        double q = 0.0;
        for (long long i = b + threadIdx.x; i < e; i += blockDim.x) { const double d = (double)x[i] - mean; q = fma(d, d, q); }
        const double sigma = sqrt(block_sum(q, sh) / dn);
        if (threadIdx.x == 0) {
            double *o = out + sidx * 4;
            if (n <= 0) { o[0] = o[1] = o[2] = o[3] = CUDART_NAN; }
            else {
                const double h = scale * sqrt(logterm / 2 / dn);
                o[0] = fmin(cap, mean + h);                                               // ST1:12
                o[1] = mean - h;                                                          // ST1:16
                o[2] = dsum / dn / dn1 - 4 * sigma / dn1 + dsum / dn1 - scale * sqrt(logterm / 2 / dn1);   // ST1:24
                o[3] = fmin(cap, mean);                                                   // ST1:28
            }
        }
    }
}

}  // namespace dcarl

using namespace dcarl;

extern "C" int dcarl_bounds(const void *x, int x_is_f64, const int64_t *seg_offsets, int64_t n_seg,
                            const dcarl_params *p, double *out, dcarl_stream_t stream) {
    if (n_seg < 0) return bad_arg(4, "n_seg < 0");
    if (!p) return bad_arg(5, "params is null");
    if (!(p->alpha > 0.0 && p->alpha <= 1.0)) return bad_arg(5, "alpha must be in (0, 1]");
    if (n_seg == 0) return 0;
    if (!seg_offsets) return bad_arg(3, "seg_offsets is null");
    if (!out) return bad_arg(6, "out is null");
    const double logterm = log(1.0 / p->alpha);
    long long grid = n_seg < 148 * 8 ? n_seg : 148 * 8;
    cudaStream_t st = (cudaStream_t)stream;
    const long long *seg = reinterpret_cast<const long long *>(seg_offsets);
    if (x_is_f64)
        k_bounds<double><<<(unsigned)grid, 256, 0, st>>>((const double *)x, seg, n_seg, logterm, p->scale, p->cap, out);
    else
        k_bounds<float><<<(unsigned)grid, 256, 0, st>>>((const float *)x, seg, n_seg, logterm, p->scale, p->cap, out);
    DCARL_LAUNCH_CHECK("k_bounds");
    return 0;
}
