// {state idx, action, return} sampler of Data_Generation.
//
// Replaces Simulation_testing/Simulation_Data_Collection/Data_Sampling/data_sampling.py:
// random_state_norm (:12-17), the rejection at :50-51 and add_an_act_data (:5-9, :55).
// scipy's norm.rvs(loc, scale) is loc + scale * standard_normal, so with the deviates
// supplied the draw is pure arithmetic; it is done in float64 with explicitly rounded
// operations (no FMA contraction) so that it reproduces the reference bit for bit.
#include "common.cuh"

namespace dcarl {

template <typename T>
__device__ __forceinline__ double ldd(const void *p, long long i) { return (double)reinterpret_cast<const T *>(p)[i]; }

__device__ __forceinline__ int state_of(double z, double loc, double sd, double span, int S) {
    // floor((loc + sd*z) / span * S), DS:14-15, each operation rounded on its own
    const double x = __dadd_rn(loc, __dmul_rn(sd, z));
    const double f = floor(__dmul_rn(__ddiv_rn(x, span), (double)S));
    return (f >= 0.0 && f < (double)S) ? (int)f : -1;
}

template <typename QT, typename ZT, typename OT>
__global__ void __launch_bounds__(256) k_sample(const QT *__restrict__ q_star, const ZT *__restrict__ z_state,
                                                 const ZT *__restrict__ z_ret, const int *__restrict__ act,
                                                 long long n, int S, int A, double loc, double sd, double span,
                                                 double sigma, int *__restrict__ out_idx, OT *__restrict__ out_R) {
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const int s = state_of((double)ld_stream(z_state + i), loc, sd, span, S);
        const int a = ld_stream(act + i);
        double r = 0.0;
        int so = -1;
        if (s >= 0 && (unsigned)a < (unsigned)A) {
            so = s;
            r = __dadd_rn((double)q_star[(long long)s * A + a], __dmul_rn(sigma, (double)ld_stream(z_ret + i)));   // DS:9
        }
        st_stream(out_idx + i, so);
        st_stream(out_R + i, (OT)r);
    }
}

// Philox4x32-10 (Salmon et al., SC'11), counter = (lo32(i), hi32(i), 0, 0), key = seed.
__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const unsigned hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const unsigned hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u; k.y += 0xBB67AE85u;
    }
    return c;
}

template <typename QT, typename OT>
__global__ void __launch_bounds__(256) k_sample_philox(const QT *__restrict__ q_star, uint2 key, long long n, int S,
                                                        int A, double loc, double sd, double span, double sigma,
                                                        int *__restrict__ out_idx, int *__restrict__ out_act,
                                                        OT *__restrict__ out_R) {
    const float two_pi = 6.283185307179586f;
    for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const uint4 u = philox4x32_10(make_uint4((unsigned)i, (unsigned)((unsigned long long)i >> 32), 0u, 0u), key);
        // Box-Muller on (u.x, u.y); u in (0, 1]
        const float u1 = ((float)(u.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
        const float u2 = (float)(u.y >> 8) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincosf(two_pi * u2, &sn, &cs);
        const double zs = (double)(rad * cs), zr = (double)(rad * sn);
        const int a = (int)(((unsigned long long)u.z * (unsigned)A) >> 32);   // uniform in [0, A)
        const int s = state_of(zs, loc, sd, span, S);
        double r = 0.0;
        if (s >= 0) r = __dadd_rn((double)q_star[(long long)s * A + a], __dmul_rn(sigma, zr));
        st_stream(out_idx + i, s);
        st_stream(out_act + i, a);
        st_stream(out_R + i, (OT)r);
    }
}

static unsigned stream_grid(long long n) {
    long long g = (n + 255) / 256;
    const long long cap = (long long)sm_count() * 16;
    return (unsigned)(g < cap ? (g > 0 ? g : 1) : cap);
}

}  // namespace dcarl

using namespace dcarl;

extern "C" int dcarl_sample_returns(const void *q_star, int q_is_f64, const void *z_state, const void *z_ret,
                                    int z_is_f64, const int32_t *act, int64_t n, int32_t S, int32_t A, double loc,
                                    double sd, double span, double sigma, int32_t *out_idx, void *out_R,
                                    int out_is_f64, dcarl_stream_t stream) {
    if (n < 0) return bad_arg(7, "n < 0");
    if (S <= 0 || A <= 0) return bad_arg(8, "S/A <= 0");
    if (!(span != 0.0)) return bad_arg(12, "span == 0");
    if (n == 0) return 0;
    if (!q_star || !z_state || !z_ret || !act) return bad_arg(1, "null input");
    if (!out_idx || !out_R) return bad_arg(14, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = stream_grid(n);
#define DCARL_SAMPLE(QT, ZT, OT)                                                                              \
    k_sample<QT, ZT, OT><<<g, 256, 0, st>>>((const QT *)q_star, (const ZT *)z_state, (const ZT *)z_ret, act, n, \
                                            S, A, loc, sd, span, sigma, out_idx, (OT *)out_R)
    if (q_is_f64) {
        if (z_is_f64) { if (out_is_f64) DCARL_SAMPLE(double, double, double); else DCARL_SAMPLE(double, double, float); }
        else { if (out_is_f64) DCARL_SAMPLE(double, float, double); else DCARL_SAMPLE(double, float, float); }
    } else {
        if (z_is_f64) { if (out_is_f64) DCARL_SAMPLE(float, double, double); else DCARL_SAMPLE(float, double, float); }
        else { if (out_is_f64) DCARL_SAMPLE(float, float, double); else DCARL_SAMPLE(float, float, float); }
    }
#undef DCARL_SAMPLE
    DCARL_LAUNCH_CHECK("k_sample");
    return 0;
}

extern "C" int dcarl_sample_returns_philox(const void *q_star, int q_is_f64, uint64_t seed, int64_t n, int32_t S,
                                           int32_t A, double loc, double sd, double span, double sigma,
                                           int32_t *out_idx, int32_t *out_act, void *out_R, int out_is_f64,
                                           dcarl_stream_t stream) {
    if (n < 0) return bad_arg(4, "n < 0");
    if (S <= 0 || A <= 0) return bad_arg(5, "S/A <= 0");
    if (!(span != 0.0)) return bad_arg(9, "span == 0");
    if (n == 0) return 0;
    if (!q_star) return bad_arg(1, "q_star is null");
    if (!out_idx || !out_act || !out_R) return bad_arg(11, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    const unsigned g = stream_grid(n);
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#define DCARL_PHILOX(QT, OT) \
    k_sample_philox<QT, OT><<<g, 256, 0, st>>>((const QT *)q_star, key, n, S, A, loc, sd, span, sigma, out_idx, out_act, (OT *)out_R)
    if (q_is_f64) { if (out_is_f64) DCARL_PHILOX(double, double); else DCARL_PHILOX(double, float); }
    else { if (out_is_f64) DCARL_PHILOX(float, double); else DCARL_PHILOX(float, float); }
#undef DCARL_PHILOX
    DCARL_LAUNCH_CHECK("k_sample_philox");
    return 0;
}
