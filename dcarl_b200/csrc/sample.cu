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

// One draw: state index (-1 when rejected, DS:50-51) and return R = Q*[s, a] + sigma * z_ret (DS:9).
template <typename QT>
__device__ __forceinline__ void draw(const QT *__restrict__ q_star, double zs, double zr, int a, int S, int A,
                                     double loc, double sd, double span, double sigma, int &so, double &r) {
    const int s = state_of(zs, loc, sd, span, S);
    r = 0.0;
    so = -1;
    if (s >= 0 && (unsigned)a < (unsigned)A) {
        so = s;
        r = __dadd_rn((double)q_star[(long long)s * A + a], __dmul_rn(sigma, zr));
    }
}

// 16-byte streaming accesses (four float32 / int32 elements)
__device__ __forceinline__ float4 ld_stream4(const float *p) {
    float4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ int4 ld_stream4(const int *p) {
    int4 v;
    asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream4(float *p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void st_stream4(int *p, int4 v) {
    asm volatile("st.global.L1::no_allocate.v4.s32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

template <typename QT, typename ZT, typename OT>
__global__ void __launch_bounds__(256) k_sample(const QT *__restrict__ q_star, const ZT *__restrict__ z_state,
                                                 const ZT *__restrict__ z_ret, const int *__restrict__ act,
                                                 long long n, int S, int A, double loc, double sd, double span,
                                                 double sigma, int *__restrict__ out_idx, OT *__restrict__ out_R,
                                                 int vec4) {
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    long long done = 0;
    if constexpr (sizeof(ZT) == 4 && sizeof(OT) == 4) {
        // float32 deviates and returns, 16-byte aligned arrays: four draws per thread and round, 16-byte accesses
        // (the scalar loop had one 4-byte load in flight per array and thread: 46 % of the HBM roofline)
        if (vec4) {
            const long long n4 = n >> 2;
            for (long long j = tid; j < n4; j += nthr) {
                const float4 zs = ld_stream4(reinterpret_cast<const float *>(z_state) + 4 * j);
                const float4 zr = ld_stream4(reinterpret_cast<const float *>(z_ret) + 4 * j);
                const int4 a = ld_stream4(act + 4 * j);
                int4 so;
                double r0, r1, r2, r3;
                draw(q_star, (double)zs.x, (double)zr.x, a.x, S, A, loc, sd, span, sigma, so.x, r0);
                draw(q_star, (double)zs.y, (double)zr.y, a.y, S, A, loc, sd, span, sigma, so.y, r1);
                draw(q_star, (double)zs.z, (double)zr.z, a.z, S, A, loc, sd, span, sigma, so.z, r2);
                draw(q_star, (double)zs.w, (double)zr.w, a.w, S, A, loc, sd, span, sigma, so.w, r3);
                st_stream4(out_idx + 4 * j, so);
                st_stream4(reinterpret_cast<float *>(out_R) + 4 * j, make_float4((float)r0, (float)r1, (float)r2, (float)r3));
            }
            done = n4 << 2;
        }
    }
    for (long long i = done + tid; i < n; i += nthr) {
        int so;
        double r;
        draw(q_star, (double)ld_stream(z_state + i), (double)ld_stream(z_ret + i), ld_stream(act + i), S, A, loc, sd, span,
             sigma, so, r);
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
                                                        OT *__restrict__ out_R, int vec4) {
    const float two_pi = 6.283185307179586f;
    auto one = [&](long long i, int &s, int &a, double &r) {     // element i: counter i, whatever thread computes it
        const uint4 u = philox4x32_10(make_uint4((unsigned)i, (unsigned)((unsigned long long)i >> 32), 0u, 0u), key);
        // Box-Muller on (u.x, u.y); u in (0, 1]
        const float u1 = ((float)(u.x >> 8) + 1.0f) * (1.0f / 16777216.0f);
        const float u2 = (float)(u.y >> 8) * (1.0f / 16777216.0f);
        const float rad = sqrtf(-2.0f * logf(u1));
        float sn, cs;
        sincosf(two_pi * u2, &sn, &cs);
        const double zs = (double)(rad * cs), zr = (double)(rad * sn);
        a = (int)(((unsigned long long)u.z * (unsigned)A) >> 32);   // uniform in [0, A)
        s = state_of(zs, loc, sd, span, S);
        r = 0.0;
        if (s >= 0) r = __dadd_rn((double)q_star[(long long)s * A + a], __dmul_rn(sigma, zr));
    };
    const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x, nthr = (long long)gridDim.x * blockDim.x;
    long long done = 0;
    if constexpr (sizeof(OT) == 4) {
        if (vec4) {          // four elements per thread and round: four independent Philox chains, 16-byte stores
            const long long n4 = n >> 2;
            for (long long j = tid; j < n4; j += nthr) {
                int4 s4, a4;
                double r0, r1, r2, r3;
                one(4 * j, s4.x, a4.x, r0); one(4 * j + 1, s4.y, a4.y, r1);
                one(4 * j + 2, s4.z, a4.z, r2); one(4 * j + 3, s4.w, a4.w, r3);
                st_stream4(out_idx + 4 * j, s4);
                st_stream4(out_act + 4 * j, a4);
                st_stream4(reinterpret_cast<float *>(out_R) + 4 * j, make_float4((float)r0, (float)r1, (float)r2, (float)r3));
            }
            done = n4 << 2;
        }
    }
    for (long long i = done + tid; i < n; i += nthr) {
        int s, a;
        double r;
        one(i, s, a, r);
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
    const int vec4 = !z_is_f64 && !out_is_f64 &&
                     ((((uintptr_t)z_state | (uintptr_t)z_ret | (uintptr_t)act | (uintptr_t)out_idx | (uintptr_t)out_R) & 15) == 0);
    const unsigned g = stream_grid(vec4 ? (n + 3) / 4 : n);
#define DCARL_SAMPLE(QT, ZT, OT)                                                                              \
    k_sample<QT, ZT, OT><<<g, 256, 0, st>>>((const QT *)q_star, (const ZT *)z_state, (const ZT *)z_ret, act, n, \
                                            S, A, loc, sd, span, sigma, out_idx, (OT *)out_R, vec4)
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
    const int vec4 = !out_is_f64 && ((((uintptr_t)out_idx | (uintptr_t)out_act | (uintptr_t)out_R) & 15) == 0);
    const unsigned g = stream_grid(vec4 ? (n + 3) / 4 : n);
    const uint2 key = make_uint2((unsigned)seed, (unsigned)(seed >> 32));
#define DCARL_PHILOX(QT, OT) \
    k_sample_philox<QT, OT><<<g, 256, 0, st>>>((const QT *)q_star, key, n, S, A, loc, sd, span, sigma, out_idx, out_act, (OT *)out_R, vec4)
    if (q_is_f64) { if (out_is_f64) DCARL_PHILOX(double, double); else DCARL_PHILOX(double, float); }
    else { if (out_is_f64) DCARL_PHILOX(float, double); else DCARL_PHILOX(float, float); }
#undef DCARL_PHILOX
    DCARL_LAUNCH_CHECK("k_sample_philox");
    return 0;
}
