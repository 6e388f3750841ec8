// Shared helpers for the dcarl_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/dcarl_b200.h"

#define FULL_MASK 0xffffffffu

namespace dcarl {

// ---- host-side error plumbing (abi.cu owns the thread-local string) ----------------------
void set_error(const char *fmt, ...);
int bad_arg(int index, const char *what);
int cuda_fail(cudaError_t e, const char *where);

#define DCARL_CUDA(call)                                                    \
    do {                                                                    \
        cudaError_t e__ = (call);                                           \
        if (e__ != cudaSuccess) return ::dcarl::cuda_fail(e__, #call);      \
    } while (0)
#define DCARL_LAUNCH_CHECK(name)                                            \
    do {                                                                    \
        cudaError_t e__ = cudaGetLastError();                               \
        if (e__ != cudaSuccess) return ::dcarl::cuda_fail(e__, name);       \
    } while (0)

int sm_count();   // of the current device (queried per call: the library keeps no state)

// ---- device helpers ----------------------------------------------------------------------
__device__ __forceinline__ unsigned lane_id() { return threadIdx.x & 31u; }
__device__ __forceinline__ unsigned lanemask_lt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
    return m;
}
__device__ __forceinline__ unsigned lanemask_le() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_le;" : "=r"(m));
    return m;
}
__device__ __forceinline__ unsigned lanemask_gt() {
    unsigned m;
    asm("mov.u32 %0, %%lanemask_gt;" : "=r"(m));
    return m;
}

// Streaming (read-once / write-once) global accesses: keep them out of L1.
__device__ __forceinline__ int ld_stream(const int *p) {
    int v;
    asm volatile("ld.global.nc.L1::no_allocate.s32 %0, [%1];" : "=r"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ uint8_t ld_stream(const uint8_t *p) {
    unsigned v;
    asm volatile("ld.global.nc.L1::no_allocate.u8 %0, [%1];" : "=r"(v) : "l"(p));
    return (uint8_t)v;
}
__device__ __forceinline__ float ld_stream(const float *p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ double ld_stream(const double *p) {
    double v;
    asm volatile("ld.global.nc.L1::no_allocate.f64 %0, [%1];" : "=d"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void st_stream(int *p, int v) {
    asm volatile("st.global.L1::no_allocate.s32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void st_stream(uint8_t *p, uint8_t v) {
    asm volatile("st.global.L1::no_allocate.u8 [%0], %1;" ::"l"(p), "r"((unsigned)v) : "memory");
}
__device__ __forceinline__ void st_stream(float *p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ void st_stream(double *p, double v) {
    asm volatile("st.global.L1::no_allocate.f64 [%0], %1;" ::"l"(p), "d"(v) : "memory");
}

// Shared-memory accesses by 32-bit shared-space address (no generic-pointer window arithmetic).
#if defined(DCARL_V_GENERIC_HOT)   // A/B variant: the same accesses through generic pointers
__device__ __forceinline__ void lds_v2f64(unsigned a, double &x, double &y) {
    const double2 v = *reinterpret_cast<const double2 *>(__cvta_shared_to_generic(a));
    x = v.x; y = v.y;
}
__device__ __forceinline__ double lds_f64(unsigned a) { return *reinterpret_cast<const double *>(__cvta_shared_to_generic(a)); }
__device__ __forceinline__ void sts_f64(unsigned a, double v) { *reinterpret_cast<double *>(__cvta_shared_to_generic(a)) = v; }
__device__ __forceinline__ void sts_v2f64(unsigned a, double x, double y) {
    *reinterpret_cast<double2 *>(__cvta_shared_to_generic(a)) = make_double2(x, y);
}
#else
#if defined(DCARL_V_NOCLOBBER)
#define DCARL_SMEM_CLOBBER
#else
#define DCARL_SMEM_CLOBBER : "memory"
#endif
__device__ __forceinline__ void lds_v2f64(unsigned a, double &x, double &y) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a) DCARL_SMEM_CLOBBER);
}
__device__ __forceinline__ double lds_f64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) DCARL_SMEM_CLOBBER);
    return v;
}
__device__ __forceinline__ void sts_f64(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) DCARL_SMEM_CLOBBER);
}
__device__ __forceinline__ void sts_v2f64(unsigned a, double x, double y) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %2};" ::"r"(a), "d"(x), "d"(y) DCARL_SMEM_CLOBBER);
}
#endif

__device__ __forceinline__ double shfl_f64(double v, int src) {
    return __shfl_sync(FULL_MASK, v, src);
}

// ---- thread-block clusters: rank, distributed shared memory, barrier ------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ unsigned cluster_nctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_nctarank;" : "=r"(r));
    return r;
}
// shared::cta address -> the same variable in CTA `rank` of the cluster (shared::cluster address)
__device__ __forceinline__ uint32_t mapa_shared(uint32_t addr, unsigned rank) {
    uint32_t r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
    return r;
}
__device__ __forceinline__ void st_shared_cluster(uint32_t addr, double v) {
    asm volatile("st.shared::cluster.f64 [%0], %1;" ::"r"(addr), "d"(v) : "memory");
}
// Every CTA of the cluster has started (required before the first access to another CTA's shared memory).
__device__ __forceinline__ void cluster_arrive_wait_entered() {
    asm volatile("barrier.cluster.arrive.relaxed.aligned;\nbarrier.cluster.wait.aligned;" ::: "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

// ---- mbarrier + 1-D TMA bulk copy (cp.async.bulk, SASS UBLKCP) ---------------------------
__device__ __forceinline__ uint32_t smem_u32(const void *p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t *bar, unsigned parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
// global -> shared bulk copy; dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void tma_load_1d(void *smem_dst, const void *gmem_src, unsigned bytes,
                                            uint64_t *bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(smem_dst)),
        "l"(gmem_src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace dcarl
