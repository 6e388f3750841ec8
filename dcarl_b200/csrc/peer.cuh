// Peer-memory all-reduce shared by the kernels that end in a collective (box query, stream stats).
#pragma once
#include "common.cuh"

namespace dcarl {

// ---- one-shot sum over the ranks of an NVLink / NVSwitch domain, through peer-mapped memory -----------
// Every rank owns a workspace {slots [2][world][n_max] doubles, flags [2][world] u32, epoch u32} that
// every other rank has mapped (CUDA IPC / symmetric memory; the host passes the mapped pointers).  Call e
// (epoch counter, kept on the device so that the launch can sit in a CUDA graph) uses parity e & 1:
//   1. write my n doubles into slot [parity][my rank] of EVERY rank (peer stores over NVLink),
//   2. system fence, then raise flag [parity][my rank] = e on every rank,
//   3. wait until all world flags of MY workspace read e,
//   4. add the world slots in rank order: every rank computes bit-identical sums.
// A rank can only be one call ahead of the slowest one (it needs everybody's flag to finish a call),
// so two parities suffice.  Messages here are <= 48 KB: latency, not bandwidth.
struct PeerCtx {
    double *slots[DCARL_MAX_PEERS];
    unsigned *flags[DCARL_MAX_PEERS];
    unsigned *epoch;
    int rank, world, n_max;
};

__device__ __forceinline__ void st_release_sys(unsigned *p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned *p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ double ld_relaxed_sys(const double *p) {
    double v;
    asm volatile("ld.relaxed.sys.global.f64 %0, [%1];" : "=d"(v) : "l"(p) : "memory");
    return v;
}

// Waits until *f has reached call `e`; gives up after kPeerWaitNs (a rank that died would otherwise leave the others
// spinning for ever, and the GPU with them).  Returns false on a timeout.
constexpr unsigned long long kPeerWaitNs = 20ull * 1000 * 1000 * 1000;
__device__ __forceinline__ bool peer_wait_flag(const unsigned *f, unsigned e) {
    unsigned long long t0 = 0;
    for (unsigned spins = 0; (int)(ld_acquire_sys(f) - e) < 0; ++spins) {
        __nanosleep(20);
        if ((spins & 1023u) == 1023u) {
            unsigned long long t;
            asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
            if (t0 == 0) t0 = t;
            else if (t - t0 > kPeerWaitNs) return false;
        }
    }
    return true;
}

// All threads of ONE CTA call this.  src: this rank's n doubles (global or shared memory);
// dst: global, receives the rank-ordered sum (may alias src).
__device__ __forceinline__ void peer_allreduce_cta(const double *src, double *dst, int n, const PeerCtx &c) {
    __shared__ unsigned s_epoch, s_lost;
    const int tid = threadIdx.x, nth = blockDim.x;
    if (tid == 0) s_lost = 0u;
    if (tid == 0) { s_epoch = *c.epoch + 1u; *c.epoch = s_epoch; }
    __syncthreads();
    const unsigned e = s_epoch;
    const size_t par = e & 1u;
    for (int q = 0; q < c.world; ++q) {
        double *d = c.slots[q] + (par * c.world + c.rank) * (size_t)c.n_max;
        for (int i = tid; i < n; i += nth) d[i] = src[i];
    }
    __threadfence_system();
    __syncthreads();
    if (tid < c.world) st_release_sys(c.flags[tid] + par * c.world + c.rank, e);
    if (tid < c.world && !peer_wait_flag(c.flags[c.rank] + par * c.world + tid, e)) s_lost = 1u;
    __syncthreads();
    const double *mine = c.slots[c.rank] + par * c.world * (size_t)c.n_max;
    const bool lost = s_lost != 0u;             // a peer never arrived: NaN, so that nothing downstream looks valid
    for (int i = tid; i < n; i += nth) {
        double sum = 0.0;
        for (int r = 0; r < c.world; ++r) sum += ld_relaxed_sys(mine + (size_t)r * c.n_max + i);
        dst[i] = lost ? __longlong_as_double(0x7ff8000000000000ll) : sum;
    }
}


// Host side: validates a dcarl_peer_ctx and turns it into the device-side view.
inline int make_peer_ctx(const dcarl_peer_ctx *h, int n, int arg_index, PeerCtx &c) {
    c.world = 1; c.rank = 0; c.n_max = 0; c.epoch = nullptr;
    for (int r = 0; r < DCARL_MAX_PEERS; ++r) { c.slots[r] = nullptr; c.flags[r] = nullptr; }
    if (!h || h->world <= 1) return 0;
    if (h->world > DCARL_MAX_PEERS) return bad_arg(arg_index, "peer context: world > DCARL_MAX_PEERS");
    if (h->rank < 0 || h->rank >= h->world) return bad_arg(arg_index, "peer context: rank outside [0, world)");
    if (n > h->n_max) return bad_arg(arg_index, "peer context: message longer than n_max");
    if (!h->epoch) return bad_arg(arg_index, "peer context: epoch is null");
    for (int r = 0; r < h->world; ++r)
        if (!h->slots[r] || !h->flags[r]) return bad_arg(arg_index, "peer context: unmapped peer pointer");
    c.world = h->world; c.rank = h->rank; c.n_max = h->n_max;
    c.epoch = reinterpret_cast<unsigned *>(h->epoch);
    for (int r = 0; r < h->world; ++r) {
        c.slots[r] = reinterpret_cast<double *>(h->slots[r]);
        c.flags[r] = reinterpret_cast<unsigned *>(h->flags[r]);
    }
    return 0;
}

}  // namespace dcarl
