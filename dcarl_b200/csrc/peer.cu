// Multi-GPU entry points: the peer-memory all-reduce as a kernel of its own, the workspace layout, and the
// NCCL form of the same sum (SURVEY 8b "dcarl_allreduce_stats").  The path's only collective is a sum of
// a summary vector of at most a few thousand doubles per batch.
#include <dlfcn.h>

#include "peer.cuh"

namespace dcarl {

__global__ void __launch_bounds__(256) k_peer_allreduce(double *data, int n, const PeerCtx c) {
    peer_allreduce_cta(data, data, n, c);
}

}  // namespace dcarl

using namespace dcarl;

extern "C" void dcarl_peer_ws_layout(int32_t world, int32_t n_max, size_t *slots_off, size_t *flags_off,
                                     size_t *epoch_off) {
    const size_t w = world > 0 ? (size_t)world : 1, n = n_max > 0 ? (size_t)n_max : 0;
    const size_t slots = 2 * w * n * sizeof(double);
    const size_t flags = (slots + 127) & ~(size_t)127;          // flags on their own 128-byte lines
    const size_t epoch = flags + ((2 * w * sizeof(unsigned) + 127) & ~(size_t)127);
    if (slots_off) *slots_off = 0;
    if (flags_off) *flags_off = flags;
    if (epoch_off) *epoch_off = epoch;
}

extern "C" size_t dcarl_peer_ws_bytes(int32_t world, int32_t n_max) {
    size_t e;
    dcarl_peer_ws_layout(world, n_max, nullptr, nullptr, &e);
    return e + 128;
}

extern "C" int dcarl_peer_allreduce(double *data, int32_t n, const dcarl_peer_ctx *peers, dcarl_stream_t stream) {
    if (n < 0) return bad_arg(2, "n < 0");
    if (n == 0) return 0;
    if (!data) return bad_arg(1, "data is null");
    PeerCtx c;
    if (int rc = make_peer_ctx(peers, n, 3, c)) return rc;
    if (c.world <= 1) return 0;                                  // nothing to add
    k_peer_allreduce<<<1, 256, 0, (cudaStream_t)stream>>>(data, n, c);
    DCARL_LAUNCH_CHECK("k_peer_allreduce");
    return 0;
}

// ---- NCCL, resolved at run time (no link-time dependency: the package also loads on boxes without it) ----
namespace {
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
nccl_allreduce_fn resolve_nccl() {
    static nccl_allreduce_fn fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        // the process usually has torch's NCCL loaded already: look there first
        void *sym = dlsym(RTLD_DEFAULT, "ncclAllReduce");
        if (!sym) {
            void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
            if (h) sym = dlsym(h, "ncclAllReduce");
        }
        fn = reinterpret_cast<nccl_allreduce_fn>(sym);
    }
    return fn;
}
}  // namespace

extern "C" int dcarl_allreduce_stats(double *stats, int32_t n, void *comm, dcarl_stream_t stream) {
    if (n < 0) return bad_arg(2, "n < 0");
    if (n == 0) return 0;
    if (!stats) return bad_arg(1, "stats is null");
    if (!comm) return bad_arg(3, "comm is null");
    nccl_allreduce_fn fn = resolve_nccl();
    if (!fn) {
        set_error("libnccl.so.2 (ncclAllReduce) could not be resolved in this process");
        return -3;
    }
    // ncclDouble = 8, ncclSum = 0 (nccl.h; stable across NCCL 2.x)
    const int rc = fn(stats, stats, (size_t)n, 8, 0, comm, (cudaStream_t)stream);
    if (rc != 0) {
        set_error("ncclAllReduce returned %d", rc);
        return 1000 + rc;
    }
    return 0;
}
