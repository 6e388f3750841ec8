// Error plumbing and the small host-only entry points of the C ABI (include/dcarl_b200.h).
#include <cstdarg>
#include <cstdio>
#include <string>

#include "common.cuh"

namespace dcarl {

static thread_local std::string g_err;

void set_error(const char *fmt, ...) {
    char buf[512];
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(buf, sizeof buf, fmt, ap);
    va_end(ap);
    g_err = buf;
}

int bad_arg(int index, const char *what) {
    set_error("invalid argument %d: %s", index, what);
    return -index;
}

int cuda_fail(cudaError_t e, const char *where) {
    set_error("CUDA error %d (%s) at %s", (int)e, cudaGetErrorString(e), where);
    return (int)e;
}

int sm_count() {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) return 148;
    return n;
}

}  // namespace dcarl

extern "C" const char *dcarl_last_error(void) { return dcarl::g_err.c_str(); }

extern "C" int dcarl_abi_version(void) { return DCARL_B200_ABI_VERSION; }

extern "C" int dcarl_device_info(int *sm_count, int *cc_major, int *cc_minor) {
    int dev = 0;
    DCARL_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    DCARL_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (sm_count) *sm_count = prop.multiProcessorCount;
    if (cc_major) *cc_major = prop.major;
    if (cc_minor) *cc_minor = prop.minor;
    return 0;
}

extern "C" void dcarl_default_params(dcarl_params *p) {
    if (!p) return;
    p->alpha = 0.05; p->scale = 150.0; p->cap = 100.0; p->prior_rule = 100.0; p->prior_other = -50.0;
    p->activation_bonus = 0.9; p->n_thres = 10; p->rule_act = 0;
}

extern "C" void dcarl_default_rls_params(dcarl_rls_params *p) {
    if (!p) return;
    p->rule_mean_gate = -0.1; p->confidence_thres = 0.5; p->visited_times_thres = 10; p->min_candidate_visits = 5;
}
