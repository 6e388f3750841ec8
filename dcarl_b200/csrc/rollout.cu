// Constant-velocity rollout surrogate.
//
// The reference's episodes are stepped by the external CARLA simulator
// (Data_From_Carla/Test_Scenarios/TestScenario_Town03.py:381-382), which cannot run here;
// this kernel restates, for synthetic scenario batches, the in-repo pieces of that loop:
//   constant-velocity prediction   Data_From_Carla/Agent/zzz/predict.py:96-100
//   collision test d^2 <= r^2      predict.py:52-59
//   step reward 0.1*sqrt(v)        TestScenario_Town03.py:403-404
//   collision -100 & done          TestScenario_Town03.py:407-410
//   pass & done                    TestScenario_Town03.py:414-415
//   episode return = sum(rewards)  Agent/drl_library/dqn/dqn_value_collect.py:119
// One thread per scenario; 80 bytes in, 12 bytes out per episode: not bandwidth bound.
#include "common.cuh"

namespace dcarl {

__global__ void __launch_bounds__(256) k_rollout_cv(const float *__restrict__ state0, const int *__restrict__ action,
                                                     long long n, int steps, float dt, float radius, float pass_x,
                                                     const float *__restrict__ speed_scale, int n_actions,
                                                     float *__restrict__ out_return, int *__restrict__ out_steps,
                                                     int *__restrict__ out_term) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 *s4 = reinterpret_cast<const float4 *>(state0 + i * 20);
    const float4 q0 = s4[0], q1 = s4[1], q2 = s4[2], q3 = s4[3], q4 = s4[4];
    // layout: ego[x y vx vy yaw] obj0[x y vx vy yaw] obj1[...] obj2[...]
    const int a = action[i];
    if ((unsigned)a >= (unsigned)n_actions) {
        out_return[i] = 0.f; out_steps[i] = 0; out_term[i] = -1;   // flagged to the caller
        return;
    }
    const float sc = speed_scale[a];
    float ex = q0.x, ey = q0.y;
    const float evx = q0.z * sc, evy = q0.w * sc;
    float ox[3] = {q1.y, q2.z, q3.w}, oy[3] = {q1.z, q2.w, q4.x};
    const float ovx[3] = {q1.w, q3.x, q4.y}, ovy[3] = {q2.x, q3.y, q4.z};
    const float v = sqrtf(__fadd_rn(__fmul_rn(evx, evx), __fmul_rn(evy, evy)));
    const float step_reward = __fmul_rn(0.1f, sqrtf(v));
    const float r2 = __fmul_rn(radius, radius);
    float ret = 0.f;
    int t, term = 0;
    for (t = 1; t <= steps; ++t) {
        ex = __fmaf_rn(evx, dt, ex);
        ey = __fmaf_rn(evy, dt, ey);
        bool hit = false;
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            ox[j] = __fmaf_rn(ovx[j], dt, ox[j]);
            oy[j] = __fmaf_rn(ovy[j], dt, oy[j]);
            const float dx = __fsub_rn(ex, ox[j]), dy = __fsub_rn(ey, oy[j]);
            const float d2 = __fmaf_rn(dx, dx, __fmul_rn(dy, dy));
            hit = hit || d2 <= r2;
        }
        if (hit) { ret = __fadd_rn(ret, -100.0f); term = 1; break; }
        ret = __fadd_rn(ret, step_reward);
        if (ex > pass_x) { term = 2; break; }
    }
    out_return[i] = ret;
    out_steps[i] = t <= steps ? t : steps;
    out_term[i] = term;
}

}  // namespace dcarl

using namespace dcarl;

extern "C" int dcarl_rollout_cv(const float *state0, const int32_t *action, int64_t n, int32_t steps, float dt,
                                float radius, float pass_x, const float *speed_scale, int32_t n_actions,
                                float *out_return, int32_t *out_steps, int32_t *out_term, dcarl_stream_t stream) {
    if (n < 0) return bad_arg(3, "n < 0");
    if (steps < 0) return bad_arg(4, "steps < 0");
    if (n_actions <= 0) return bad_arg(9, "n_actions <= 0");
    if (n == 0) return 0;
    if (!state0 || !action || !speed_scale) return bad_arg(1, "null input");
    if ((uintptr_t)state0 & 15) return bad_arg(1, "state0 must be 16-byte aligned");
    if (!out_return || !out_steps || !out_term) return bad_arg(10, "null output");
    cudaStream_t st = (cudaStream_t)stream;
    // out-of-range actions are flagged through out_term = -1
    k_rollout_cv<<<(unsigned)((n + 255) / 256), 256, 0, st>>>(state0, action, n, steps, dt, radius, pass_x,
                                                              speed_scale, n_actions, out_return, out_steps,
                                                              out_term);
    DCARL_LAUNCH_CHECK("k_rollout_cv");
    return 0;
}
