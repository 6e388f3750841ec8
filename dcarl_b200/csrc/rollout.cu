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
#include <math_constants.h>

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

// ---- candidate-trajectory rollout: the return of every action of a scenario ---------------------------
// The reference's collector (Data_From_Carla/Agent/drl_library/dqn/dqn_value_collect.py:86-145, "DVC")
// runs one CARLA episode per action: at the trigger it locks in candidate `used_action` of the planner
// (DVC:92-101; 0 = the brake / backup trajectory, c+1 = candidate c, JTP:132-140), follows that ONE
// trajectory to the end of the episode (DVC:106-113), sums the step rewards (DVC:119) and records
// (state, action, episode return) (DVC:128-137).  CARLA cannot run here, so the episode is idealised
// kinematically: the ego is AT point k of the locked trajectory at time k*DT (the trajectory's own time
// grid, JTP:308), its speed is the distance between consecutive points / DT, the other vehicles move at
// constant velocity (PRD:96-100).  Everything else is the reference's rule:
//   step reward 0.1*sqrt(v) per simulator tick            TestScenario_Town03.py ("TS"):403-404
//   collision: reward = -100, done                        TS:407-410; the collision predicate is the
//       planner's own, predict.check_collision PRD:49-59 — front/back circles of the nearest
//       OBSTACLES_CONSIDERED vehicles (PRD:64-110), d^2 <= radius^2 at t = 2, 4, ... <
//       min(len(fp.x) - 1, len(path.t) - 1) — evaluated in time order, the first hit ends the episode
//   pass (y < pass_y): done, the step reward stands       TS:414-415, :196-201
//   stuck (speed < 0.1 for more than stay_thres seconds, counted from the last moving step or the first
//       step): reward = 0, done                           TS:419-422, :182-194
//   the three are applied in that order, `pass` shadowing `stuck` (TS:414-419 if / elif)
// The brake action keeps the ego at the first point with speed 0 (the backup trajectory is the cheapest
// candidate with all desired speeds zeroed, JTP:147-150) over that candidate's horizon.
// One warp per scenario, one lane per action.
struct RolloutCandArgs {
    const double *paths;      // [n, C, Tm, 2]
    const int *n_valid;       // [n, C]
    const double *cost;       // [n, C]
    const double *ego_xy;     // [n, 2]
    const double *obs;        // [n, max_obs, 5]
    const int *n_obs;         // [n]
    long long n;
    int C, Tm, max_obs;
    dcarl_rollout_params p;
    double *out_return;       // [n, C+1]
    int *out_steps, *out_term;
};

__global__ void __launch_bounds__(128) k_rollout_candidates(const RolloutCandArgs a) {
    __shared__ double circ_s[4][8 * 16];          // per warp: {x0, y0, vx, vy} of up to 32 circles
    const int lane = lane_id();
    const int w = threadIdx.x >> 5;
    const long long sc = (long long)blockIdx.x * 4 + w;
    if (sc >= a.n) return;
    double *circ = circ_s[w];
    const dcarl_rollout_params &p = a.p;
    const int C = a.C, Tm = a.Tm;

    // PRD:64-110: the nearest obstacles (stable order by distance to the ego), a front and a back circle each
    const int no = min(min(a.n_obs[sc], a.max_obs), 32);
    const int nuse = min(min(no, p.obstacles_considered), 16);
    {
        double ox = 0.0, oy = 0.0, ovx = 0.0, ovy = 0.0, oyaw = 0.0;
        if (lane < no) {
            const double *o = a.obs + ((size_t)sc * a.max_obs + lane) * 5;
            ox = o[0]; oy = o[1]; ovx = o[2]; ovy = o[3]; oyaw = o[4];
        }
        const double dx = ox - a.ego_xy[2 * sc], dy = oy - a.ego_xy[2 * sc + 1];
        const double dist = lane < no ? sqrt(dx * dx + dy * dy) : CUDART_INF;
        int rank = 0;
        for (int j = 0; j < no; ++j) {
            const double dj = __shfl_sync(FULL_MASK, dist, j);
            rank += (dj < dist || (dj == dist && j < lane)) ? 1 : 0;
        }
        if (lane < no && rank < nuse) {
            double sn, cs;
            sincos(oyaw, &sn, &cs);
            const double gx = cs * p.move_gap, gy = sn * p.move_gap;
            double *f = circ + 8 * rank;
            f[0] = ox + gx; f[1] = oy + gy; f[2] = ovx; f[3] = ovy;
            f[4] = ox - gx; f[5] = oy - gy; f[6] = ovx; f[7] = ovy;
        }
    }
    // the brake action runs over the horizon of the cheapest candidate (sorted() is stable: lowest index on ties)
    int cheapest = 0;
    {
        double best = CUDART_INF;
        for (int c = 0; c < C; ++c) {
            const double cc = a.cost[sc * C + c];
            if (cc < best) { best = cc; cheapest = c; }
        }
    }
    __syncwarp();

    const double r2 = p.robot_radius * p.robot_radius;
    for (int act = lane; act <= C; act += 32) {
        const int c = act == 0 ? cheapest : act - 1;
        const double *xy = a.paths + ((size_t)sc * C + c) * Tm * 2;
        const int n = min(a.n_valid[sc * C + c], Tm);
        const int lim = min(n - 1, p.n_pred - 1);
        double ret = 0.0, t_moving = 0.0;
        int term = 0, steps = n;                                  // 0 = the trajectory ran out
        for (int k = 0; k < n; ++k) {
            double x, y, v;
            if (act == 0) { x = xy[0]; y = xy[1]; v = 0.0; }
            else {
                x = xy[2 * k]; y = xy[2 * k + 1];
                const int k0 = k + 1 < n ? k : k - 1;             // the last point keeps the last speed
                if (n >= 2) {
                    const double ex = xy[2 * k0 + 2] - xy[2 * k0], ey = xy[2 * k0 + 3] - xy[2 * k0 + 1];
                    v = sqrt(ex * ex + ey * ey) / p.dt;
                } else v = 0.0;
            }
            double reward = p.ticks_per_point * (0.1 * sqrt(v));   // TS:403-404
            bool done = false;
            int why = 0;
            if (k >= 2 && (k & 1) == 0 && k < lim) {               // PRD:52-59
                bool hit = false;
                for (int q = 0; q < 2 * nuse; ++q) {
                    const double *f = circ + 4 * q;
                    const double px = f[0] + (k * p.dt) * f[2], py = f[1] + (k * p.dt) * f[3];
                    const double ex = px - x, ey = py - y;
                    hit = hit || (ex * ex + ey * ey <= r2);
                }
                if (hit) { reward = p.collision_reward; done = true; why = 1; }   // TS:407-410
            }
            if (y < p.pass_y) { done = true; if (why == 0) why = 2; }             // TS:414-415
            else {                                                                // TS:419-422, :182-194
                const double now = k * p.dt;
                if (k > 0 && v < p.stuck_speed) {
                    if (now - t_moving > p.stay_thres) { reward = 0.0; done = true; why = 3; }
                } else t_moving = now;
            }
            ret += reward;                                                        // DVC:119
            if (done) { term = why; steps = k + 1; break; }
        }
        a.out_return[sc * (C + 1) + act] = ret;
        a.out_steps[sc * (C + 1) + act] = steps;
        a.out_term[sc * (C + 1) + act] = term;
    }
}

extern "C" void dcarl_default_rollout_params(dcarl_rollout_params *p) {
    if (!p) return;
    p->dt = 0.3; p->robot_radius = 1.0; p->move_gap = 1.0;          // JTP:19, 29, 31
    p->n_pred = 15; p->obstacles_considered = 4;                    // PRD:91 over JTP:20 (MAXT 4.2), JTP:28
    p->ticks_per_point = 6.0;                                       // DT 0.3 s / simulator tick 0.05 s (TS:84)
    p->collision_reward = -100.0;                                   // TS:409
    p->pass_y = 73.7;                                               // TS:198
    p->stuck_speed = 0.1; p->stay_thres = 2.0;                      // TS:188, :182
}

extern "C" int dcarl_rollout_candidates(const double *paths, const int32_t *n_valid, const double *cost,
                                        const double *ego_xy, const double *obs, const int32_t *n_obs,
                                        int32_t max_obs, int64_t n, int32_t C, int32_t Tm,
                                        const dcarl_rollout_params *p, double *out_return, int32_t *out_steps,
                                        int32_t *out_term, dcarl_stream_t stream) {
    if (max_obs < 0) return bad_arg(7, "max_obs < 0");
    if (n < 0) return bad_arg(8, "n < 0");
    if (C <= 0 || C > 32) return bad_arg(9, "C outside [1, 32]");
    if (Tm <= 0 || Tm > 32) return bad_arg(10, "Tm outside [1, 32]");
    if (!p) return bad_arg(11, "params is null");
    if (!(p->dt > 0.0)) return bad_arg(11, "dt <= 0");
    if (p->obstacles_considered < 0 || p->obstacles_considered > 16) return bad_arg(11, "obstacles_considered outside [0, 16]");
    if (n == 0) return 0;
    if (!paths || !n_valid || !cost || !ego_xy || !n_obs) return bad_arg(1, "null input");
    if (max_obs > 0 && !obs) return bad_arg(5, "obs is null");
    if (!out_return || !out_steps || !out_term) return bad_arg(12, "null output");
    RolloutCandArgs a;
    a.paths = paths; a.n_valid = n_valid; a.cost = cost; a.ego_xy = ego_xy; a.obs = obs; a.n_obs = n_obs;
    a.n = n; a.C = C; a.Tm = Tm; a.max_obs = max_obs; a.p = *p;
    a.out_return = out_return; a.out_steps = out_steps; a.out_term = out_term;
    k_rollout_candidates<<<(unsigned)((n + 3) / 4), 128, 0, (cudaStream_t)stream>>>(a);
    DCARL_LAUNCH_CHECK("k_rollout_candidates");
    return 0;
}
