#!/usr/bin/env python
"""bench.py — scenario-steps/sec of the DCARL hot path on B200 (BASELINE.json metric).

A "step" is one pass of the streaming confidence-bound estimator + argmax selector over one
batch of synthetic input: BASELINE config 2, "Simulation_2 test_DCARL.py, 4096 parallel scenario
instances" — 4096 instances x 20 000 rows, S=20 states, A=11 actions, float32 rows in, float32
value + int32 argmax + float32 overall_value out (24 algorithmic bytes per scenario-step), plus the
summary-statistics reduction (and, for N > 1, its all-reduce).  N > 1 is weak scaling: every rank
runs its own 4096 instances, no data-path collective.

    python bench.py [--gpus N] [--steps K] [--warmup W]            # this repo's CUDA path
    python bench.py --impl reference [...]                         # the reference's CPU path (port)

Prints ONE JSON line (rank 0).  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

S, A = 20, 11
N_INST, N_STEPS = 4096, 20000
BYTES_PER_STEP = 12 + 8 + 4          # idx+act+R read, value+act written, overall written
METRIC = "scenario-steps/sec (rollout+confidence)"
WORKLOAD = "config2: Simulation_2 stream, 4096 instances x 20000 steps, S=20, A=11, f32 rows, overall_value on"


def _cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


# ---- synthetic workload ---------------------------------------------------
def synth_host(n_inst, n_steps, seed):
    rng = np.random.default_rng(seed)
    q = rng.uniform(-50, 100, (n_inst, S, A))
    idx = np.floor(rng.normal(3, 1, (n_inst, n_steps)) / 6 * S).astype(np.int64)
    bad = (idx < 0) | (idx >= S)
    while bad.any():
        idx[bad] = np.floor(rng.normal(3, 1, int(bad.sum())) / 6 * S).astype(np.int64)
        bad = (idx < 0) | (idx >= S)
    act = rng.integers(0, A, (n_inst, n_steps))
    R = (q[np.arange(n_inst)[:, None], idx, act] + 50 * rng.standard_normal((n_inst, n_steps))).astype(np.float32)
    return idx.astype(np.int32), act.astype(np.int32), R


def synth_device(n_inst, n_steps, seed, device, S=S, A=A, live=None):
    """Sim-2-shaped rows by default; S=1, A=30, live=11 gives the Simulation_1 shape (11 live actions)."""
    import torch
    g = torch.Generator(device=device).manual_seed(seed)
    q = torch.rand((n_inst, S * A), generator=g, device=device) * 150 - 50
    idx = torch.empty((n_inst, n_steps), dtype=torch.int32, device=device)
    act = torch.empty((n_inst, n_steps), dtype=torch.int32, device=device)
    R = torch.empty((n_inst, n_steps), dtype=torch.float32, device=device)
    chunk = 512
    for lo in range(0, n_inst, chunk):
        hi = min(lo + chunk, n_inst)
        z = torch.randn((hi - lo, n_steps), generator=g, device=device)
        bad = (z < -3) | (z >= 3)
        while bool(bad.any()):                       # redraw the 0.27 % that fall outside [0, S)
            z[bad] = torch.randn(int(bad.sum()), generator=g, device=device)
            bad = (z < -3) | (z >= 3)
        i = torch.floor((3 + z) / 6 * S).clamp_(0, S - 1).long()
        a = torch.randint(0, live or A, (hi - lo, n_steps), generator=g, device=device)
        R[lo:hi] = q[lo:hi].gather(1, i * A + a) + 50 * torch.randn((hi - lo, n_steps), generator=g, device=device)
        idx[lo:hi] = i.to(torch.int32)
        act[lo:hi] = a.to(torch.int32)
    return idx, act, R


# ---- CPU legs (the oracle is used here as the thing timed: bench.py's cpu_baseline / reference arm) --
def _literal_worker(args):
    from oracle import pyref
    idx, act, R = args
    data = np.stack([idx, np.zeros(len(idx)), act, R], axis=1).astype(np.float64)
    out = pyref.stream(data, S, A)
    return out["act"]


def cpu_literal(idx, act, R, procs):
    """The reference's own algorithm (literal Python/NumPy restatement of test_DCARL.py:72-105),
    one instance per task over `procs` processes.  Returns (steps/s, per-instance argmax traces)."""
    import multiprocessing as mp
    jobs = [(idx[i], act[i], R[i]) for i in range(len(idx))]
    ctx = mp.get_context("fork")
    with ctx.Pool(procs) as pool:
        pool.map(_literal_worker, [(j[0][:64], j[1][:64], j[2][:64]) for j in jobs[:procs]])   # fork/import warm-up
        t0 = time.perf_counter()
        acts = pool.map(_literal_worker, jobs, chunksize=1)
        dt = time.perf_counter() - t0
    return idx.size / dt, np.stack(acts), dt


def cpu_c_port(idx, act, R, threads):
    """Same algorithm as plain C (oracle/dcarl_oracle.c, bit-exact with numpy), ctypes releases the GIL."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle import cport
    cport.lib()
    R64 = R.astype(np.float64)
    parts = np.array_split(np.arange(len(idx)), max(1, min(threads, len(idx))))
    t0 = time.perf_counter()
    with ThreadPoolExecutor(threads) as ex:
        outs = list(ex.map(lambda p: cport.stream(idx[p], act[p], R64[p], S, A), parts))
    dt = time.perf_counter() - t0
    return idx.size / dt, np.concatenate([o["act"] for o in outs]), dt


# ---- clocks ----------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    def __init__(self, index, period=0.004):
        super().__init__(daemon=True)
        self.period, self.samples, self.reasons, self.sm_max, self.ok = period, [], set(), None, False
        self._stop_ev = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.sm_max = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.ok = True
        except Exception as e:   # noqa: BLE001
            self.err = repr(e)

    def run(self):
        if not self.ok:
            return
        nv = self.nv
        names = {"hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                 "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                 "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                 "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
                 "hw_power_brake": getattr(nv, "nvmlClocksThrottleReasonHwPowerBrakeSlowdown", 0x80)}
        while not self._stop_ev.is_set():
            try:
                self.samples.append(int(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                r = int(nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                for k, bit in names.items():
                    if r & bit:
                        self.reasons.add(k)
            except Exception:   # noqa: BLE001
                pass
            self._stop_ev.wait(self.period)

    def stop(self):
        self._stop_ev.set()
        self.join(timeout=2)

    def summary(self):
        if not self.ok or not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": self.sm_max, "reasons": sorted(self.reasons), "samples": 0}
        return {"sm_mhz": float(np.median(self.samples)), "sm_max_mhz": self.sm_max,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


def physical_gpu_index(local):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local])
        except (ValueError, IndexError):
            return local
    return local


def measured_peak():
    try:
        return float(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]), "measured"
    except Exception:   # noqa: BLE001
        return 6650.0, "fallback"


def ncu_traffic():
    """dram bytes per launch of the stream kernel from the committed ncu summary, if any."""
    try:
        d = json.load(open(os.path.join(ROOT, "profiles", "stream_rows_ncu_summary.json")))
        return d.get("dram_bytes_per_launch")
    except Exception:   # noqa: BLE001
        return None


# ---- this repo's arm ---------------------------------------------------------------------------------
def run_ours(args):
    import torch
    import dcarl_b200 as dc
    from dcarl_b200 import dist as ddist
    rank, world, local = ddist.init_from_env()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — this arm has no CPU fallback (use --impl reference for the CPU path)")
    dev = torch.device(f"cuda:{local}")
    torch.cuda.set_device(dev)
    n_inst, n_steps = args.n_inst, args.n_steps
    idx, act, R = synth_device(n_inst, n_steps, 1000 + rank, dev)
    plan = dc.StreamPlan(n_inst, n_steps, S, A, r_dtype=torch.float32, out_dtype=torch.float32,
                         want_overall=True, want_stats=True)
    stream = torch.cuda.current_stream()

    def step():
        res = plan.run_device(idx, act, R)
        ddist.allreduce_stats(res.stats)
        return res

    for _ in range(max(args.warmup, 3)):
        step()
    plan.raise_on_bad_rows()
    sampler = ClockSampler(physical_gpu_index(local))
    ddist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    for i in range(args.steps):
        # events bracket the dominant kernel (k_stream_rows) on the launching stream
        plan.act_hist.zero_(); plan.bad_rows.zero_()
        k_ev[i][0].record(stream)
        dc._abi.check(dc._abi.lib().dcarl_stream_bounds(
            dc._dev.ptr(idx), dc._dev.ptr(act), dc._dev.ptr(R), 0, n_inst, n_steps, S, A, plan.params,
            dc._dev.ptr(plan.value), 0, dc._dev.ptr(plan.act), dc._dev.ptr(plan.activation_step),
            dc._dev.ptr(plan.overall), dc._dev.ptr(plan.act_hist), dc._dev.ptr(plan.bad_rows),
            dc._dev.ptr(plan.ws), plan._ws_bytes, 0, dc._dev.stream_handle(stream)))
        k_ev[i][1].record(stream)
        dc._abi.check(dc._abi.lib().dcarl_stream_stats(
            dc._dev.ptr(plan.value), 0, dc._dev.ptr(plan.overall), dc._dev.ptr(plan.activation_step),
            dc._dev.ptr(plan.act_hist), n_inst, n_steps, S, A, dc._dev.ptr(plan.stats), dc._dev.stream_handle(stream)))
        ddist.allreduce_stats(plan.stats)
    e1.record(stream)
    torch.cuda.synchronize()
    ddist.barrier()
    ms_total = e0.elapsed_time(e1)
    if len(sampler.samples) < 5:            # a short timed region: keep sampling over an identical probe loop
        # rank-local on purpose: the number of iterations differs per rank, so NO collective in here
        t_end = time.perf_counter() + 0.5
        while time.perf_counter() < t_end:
            plan.run_device(idx, act, R)
        torch.cuda.synchronize()
        clocks_from = "timed region + identical 0.5 s probe loop"
    else:
        clocks_from = "timed region"
    sampler.stop()
    k_ms = float(np.mean([a.elapsed_time(b) for a, b in k_ev]))
    t = torch.tensor([ms_total, k_ms], dtype=torch.float64, device=dev)
    ddist.allreduce_max(t)
    ms_total, k_ms_max = float(t[0]), float(t[1])
    ms_per_step = ms_total / args.steps
    total_steps = world * n_inst * n_steps
    value = total_steps / (ms_per_step * 1e-3)
    stats = plan.stats.cpu().numpy()

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ---------------------
    pin = plan.pinned()
    pin["idx"].copy_(idx); pin["act"].copy_(act); pin["R"].copy_(R)
    torch.cuda.synchronize()
    e2e_steps = max(2, min(args.steps, 5))
    plan.run_pinned()                                   # warm-up (first touch of the pinned pages)
    ddist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        plan.run_pinned()                               # returns after the stream drained
        ddist.allreduce_stats(plan.stats)
    torch.cuda.synchronize()
    e2e_ms = (time.perf_counter() - t0) * 1e3 / e2e_steps
    t = torch.tensor([e2e_ms], dtype=torch.float64, device=dev)
    ddist.allreduce_max(t)
    e2e_ms = float(t[0])
    same = bool(torch.equal(pin["value"], plan.value.cpu())) and bool(torch.equal(pin["oact"], plan.act.cpu()))

    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = n_inst * n_steps * BYTES_PER_STEP / (k_ms_max * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "scenario-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic",
            "config": {"workload": WORKLOAD, "n_inst_per_gpu": n_inst, "n_steps": n_steps, "S": S, "A": A,
                       "io": "f32 rows in, f32 value/overall + i32 argmax out; float64 accumulation and bounds",
                       "cache": "inputs (983 MB) + outputs (983 MB) per step exceed the 126 MB L2",
                       "parallelism": f"instances sharded, {world} rank(s), one stats all-reduce per step"},
            "clocks": dict(sampler.summary(), sampled_over=clocks_from),
            "e2e": {"value": total_steps / (e2e_ms * 1e-3), "unit": "scenario-steps/s",
                    "h2d_bytes_per_step": plan.h2d_bytes(), "d2h_bytes_per_step": plan.d2h_bytes(),
                    "ms_per_step": e2e_ms, "outputs_match_device_run": same,
                    "api": "dcarl_b200.StreamPlan.run_pinned (pinned host in/out, copies in the timed region)"},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "kernel": "k_stream_rows",
                         "kernel_ms": k_ms_max, "algorithmic_bytes_per_step": BYTES_PER_STEP, "peak_source": peak_src},
            "stats_check": {"total_steps": float(stats[4]), "expected": float(total_steps),
                            "activated_pairs": float(stats[2])},
        }
    if rank == 0 and world == 1 and not args.no_cpu:
        out.update(cpu_legs(plan, idx, act, R))
    if rank == 0 and world == 1 and not args.no_secondary:
        del idx, act, R
        plan = None
        torch.cuda.empty_cache()
        try:
            out["secondary"] = secondary(dev, measured_peak()[0])
        except Exception as e:   # noqa: BLE001  (secondary numbers must never cost the headline line)
            out["secondary"] = {"error": repr(e)}
    if rank == 0:
        print(json.dumps(out), flush=True)
    if world > 1:
        import torch.distributed as tdist
        tdist.destroy_process_group()


def secondary(dev, peak):
    """Roofline of the other kernels on the path (BASELINE configs 3 and 4, single-GPU share):
    CUDA-event time of back-to-back launches on inputs larger than L2."""
    import torch
    import dcarl_b200 as dc
    out = {}

    def timed(fn, reps):
        fn(); fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / reps

    # config 4 (one GPU's view): 1e7-row buffer, 256 candidate actions, Q = 1 and Q = 64
    g = torch.Generator(device=dev).manual_seed(3000)
    M, NA = 10_000_000, 256
    dist = torch.tensor(dc.VISITED_STATE_DIST[0], dtype=torch.float64, device=dev)
    centers = torch.randn((4096, 20), generator=g, device=dev) * dist[:20].float()
    buf = dc.BoxBuffer(20, capacity=M)
    st = centers[torch.randint(0, 4096, (M,), generator=g, device=dev)]
    st += torch.randn((M, 20), generator=g, device=dev) * (0.25 * dist[:20].float())
    a = torch.randint(0, NA, (M, 1), generator=g, device=dev).float()
    ret = -(0.95 ** torch.randint(0, 10, (M,), generator=g, device=dev).float()) * (torch.rand(M, generator=g, device=dev) < 0.5)
    buf.append(torch.cat([st, a], 1), ret)
    del st, a, ret
    for Q in (1, 64):
        q = centers[:Q].double().contiguous()
        mom = torch.empty((Q, NA, 3), dtype=torch.float64, device=dev)
        ms = timed(lambda: buf.moments(q, NA, out=mom), 20)
        gbs = M * 88 / (ms * 1e-3) / 1e9
        out[f"box_moments_Q{Q}"] = {"rows": M, "n_actions": NA, "ms": ms, "rows_per_s": M / (ms * 1e-3),
                                    "queries_per_s": Q / (ms * 1e-3), "achieved_GBs": gbs, "frac": gbs / peak,
                                    "algorithmic_bytes_per_row": 88, "matched_rows": float(mom[:, :, 0].sum())}
    run = buf.query_graph(1, NA)
    q1 = centers[:1].double().contiguous()
    ms = timed(lambda: run(q1), 50)
    out["box_query_graph_Q1"] = {"rows": M, "ms": ms, "queries_per_s": 1 / (ms * 1e-3), "achieved_GBs": M * 88 / (ms * 1e-3) / 1e9,
                                 "frac": M * 88 / (ms * 1e-3) / 1e9 / peak,
                                 "note": "whole act_test query as one CUDA-graph replay: query upload + zero + prepare + scan + select"}
    del buf, run
    # config 3 mode A: 65 536 episodes x 200 draws, deviates supplied (16 B/row) and on-device Philox (12 B/row)
    n = 65536 * 200
    qs = torch.rand((20, 11), generator=g, device=dev) * 150 - 50
    zs, zr = torch.randn(n, generator=g, device=dev), torch.randn(n, generator=g, device=dev)
    ac = torch.randint(0, 11, (n,), generator=g, device=dev, dtype=torch.int32)
    ms = timed(lambda: dc.sample_returns(qs, zs, ac, zr), 50)
    out["sample_returns"] = {"rows": n, "ms": ms, "rows_per_s": n / (ms * 1e-3), "achieved_GBs": n * 20 / (ms * 1e-3) / 1e9,
                             "frac": n * 20 / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_row": 20,
                             "note": "12 B read (z_state, z_ret, act) + 8 B written (idx, R); includes two torch.empty per call"}
    ms = timed(lambda: dc.sample_returns_philox(qs, n, 7), 50)
    out["sample_returns_philox"] = {"rows": n, "ms": ms, "rows_per_s": n / (ms * 1e-3),
                                    "achieved_GBs": n * 12 / (ms * 1e-3) / 1e9, "frac": n * 12 / (ms * 1e-3) / 1e9 / peak,
                                    "algorithmic_bytes_per_row": 12}
    # config 3 mode B: constant-velocity rollout surrogate, 65 536 episodes x 200 steps (not bandwidth bound)
    ne = 65536
    s0 = torch.zeros((ne, 20), device=dev)
    s0[:, 2] = torch.rand(ne, generator=g, device=dev) * 6 + 2
    for j in range(3):
        b = 5 + 5 * j
        s0[:, b] = torch.rand(ne, generator=g, device=dev) * 60 - 30
        s0[:, b + 1] = torch.rand(ne, generator=g, device=dev) * 12 - 6
        s0[:, b + 2] = torch.rand(ne, generator=g, device=dev) * 10 - 5
        s0[:, b + 3] = torch.rand(ne, generator=g, device=dev) * 2 - 1
    ra = torch.randint(0, 11, (ne,), generator=g, device=dev, dtype=torch.int32)
    ms = timed(lambda: dc.rollout_cv(s0, ra), 50)
    _, steps_run, _ = dc.rollout_cv(s0, ra)
    out["rollout_cv"] = {"episodes": ne, "ms": ms, "kinematic_steps_per_s": float(steps_run.sum()) / (ms * 1e-3),
                         "episodes_per_s": ne / (ms * 1e-3), "note": "96 B per episode: latency/compute bound, no roofline fraction"}
    del s0, ra, zs, zr, ac
    out["frenet_plan"] = frenet_leg(dev)
    return out


def synth_frenet(n, n_paths, seed, max_obs=4):
    """Planner inputs for the Frenet leg: gently curved lanes with 2 m knots, a start state along a
    random lane, slow obstacles on the corridor ahead (vectorised numpy; same recipe as the tests)."""
    rng = np.random.default_rng(seed)
    knots, off = [], [0]
    for _ in range(n_paths):
        L = rng.uniform(80, 220)
        sk = np.linspace(0, L, int(round(L / 2)))
        amp, w, ph, th = rng.uniform(0, 8), rng.uniform(15, 60), rng.uniform(0, 6.28), rng.uniform(-3.14, 3.14)
        x, y = sk, amp * np.sin(sk / w + ph)
        knots.append(np.stack([x * np.cos(th) - y * np.sin(th), x * np.sin(th) + y * np.cos(th)], 1) + rng.uniform(-100, 100, 2))
        off.append(off[-1] + len(sk))
    off = np.array(off, np.int64)
    pid = rng.integers(0, n_paths, n).astype(np.int32)
    nk = (off[1:] - off[:-1])[pid]
    j = (off[:-1][pid] + (rng.uniform(0.02, 0.7, n) * (nk - 2)).astype(np.int64))
    allk = np.concatenate(knots)
    u = allk[j + 1] - allk[j]
    seglen = np.hypot(u[:, 0], u[:, 1])
    u /= seglen[:, None]
    nrm = np.stack([-u[:, 1], u[:, 0]], 1)
    c_d = rng.uniform(-3, 3, n)
    s0 = (j - off[:-1][pid]) * 2.0 + 1.0                       # knots are ~2 m apart
    start = np.stack([s0, c_d, rng.normal(0, 0.4, n), np.zeros(n), rng.uniform(0.2, 12.0, n)], 1)
    ego = allk[j] + u * 1.0 + nrm * c_d[:, None]
    ahead, side = rng.uniform(2, 30, (n, max_obs)), rng.uniform(-5, 5, (n, max_obs))
    pos = ego[:, None, :] + u[:, None, :] * ahead[..., None] + nrm[:, None, :] * (side - c_d[:, None])[..., None]
    v, yaw = rng.uniform(0, 2, (n, max_obs)), rng.uniform(-3.14, 3.14, (n, max_obs))
    obs = np.concatenate([pos, (v * np.cos(yaw))[..., None], (v * np.sin(yaw))[..., None], yaw[..., None]], 2)
    n_obs = rng.integers(0, max_obs + 1, n).astype(np.int32)
    return knots, off, pid, start, ego, obs, n_obs


def frenet_leg(dev):
    """SURVEY row f3 at BASELINE config-5's scenario count: one planning tick (10 candidates x 14
    points, collision-checked pick) for 262 144 scenarios per launch, device-resident inputs; the
    oracle's C restatement of the reference planner on the host cores beside it, on a sample."""
    import torch
    from concurrent.futures import ThreadPoolExecutor
    import dcarl_b200 as dc
    from oracle import cport
    n, n_paths = 262144, 64
    knots, off, pid, start, ego, obs, n_obs = synth_frenet(n, n_paths, 5000)
    routes = dc.Routes(knots)
    p = dc.frenet_params()
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    d = (t(pid, torch.int32), t(start, torch.float64), t(ego, torch.float64), t(obs, torch.float64), t(n_obs, torch.int32))
    res = dc.frenet_plan(routes, *d, p)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    reps = 10
    e0.record()
    for _ in range(reps):
        res = dc.frenet_plan(routes, *d, p)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    C, T = p.n_d * p.n_T * p.n_v, max(p.n_t[:p.n_T])
    # CPU: the oracle on a sample, all cores (ctypes releases the GIL)
    cores = _cores()
    ns = min(n, 8192 * cores)
    coef = np.concatenate([cport.spline_build(k) for k in knots])
    po = cport.frenet_params()
    parts = np.array_split(np.arange(ns), cores)
    t0 = time.perf_counter()
    with ThreadPoolExecutor(cores) as ex:
        outs = list(ex.map(lambda q: cport.frenet_plan(coef, off, pid[q], start[q], ego[q], obs[q], n_obs[q], po, want_paths=False)["index"], parts))
    dt = time.perf_counter() - t0
    cpu_idx = np.concatenate(outs)
    gpu_idx = res.index[:ns].cpu().numpy()
    return {"scenarios": n, "candidates": C, "points_per_candidate": T, "ms": ms, "scenarios_per_s": n / (ms * 1e-3),
            "candidate_points_per_s": n * C * T / (ms * 1e-3),
            "io_bytes_per_scenario": 5 * 8 + 2 * 8 + obs.shape[1] * 40 + 8 + 4 + C * 13 + T * 24 + 4,
            "selected_index_histogram": np.bincount(res.index.cpu().numpy(), minlength=C + 1).tolist(),
            "cpu_oracle": {"scenarios_per_s": ns / dt, "cores": cores, "sample": f"{ns} scenarios, C restatement of the reference planner, {dt:.2f} s"},
            "index_mismatches_vs_cpu_sample": int((cpu_idx != gpu_idx).sum()),
            "note": "compute-bound float64 (one warp per scenario): no bandwidth roofline; the reference's Python planner "
                    "takes ~10 ms per scenario per core (measured in the build container, not on this box)"}


def cpu_legs(plan, idx, act, R):
    """cpu_baseline: the reference's algorithm on the host cores, on a bounded sample of the SAME
    inputs (first instances of rank 0), cross-checked against the GPU's argmax traces."""
    cores = _cores()
    procs = min(cores, 64)
    n_lit = min(idx.shape[0], 12 * procs)       # ~10-15 s of CPU work
    hi, ha, hr = idx[:n_lit].cpu().numpy(), act[:n_lit].cpu().numpy(), R[:n_lit].cpu().numpy()
    gpu_act = plan.act[:n_lit].cpu().numpy()
    v_lit, acts_lit, dt_lit = cpu_literal(hi, ha, hr, procs)
    n_c = idx.shape[0]          # the C port is fast: time the whole batch
    ci, ca, cr = idx[:n_c].cpu().numpy(), act[:n_c].cpu().numpy(), R[:n_c].cpu().numpy()
    v_c, acts_c, dt_c = cpu_c_port(ci, ca, cr, cores)
    mism = int((acts_lit != gpu_act).sum()) + int((acts_c != plan.act[:n_c].cpu().numpy()).sum())
    return {
        "cpu_baseline": {"value": v_lit, "unit": "scenario-steps/s", "cores": procs, "kind": "port",
                         "sample": f"{n_lit} of the {idx.shape[0]} instances x {idx.shape[1]} steps, literal "
                                   f"Python/NumPy restatement of test_DCARL.py:72-105, {procs} processes, {dt_lit:.1f} s"},
        "cpu_baseline_c": {"value": v_c, "unit": "scenario-steps/s", "cores": cores, "kind": "port",
                           "sample": f"{n_c} instances, plain-C restatement (bit-exact with numpy), {cores} threads, {dt_c:.1f} s"},
        "parity_vs_cpu_sample": {"argmax_mismatches": mism, "rows_checked": int(acts_lit.size + acts_c.size)},
    }


# ---- the reference arm: the reference's own CPU implementation of the path --------------------------
def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = _cores()
    procs = min(cores, 64)
    # ~3.5 s per 20 000-step instance per core: size the per-step sample so K+W steps end in minutes
    per_core = int(min(4, max(1, 150.0 / (args.steps + 1) / 3.5)))
    n_lit = per_core * procs
    idx, act, R = synth_host(n_lit, args.n_steps, 1000)
    warm = max(args.warmup, 0)
    for _ in range(min(warm, 1)):
        cpu_literal(idx, act, R, procs)
    times = []
    for _ in range(args.steps):
        v, _, dt = cpu_literal(idx, act, R, procs)
        times.append(dt)
    ms = float(np.mean(times)) * 1e3
    value = idx.size / (ms * 1e-3)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scenario-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": min(warm, 1), "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "n_steps": args.n_steps, "S": S, "A": A,
                   "note": "the reference is single-process Python/NumPy and cannot travel to the GPU box; this arm "
                           "times its literal restatement (oracle/pyref.py, pinned bit-exact against the reference's "
                           "outputs) on all host cores, one instance per task"},
        "cpu_baseline": {"value": value, "unit": "scenario-steps/s", "cores": procs, "kind": "port",
                         "sample": f"{n_lit} instances x {args.n_steps} steps per step"},
        "e2e": {"value": value, "unit": "scenario-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--n-inst", type=int, default=N_INST, help="instances per GPU (default: config 2)")
    ap.add_argument("--n-steps", type=int, default=N_STEPS)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline legs")
    ap.add_argument("--no-secondary", action="store_true", help="skip the box / sampler / rollout rooflines")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else 50
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
