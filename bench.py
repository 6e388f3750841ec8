#!/usr/bin/env python
"""bench.py — scenario-steps/sec of the DCARL hot path on B200 (BASELINE.json metric).

A "step" is one pass of the streaming confidence-bound estimator + argmax selector over one
batch of synthetic input: BASELINE config 2, "Simulation_2 test_DCARL.py, 4096 parallel scenario
instances" — 4096 instances x 20 000 rows, S=20 states, A=11 actions, float32 rows in, float32
value + int32 argmax + float32 overall_value out (24 algorithmic bytes per scenario-step), plus the
summary-statistics reduction (and, for N > 1, its sum over ranks).  Steps are pipelined over four output sets: the
summary of step i runs on a side stream beside the stream kernel of step i + 1; the timed region ends when the last
summary is in.  N > 1 is weak scaling: every rank runs its own 4096 instances, no data-path collective.

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


def config_dict(n_inst, n_steps, world):
    """The `config` object both arms print (the reference arm times a bounded sample of this workload)."""
    return {"workload": WORKLOAD, "n_inst_per_gpu": n_inst, "n_steps": n_steps, "S": S, "A": A,
            "io": "f32 rows in, f32 value/overall + i32 argmax out; float64 accumulation and bounds",
            "cache": "inputs (983 MB) + outputs (983 MB) per step exceed the 126 MB L2",
            "parallelism": f"instances sharded, {world} rank(s), one stats all-reduce per step"}   # (peer-memory sum in the summary kernel when the ranks can map each other)


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


def pcie_probe(dev, nb_in, nb_out, ddist, reps=3):
    """Bare cudaMemcpyAsync of one step's bytes, both directions at once on two streams, no kernel:
    the floor of the end-to-end time on this host (max over ranks, all ranks copying at once)."""
    import torch
    hin = torch.empty(nb_in, dtype=torch.uint8).pin_memory()
    hout = torch.empty(nb_out, dtype=torch.uint8).pin_memory()
    din = torch.empty(nb_in, dtype=torch.uint8, device=dev)
    dout = torch.zeros(nb_out, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def once():
        with torch.cuda.stream(s1):
            din.copy_(hin, non_blocking=True)
        with torch.cuda.stream(s2):
            hout.copy_(dout, non_blocking=True)
        s1.synchronize(); s2.synchronize()

    once()
    ddist.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        once()
    ms = (time.perf_counter() - t0) * 1e3 / reps
    t = torch.tensor([ms], dtype=torch.float64, device=dev)
    ddist.allreduce_max(t)
    ms = float(t[0])
    return {"ms": ms, "h2d_GBs": nb_in / ms / 1e6, "d2h_GBs": nb_out / ms / 1e6,
            "note": "per rank: pinned-host copies of one step's bytes, both directions concurrently, no kernel"}


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
    # NSETS output sets: batch b's summary kernel (+ the sum over ranks) runs on a side stream while the stream kernels of
    # the next batches fill the other sets ("overlap batch b's all-reduce with batch b+1's kernel via two
    # streams").  The sum over ranks makes every rank wait for the slowest one of THAT step; with four sets a rank can be
    # three steps ahead before it has to, so the jitter of single steps is absorbed (measured at 8 ranks with two sets:
    # 12 to 69 us per step beyond the slowest rank's kernel, depending on the run).
    NSETS = 4
    plans = [dc.StreamPlan(n_inst, n_steps, S, A, r_dtype=torch.float32, out_dtype=torch.float32,
                           want_overall=True, want_stats=True) for _ in range(NSETS)]
    plan = plans[0]
    stream = torch.cuda.current_stream()
    side = torch.cuda.Stream(device=dev, priority=-1)
    # N > 1: the summary kernel ends with the sum over ranks through peer-mapped memory (one launch, no NCCL call
    # on the step's critical path); without peer mapping the vector goes through NCCL after the kernel
    import ctypes as C
    L = dc._abi.lib()
    peers = ddist.PeerGroup(ddist.STATS_HEADER + A, None, dev)
    fused_sum = world > 1 and peers.ctx is not None
    reds = [torch.zeros(ddist.STATS_HEADER + A, dtype=torch.float64, device=dev) for _ in range(NSETS)]

    def stats_step(pl, hist_i, red, st):
        """summary vector of the batch in `pl`'s outputs (+ its action histogram) -> red (summed over ranks), on stream st"""
        if fused_sum:
            dc._abi.check(L.dcarl_stream_stats_reduce(
                dc._dev.ptr(pl.value), 0, dc._dev.ptr(pl.overall), dc._dev.ptr(pl.activation_step),
                dc._dev.ptr(hist_i), n_inst, n_steps, S, A, dc._dev.ptr(red), ddist.STATS_HEADER + A, 0,
                C.byref(peers.ctx), dc._dev.stream_handle(st)))
        else:
            dc._abi.check(L.dcarl_stream_stats(
                dc._dev.ptr(pl.value), 0, dc._dev.ptr(pl.overall), dc._dev.ptr(pl.activation_step),
                dc._dev.ptr(hist_i), n_inst, n_steps, S, A, dc._dev.ptr(red), dc._dev.stream_handle(st)))
            if world > 1:
                with torch.cuda.stream(st):
                    ddist.allreduce_stats(red)

    def run_steps(k, k_ev=None):
        """k pipelined steps: stream kernel of step i on the main stream, its summary (+ sum over ranks) on the side
        stream; step i + NSETS reuses step i's output set and waits for its summary.  Every step has its own zeroed
        action histogram (one memset for all of them, outside the loop); bad_rows accumulates over the steps."""
        done_rows = [torch.cuda.Event() for _ in range(k)]
        done_stats = [torch.cuda.Event() for _ in range(k)]
        hist = torch.zeros((k, A), dtype=torch.int64, device=dev)
        for i in range(k):
            pl = plans[i % NSETS]
            if i >= NSETS:
                stream.wait_event(done_stats[i - NSETS])
            if k_ev is not None:
                k_ev[i][0].record(stream)     # events bracket the dominant kernel (k_stream_rows) on the launching stream
            dc._abi.check(L.dcarl_stream_bounds(
                dc._dev.ptr(idx), dc._dev.ptr(act), dc._dev.ptr(R), 0, n_inst, n_steps, S, A, pl.params,
                dc._dev.ptr(pl.value), 0, dc._dev.ptr(pl.act), dc._dev.ptr(pl.activation_step),
                dc._dev.ptr(pl.overall), dc._dev.ptr(hist[i]), dc._dev.ptr(pl.bad_rows),
                dc._dev.ptr(pl.ws), pl._ws_bytes, 0, dc._dev.stream_handle(stream)))
            if k_ev is not None:
                k_ev[i][1].record(stream)
            done_rows[i].record(stream)
            side.wait_event(done_rows[i])
            stats_step(pl, hist[i], reds[i % NSETS], side)
            done_stats[i].record(side)
        stream.wait_event(done_stats[-1])     # the timed region ends when the last summary is in (the side stream is in order)

    for pl in plans:
        pl.bad_rows.zero_()
    run_steps(max(args.warmup, 3))
    torch.cuda.synchronize()
    for pl in plans:
        pl.raise_on_bad_rows()
    sampler = ClockSampler(physical_gpu_index(local))
    ddist.barrier()
    torch.cuda.synchronize()
    sampler.start()
    k_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(stream)
    run_steps(args.steps, k_ev)
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
    stats = reds[(args.steps - 1) % NSETS].cpu().numpy()

    # ---- e2e: host buffers in, host buffers out, copies inside the timed region ---------------------
    # The public call is StreamPlan.run_pinned.  Headline: the compact wire format (one-byte keys in, one-byte
    # argmax out: 6 + 9 bytes per scenario-step on the link); the int32 format (12 + 12) is timed beside it.
    def e2e_leg(key_dtype):
        ep = dc.StreamPlan(n_inst, n_steps, S, A, r_dtype=torch.float32, out_dtype=torch.float32,
                           want_overall=True, want_stats=True, key_dtype=key_dtype)
        pin = ep.pinned()
        pin["idx"].copy_(idx.to(key_dtype)); pin["act"].copy_(act.to(key_dtype)); pin["R"].copy_(R)
        torch.cuda.synchronize()
        k = max(2, min(args.steps, 5))
        ep.run_pinned()                                   # warm-up (first touch of the pinned pages)
        ddist.barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(k):
            ep.run_pinned()                               # returns after the stream drained
            ddist.allreduce_stats(ep.stats)
        torch.cuda.synchronize()
        ms = (time.perf_counter() - t0) * 1e3 / k
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        ddist.allreduce_max(t)
        same = (bool(torch.equal(pin["value"], plan.value.cpu())) and bool(torch.equal(pin["overall"], plan.overall.cpu()))
                and bool(torch.equal(pin["oact"].to(torch.int32), plan.act.cpu())))
        res = {"ms_per_step": float(t[0]), "h2d_bytes_per_step": ep.h2d_bytes(), "d2h_bytes_per_step": ep.d2h_bytes(),
               "outputs_match_device_run": same}
        del ep, pin
        torch.cuda.empty_cache()
        return res

    e2e_c = e2e_leg(torch.uint8)
    e2e_w = e2e_leg(torch.int32)
    probe = pcie_probe(dev, e2e_c["h2d_bytes_per_step"], e2e_c["d2h_bytes_per_step"], ddist)

    out = None
    if rank == 0:
        peak, peak_src = measured_peak()
        achieved = n_inst * n_steps * BYTES_PER_STEP / (k_ms_max * 1e-3) / 1e9
        out = {
            "metric": METRIC, "value": value, "unit": "scenario-steps/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms_per_step,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64", "io_dtype": "f32 rows in, f32 values out (the arithmetic type is float64: moments and bounds)",
        "data": "synthetic",
            "config": config_dict(n_inst, n_steps, world),
            "clocks": dict(sampler.summary(), sampled_over=clocks_from),
            "e2e": {"value": total_steps / (e2e_c["ms_per_step"] * 1e-3), "unit": "scenario-steps/s",
                    "h2d_bytes_per_step": e2e_c["h2d_bytes_per_step"], "d2h_bytes_per_step": e2e_c["d2h_bytes_per_step"],
                    "ms_per_step": e2e_c["ms_per_step"], "outputs_match_device_run": e2e_c["outputs_match_device_run"],
                    "pcie_GBs": {"h2d": e2e_c["h2d_bytes_per_step"] / e2e_c["ms_per_step"] / 1e6,
                                 "d2h": e2e_c["d2h_bytes_per_step"] / e2e_c["ms_per_step"] / 1e6,
                                 "note": "per rank, bytes of one step / its end-to-end time; the two directions overlap"},
                    "wire": "compact: uint8 idx/act + f32 R in, f32 value/overall + uint8 argmax out (6 + 9 B per scenario-step)",
                    "api": "dcarl_b200.StreamPlan(key_dtype=uint8).run_pinned (pinned host in/out, copies in the timed region)",
                    "int32_wire": {"value": total_steps / (e2e_w["ms_per_step"] * 1e-3), "ms_per_step": e2e_w["ms_per_step"],
                                   "h2d_bytes_per_step": e2e_w["h2d_bytes_per_step"],
                                   "d2h_bytes_per_step": e2e_w["d2h_bytes_per_step"],
                                   "outputs_match_device_run": e2e_w["outputs_match_device_run"]},
                    "bare_copy_probe": probe},
            "gpu_launches": 2 * args.steps,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": ncu_traffic(), "kernel": "k_stream_rows",
                         "kernel_ms": k_ms_max, "algorithmic_bytes_per_step": BYTES_PER_STEP, "peak_source": peak_src},
            "stats_check": {"total_steps": float(stats[4]), "expected": float(total_steps),
                            "activated_pairs": float(stats[2]),
                            "collective": ("sum over ranks through peer memory at the end of the summary kernel" if fused_sum
                                           else "NCCL all-reduce after the summary kernel" if world > 1 else "none (one rank)"),
                            "pipeline": "summary (+ sum over ranks) of step i on a side stream beside the stream kernels of "
                                        "the next steps (four output sets)"},
        }
    if rank == 0 and world == 1 and not args.no_cpu:
        out.update(cpu_legs(plan, idx, act, R))
    # BASELINE configs 4 and 5 at this world size (every rank takes part; rank 0 reports)
    del idx, act, R
    plan = plans = None
    torch.cuda.empty_cache()
    if not args.no_configs:
        peak = measured_peak()[0]
        for name, leg in (("config4", config4_leg), ("config5", config5_leg)):
            try:
                res = leg(dev, rank, world, peak, args)
            except Exception as e:   # noqa: BLE001  (a failing leg must not cost the headline line; every rank fails alike)
                res = {"error": repr(e)}
            if rank == 0:
                out[name] = res
    if rank == 0 and world == 1 and not args.no_secondary:
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
    """Roofline of the other kernels on the path (BASELINE config 3, planner row f3): CUDA-event time of
    back-to-back launches on inputs larger than L2.  Configs 4 and 5 have their own legs (every world size)."""
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

    g = torch.Generator(device=dev).manual_seed(3000)
    # config 3 mode A: 65 536 episodes x 200 draws, deviates supplied (16 B/row) and on-device Philox (12 B/row)
    n = 65536 * 200
    qs = torch.rand((20, 11), generator=g, device=dev) * 150 - 50
    zs, zr = torch.randn(n, generator=g, device=dev), torch.randn(n, generator=g, device=dev)
    ac = torch.randint(0, 11, (n,), generator=g, device=dev, dtype=torch.int32)
    o_idx, o_R = torch.empty(n, dtype=torch.int32, device=dev), torch.empty(n, dtype=torch.float32, device=dev)
    ms = timed(lambda: dc.sample_returns(qs, zs, ac, zr, out=(o_idx, o_R)), 50)
    out["sample_returns"] = {"rows": n, "ms": ms, "rows_per_s": n / (ms * 1e-3), "achieved_GBs": n * 20 / (ms * 1e-3) / 1e9,
                             "frac": n * 20 / (ms * 1e-3) / 1e9 / peak, "algorithmic_bytes_per_row": 20,
                             "note": "12 B read (z_state, z_ret, act) + 8 B written (idx, R); caller-held outputs"}
    o_act = torch.empty(n, dtype=torch.int32, device=dev)
    ms = timed(lambda: dc.sample_returns_philox(qs, n, 7, out=(o_idx, o_act, o_R)), 50)
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
    out["closed_loop"] = closed_loop_leg(dev)
    return out


# ---- BASELINE config 4: confidence over a 1e7-row {s,a,R} buffer, 256 candidate actions, row-sharded -------
C4_ROWS, C4_ACTIONS, C4_CHUNKS = 10_000_000, 256, 8


def c4_chunk(c, dev):
    """Rows [c, c+1) * C4_ROWS / 8 of the global buffer: the same bytes whatever the world size."""
    import torch
    import dcarl_b200 as dc
    g = torch.Generator(device=dev).manual_seed(3000)
    dist = torch.tensor(dc.VISITED_STATE_DIST[0], dtype=torch.float64, device=dev)
    centers = torch.randn((4096, 20), generator=g, device=dev) * dist[:20].float()
    g = torch.Generator(device=dev).manual_seed(3100 + c)
    m = C4_ROWS // C4_CHUNKS
    st = centers[torch.randint(0, 4096, (m,), generator=g, device=dev)]
    st += torch.randn((m, 20), generator=g, device=dev) * (0.25 * dist[:20].float())
    a = torch.randint(0, C4_ACTIONS, (m, 1), generator=g, device=dev).float()
    ret = -(0.95 ** torch.randint(0, 10, (m,), generator=g, device=dev).float()) * (torch.rand(m, generator=g, device=dev) < 0.5)
    return centers, torch.cat([st, a], 1), ret


def config4_leg(dev, rank, world, peak, args):
    """Strong scaling: the buffer is row-sharded M/world per rank, the queries are replicated; one query
    pass = scan of the shard + sum of the raw moments over ranks + act_test pick on every rank.  Checked in
    the run against the oracle's single-pass restatement on the same rows (host cores)."""
    import torch
    import dcarl_b200 as dc
    from dcarl_b200 import dist as ddist
    from oracle import cport
    if C4_CHUNKS % world:
        return {"skipped": f"world size {world} does not divide the {C4_CHUNKS} buffer chunks"}
    mine = range(rank * C4_CHUNKS // world, (rank + 1) * C4_CHUNKS // world)
    buf = dc.BoxBuffer(20, capacity=len(mine) * (C4_ROWS // C4_CHUNKS))
    for c in mine:
        centers, st, ret = c4_chunk(c, dev)
        buf.append(st, ret)
        del st, ret
    out = {"rows": C4_ROWS, "rows_per_gpu": len(buf), "n_actions": C4_ACTIONS, "n_gpus": world, "scaling": "strong",
           "algorithmic_bytes_per_row": 88}
    q64 = centers[:64].double().contiguous()
    results = {}
    for Q in (1, 64):
        q = q64[:Q].contiguous()
        run = ddist.ShardedBoxQuery(buf, Q, C4_ACTIONS, graph=True)
        reps = 50 if Q == 1 else 10
        run(q); run(q)
        ddist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(reps):
            mom, mv, pick = run(q)
        e1.record()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
        ddist.allreduce_max(t)
        ms = float(t[0])
        gbs = C4_ROWS * 88 / (ms * 1e-3) / 1e9
        results[Q] = (mom.clone(), pick.clone())
        out[f"Q{Q}"] = {"ms_per_query_pass": ms, "queries_per_s": Q / (ms * 1e-3), "rows_per_s": C4_ROWS / (ms * 1e-3),
                        "achieved_GBs": gbs, "frac_of_aggregate_peak": gbs / (peak * world),
                        "matched_rows": float(mom[:, :, 0].sum()), "path": run.path + " (CUDA-graph replay)"}
    # parity: the oracle on this rank's rows, partial moments summed over ranks like the device's
    cores = max(1, _cores() // world)
    t0 = time.perf_counter()
    ref = cport.box_raw(buf.state.cpu().numpy(), buf.ret.cpu().numpy(), dc.VISITED_STATE_DIST[0], q64.cpu().numpy(),
                        C4_ACTIONS, threads=cores)
    dt = time.perf_counter() - t0
    ref_t = torch.from_numpy(ref).to(dev)
    ddist.allreduce_stats(ref_t)
    ref = ref_t.cpu().numpy()
    n = np.maximum(ref[:, :, 0], 1.0)
    nmv = np.stack([ref[:, :, 0], np.where(ref[:, :, 0] > 0, ref[:, :, 1] / n, -1.0),
                    np.where(ref[:, :, 0] > 0, np.maximum(ref[:, :, 2] / n - (ref[:, :, 1] / n) ** 2, 0.0), -1.0)], 2)
    ref_pick = cport.box_select(nmv)
    mism, sum_err, pick_mism = 0, 0.0, 0
    for Q in (1, 64):
        mom, pick = results[Q][0].cpu().numpy(), results[Q][1].cpu().numpy()
        mism += int((mom[:, :, 0] != ref[:Q, :, 0]).sum())
        sum_err = max(sum_err, float(np.abs(mom[:, :, 1:] - ref[:Q, :, 1:]).max()))
        pick_mism += int((pick != ref_pick[:Q]).sum())
    out["parity"] = {"count_mismatches": mism, "pick_mismatches": pick_mism, "max_abs_sum_error": sum_err,
                     "bins_checked": 65 * C4_ACTIONS, "rows_checked_per_rank": len(buf),
                     "oracle": f"oracle/dcarl_oracle.c orc_box_raw on {cores} host threads per rank, {dt:.1f} s"}
    out["cpu_baseline"] = {"value": len(buf) * 64 / dt, "unit": "row-queries/s", "cores": cores, "kind": "port",
                           "sample": f"{len(buf)} rows x 64 queries per rank"}
    return out


# ---- BASELINE config 5: mixed Simulation_1 + Simulation_2 batch, 262 144 scenarios x 500 steps, sharded -------
C5_INST, C5_STEPS, C5_CHUNK = 262144, 500, 4096


def config5_leg(dev, rank, world, peak, args):
    """Strong scaling: even instance ids are Simulation_1-shaped (S=1, A=30, 11 live actions), odd ids
    Simulation_2-shaped (S=20, A=11); each rank owns a contiguous block of 4096-instance chunks (the same
    rows whatever the world size), runs its two launches and ONE all-reduce of the merged stats vector."""
    import torch
    import dcarl_b200 as dc
    from dcarl_b200 import dist as ddist
    from oracle import cport
    n_chunks = C5_INST // 2 // C5_CHUNK                       # per shape
    if n_chunks % world:
        return {"skipped": f"world size {world} does not divide the {n_chunks} chunks per shape"}
    mine = range(rank * n_chunks // world, (rank + 1) * n_chunks // world)
    shapes = ((1, 30, 11), (20, 11, None))
    data, plans = [], []
    for si, (s_, a_, live) in enumerate(shapes):
        parts = [synth_device(C5_CHUNK, C5_STEPS, 4000 + 2 * c + si, dev, S=s_, A=a_, live=live) for c in mine]
        data.append(tuple(torch.cat([p[j] for p in parts]) for j in range(3)))
        plans.append(dc.StreamPlan(len(mine) * C5_CHUNK, C5_STEPS, s_, a_, want_overall=True, want_stats=True))
    del parts
    batch = ddist.MixedStreamBatch(plans, keep_local=False)   # the last summary kernel ends with the sum over ranks

    def step():
        return batch.run(data)

    for _ in range(3):
        stats = step()
    ddist.barrier()
    torch.cuda.synchronize()
    reps = 50
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        stats = step()
    e1.record()
    torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], dtype=torch.float64, device=dev)
    ddist.allreduce_max(t)
    ms = float(t[0])
    for pl in plans:
        pl.raise_on_bad_rows()
    total = C5_INST * C5_STEPS
    gbs = total * BYTES_PER_STEP / (ms * 1e-3) / 1e9
    # parity 1: sampled instances of this rank against the oracle (argmax and activation step bit-exact)
    rng = np.random.default_rng(77 + rank)
    mism, rows_checked, act_mism = 0, 0, 0
    for (s_, a_, _), pl, (i_, a__, r_) in zip(shapes, plans, data):
        pick = torch.from_numpy(rng.choice(pl.n_inst, 16, replace=False)).to(dev)
        ref = cport.stream(i_[pick].cpu().numpy(), a__[pick].cpu().numpy(), r_[pick].cpu().numpy().astype(np.float64), s_, a_)
        mism += int((ref["act"] != pl.act[pick].cpu().numpy()).sum())
        act_mism += int((ref["activation_step"] != pl.activation_step[pick].cpu().numpy()).sum())
        rows_checked += 16 * C5_STEPS
    pm = torch.tensor([mism, act_mism, rows_checked], dtype=torch.float64, device=dev)
    ddist.allreduce_stats(pm)
    # parity 2: the all-reduced vector against the rank-ordered sum of the local vectors
    stats = stats.clone()
    batch.keep_local = True                                   # one more batch that keeps this rank's own vector
    stats_unfused = batch.run(data)
    local = batch.local_stats.clone()
    gathered = ddist.allgather(local)
    fixed = torch.stack(gathered).sum(0)
    st, fx = stats.cpu().numpy(), fixed.cpu().numpy()
    return {"scenarios": C5_INST, "steps": C5_STEPS, "scenarios_per_gpu": 2 * len(mine) * C5_CHUNK, "n_gpus": world,
            "scaling": "strong", "ms_per_batch": ms, "value": total / (ms * 1e-3), "unit": "scenario-steps/s",
            "achieved_GBs": gbs, "frac_of_aggregate_peak": gbs / (peak * world), "algorithmic_bytes_per_step": BYTES_PER_STEP,
            "timed": "two stream launches + two chained summary kernels, the second ending with the sum over ranks "
                     "(peer memory), 50 batches back to back",
            "collective": batch.peers.backend,
            "parity": {"argmax_mismatches": int(pm[0].item()), "activation_step_mismatches": int(pm[1].item()),
                       "rows_checked": int(pm[2].item()),
                       "fused_equals_unfused_collective": bool(torch.equal(stats, stats_unfused)),
                       "allreduced_vs_rank_ordered_sum": {"integer_fields_equal": bool(np.array_equal(st[2:], fx[2:])),
                                                          "max_rel_err_float_fields": float(np.abs(st[:2] - fx[:2]).max() / max(1.0, np.abs(fx[:2]).max()))}},
            "stats": {"sum_final_value": float(st[0]), "sum_final_overall": float(st[1]), "activated_pairs": int(st[2]),
                      "sum_activation_step": int(st[3]), "total_steps": int(st[4]), "action_histogram": [int(x) for x in st[5:]],
                      "note": "the integer fields are identical for every world size (same rows, exact sums)"}}


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


def closed_loop_leg(dev):
    """The collector's loop the metric is named after (dqn_value_collect.py:86-145 under the reward / termination
    rules of TestScenario_Town03.py:403-422), on the device from end to end for a batch of scenarios: planner tick
    (10 candidates) -> episode return of every candidate against constant-velocity obstacles -> (state, action,
    return) rows into the box buffer -> box-kernel confidence of every scenario's state -> act_test pick.  Stage
    times from CUDA events; the returns of a sample are checked against the oracle's restatement of those lines."""
    import torch
    import dcarl_b200 as dc
    from oracle import pyref
    n, n_paths = 32768, 64
    knots, off, pid, start, ego_xy, obs, n_obs = synth_frenet(n, n_paths, 6000, max_obs=3)
    rng = np.random.default_rng(6001)
    ang, v = rng.uniform(-np.pi, np.pi, n), rng.uniform(3, 9, n)
    ego = np.c_[ego_xy, v * np.cos(ang), v * np.sin(ang), v]
    routes = dc.Routes(knots)
    t = lambda a, dt: torch.from_numpy(np.ascontiguousarray(a)).to(dev, dt)
    d_pid, d_ego, d_obs, d_nobs = t(pid, torch.int32), t(ego, torch.float64), t(obs, torch.float64), t(n_obs, torch.int32)
    rp = dc.rollout_params(pass_y=float(np.median(ego[:, 1])))
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    reps, stage = 3, np.zeros(4)
    for rep in range(reps + 1):                        # the first pass is the warm-up
        buf = dc.BoxBuffer(20, capacity=n * 11)
        torch.cuda.synchronize()
        ev[0].record()
        st0 = dc.frenet_start(routes, d_pid, d_ego)
        exy = d_ego[:, :2].contiguous()
        plan = dc.frenet_plan(routes, d_pid, st0, exy, d_obs, d_nobs, want_paths=True)
        ev[1].record()
        ret, steps, term = dc.rollout_candidates(plan, exy, d_obs, d_nobs, rp)
        ev[2].record()
        A = ret.shape[1]
        st = torch.zeros((n, 20), dtype=torch.float64, device=dev)
        st[:, 0:4] = d_ego[:, 0:4]
        st[:, 4] = torch.atan2(d_ego[:, 3], d_ego[:, 2])
        present = (torch.arange(3, device=dev)[None, :] < d_nobs[:, None]).to(torch.float64)
        st[:, 5:20] = (d_obs[:, :3, :] * present[:, :, None]).reshape(n, 15)
        rows = torch.cat([st[:, None, :].expand(n, A, 20),
                          torch.arange(A, dtype=torch.float64, device=dev)[None, :, None].expand(n, A, 1)], 2)
        buf.append(rows.reshape(n * A, 21), ret.reshape(n * A))
        ev[3].record()
        mom = buf.moments(st, A)                       # n queries against the n x A rows just recorded: ONE launch,
        mv, pick = buf.select(mom)                     # a CTA per chunk of 64 queries scans the (L2-resident) buffer
        ev[4].record()
        torch.cuda.synchronize()
        if rep:
            stage += [ev[i].elapsed_time(ev[i + 1]) for i in range(4)]
        del buf
    stage /= reps
    # parity of the rollout on a sample: returns / steps / termination codes of every action
    paths, nv, cost = plan.paths.cpu().numpy(), plan.n_valid.cpu().numpy(), plan.cost.cpu().numpy()
    g_ret, g_steps, g_term = ret.cpu().numpy(), steps.cpu().numpy(), term.cpu().numpy()
    bad, max_err = 0, 0.0
    sample = rng.choice(n, 64, replace=False)
    for i in sample:
        want = pyref.rollout_candidates(paths[i], nv[i], cost[i], ego_xy[i], [tuple(x) for x in obs[i, :n_obs[i]]], considered=4,
                                        move_gap=1.0, n_pred=15, dt=0.3, radius=1.0, pass_y=rp.pass_y, stay_thres=rp.stay_thres)
        bad += int([w[1] for w in want] != g_steps[i].tolist()) + int([w[2] for w in want] != g_term[i].tolist())
        max_err = max(max_err, float(np.abs(np.array([w[0] for w in want]) - g_ret[i]).max()))
    total = float(stage.sum())
    return {"scenarios": n, "actions_per_scenario": int(A), "ms": {"planner_tick": stage[0], "candidate_rollout": stage[1],
                                                                  "record_rows": stage[2], "confidence_and_pick": stage[3], "total": total},
            "scenarios_per_s": n / (total * 1e-3), "candidate_episodes_per_s": n * int(A) / (stage[1] * 1e-3),
            "kinematic_steps_per_s": float(g_steps.sum()) / (stage[1] * 1e-3), "rows_recorded": n * int(A),
            "picks_other_than_rule": int((pick != 0).sum()),
            "parity": {"sampled_scenarios": len(sample), "step_or_termination_mismatches": bad, "max_abs_return_error": max_err,
                       "oracle": "oracle/pyref.py rollout_candidates (restates dqn_value_collect.py:86-145, TestScenario_Town03.py:403-422, predict.py:52-59)"},
            "note": "confidence_and_pick answers the n queries against the n x A rows just recorded in one filtered launch "
                    "(n / 64 query chunks, a CTA per chunk over the L2-resident buffer) + the select"}


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
    per_core = int(min(4, max(1, 150.0 / (args.steps + max(args.warmup, 3)) / 3.5)))
    n_lit = per_core * procs
    idx, act, R = synth_host(n_lit, args.n_steps, 1000)
    warm = max(args.warmup, 3)                   # same rule as our arm
    for _ in range(warm):
        cpu_literal(idx, act, R, procs)
    times = []
    for _ in range(args.steps):
        v, _, dt = cpu_literal(idx, act, R, procs)
        times.append(dt)
    ms = float(np.mean(times)) * 1e3
    value = idx.size / (ms * 1e-3)
    out = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "scenario-steps/s", "n_gpus": world,
        "steps": args.steps, "warmup": warm, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "io_dtype": "f32 rows in, f32 values out (the arithmetic type is float64: moments and bounds)",
        "data": "synthetic",
        "config": config_dict(args.n_inst, args.n_steps, world),
        "sampled_instances": n_lit,
        "note": "the reference is single-process Python/NumPy and cannot travel to the GPU box; this arm times its "
                "literal restatement (oracle/pyref.py, pinned bit-exact against the reference's outputs) on all host "
                "cores, one instance per task, on a bounded sample of the config's instances per step (rates are per "
                "scenario-step, so the sample size does not enter the ratio)",
        "cpu_baseline": {"value": value, "unit": "scenario-steps/s", "cores": procs, "kind": "port",
                         "sample": f"{n_lit} of the {args.n_inst} instances x {args.n_steps} steps per step"},
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
    ap.add_argument("--no-secondary", action="store_true", help="skip the sampler / rollout / planner legs")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE config-4 and config-5 legs")
    args = ap.parse_args()
    if args.steps is None:
        args.steps = 3 if args.impl == "reference" else 50
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
