#!/usr/bin/env python
"""Time the stream-kernel variants on the BASELINE config-2 workload and cross-check their outputs.

    python tools/stream_sweep.py [--variants 0,1] [--steps 10] [--once] [--n-inst 4096] [--n-steps 20000]

One JSON line per variant (also appended to gpurun_out/stream_sweep.jsonl): kernel ms (CUDA events
around `steps` back-to-back launches on the launching stream), scenario-steps/s, algorithmic GB/s
at 24 B per scenario-step, and the comparison with the first variant listed (argmax / activation
step bit-exact, value and overall_value by max abs difference).  `--once` launches every variant
exactly once (for a pass under ncu).
"""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--variants", default="0")
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--n-inst", type=int, default=4096)
    ap.add_argument("--n-steps", type=int, default=20000)
    ap.add_argument("--S", type=int, default=None)
    ap.add_argument("--A", type=int, default=None)
    ap.add_argument("--live", type=int, default=None, help="actions that occur in the rows (Simulation_1: 11 of 30)")
    ap.add_argument("--once", action="store_true")
    ap.add_argument("--no-overall", action="store_true")
    args = ap.parse_args()

    import torch
    import bench
    import dcarl_b200 as dc

    dev = torch.device("cuda:0")
    torch.cuda.set_device(dev)
    S, A = args.S or bench.S, args.A or bench.A
    idx, act, R = bench.synth_device(args.n_inst, args.n_steps, 1234, dev, S=S, A=A, live=args.live)
    rows = args.n_inst * args.n_steps
    bps = 20 if args.no_overall else 24
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    ref = None
    for v in [int(x) for x in args.variants.split(",")]:
        plan = dc.StreamPlan(args.n_inst, args.n_steps, S, A, want_overall=not args.no_overall,
                             want_stats=False, variant=v, device=dev)
        out = {"variant": v, "n_inst": args.n_inst, "n_steps": args.n_steps, "S": S, "A": A}
        try:
            if args.once:
                plan.run_device(idx, act, R)
                torch.cuda.synchronize()
            else:
                for _ in range(args.warmup):
                    plan.run_device(idx, act, R)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(args.steps):
                    plan.run_device(idx, act, R)
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / args.steps
                out.update(ms=ms, steps_per_s=rows / ms * 1e3, algorithmic_GBs=rows * bps / ms / 1e6)
            plan.raise_on_bad_rows()
            cur = dict(value=plan.value.clone(), act=plan.act.clone(), astep=plan.activation_step.clone(),
                       overall=None if plan.overall is None else plan.overall.clone())
            if ref is None:
                ref = cur
            else:
                out["act_mismatches"] = int((cur["act"] != ref["act"]).sum().item())
                out["activation_mismatches"] = int((cur["astep"] != ref["astep"]).sum().item())
                out["value_max_abs_diff"] = float((cur["value"].double() - ref["value"].double()).abs().max().item())
                if cur["overall"] is not None:
                    d = (cur["overall"].double() - ref["overall"].double()).abs()
                    out["overall_max_abs_diff"] = float(d.max().item())
                    out["overall_max_rel_diff"] = float((d / ref["overall"].double().abs().clamp_min(1.0)).max().item())
        except Exception as ex:   # a variant that fails must not hide the others
            out["error"] = f"{type(ex).__name__}: {ex}"
        line = json.dumps(out)
        print(line, flush=True)
        with open(os.path.join(ROOT, "gpurun_out", "stream_sweep.jsonl"), "a") as f:
            f.write(line + "\n")
        del plan


if __name__ == "__main__":
    main()
