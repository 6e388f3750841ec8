#!/usr/bin/env python
"""Summarise an .ncu-rep (read on the CPU box): headline metrics + instructions per CUDA source line.
usage: python tools/ncu_summary.py gpurun_out/stream_rows.ncu-rep [units_per_launch] [min_instr_per_unit] [summary.json]
With a fourth argument the DRAM traffic of the launch is also written as JSON (bench.py reads
profiles/stream_rows_ncu_summary.json for roofline.traffic)."""
import csv
import io
import subprocess
import sys

rep = sys.argv[1]
units = float(sys.argv[2]) if len(sys.argv) > 2 else 2.56e6     # 32-row batches of config 2
thr = float(sys.argv[3]) if len(sys.argv) > 3 else 1.5
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, unit, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "dram__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread",
        "launch__occupancy_limit_shared_mem", "launch__occupancy_limit_registers", "launch__waves_per_multiprocessor",
        "sm__cycles_elapsed.avg", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "smsp__thread_inst_executed_per_inst_executed.ratio", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_lsu.sum", "sm__inst_executed_pipe_alu.sum", "sm__inst_executed_pipe_fma.sum",
        "sm__inst_executed_pipe_fp64.sum", "sm__inst_executed_pipe_xu.sum", "sm__inst_executed_pipe_uniform.sum",
        "sm__inst_executed_pipe_cbu.sum", "sm__inst_executed_pipe_adu.sum",
        "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
print("kernel:", vals[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "?")
if len(sys.argv) > 4:
    import json
    g = lambda name: float(vals[hdr.index(name)].replace(",", ""))
    scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
    rd = g("dram__bytes_read.sum") * scale[unit[hdr.index("dram__bytes_read.sum")]]
    wr = g("dram__bytes_write.sum") * scale[unit[hdr.index("dram__bytes_write.sum")]]
    tu = {"ns": 1e-3, "us": 1.0, "ms": 1e3}[unit[hdr.index("gpu__time_duration.sum")]]
    json.dump({"kernel": vals[hdr.index("Kernel Name")], "source": f"ncu --set full, one launch ({rep})",
               "dram_bytes_read": rd, "dram_bytes_write": wr, "dram_bytes_per_launch": rd + wr,
               "inst_executed": g("smsp__inst_executed.sum"),
               "issue_active_pct": g("smsp__issue_active.avg.pct_of_peak_sustained_active"),
               "gpu_time_us_under_ncu": g("gpu__time_duration.sum") * tu}, open(sys.argv[4], "w"), indent=1)
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"  {w:70s} {vals[i]:>18s} {unit[i]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        try:
            if float(vals[i]) > 0.15:
                print(f"  stall {h[len('smsp__average_warps_issue_stalled_'):-len('_per_issue_active.ratio')]:40s} {vals[i]}")
        except ValueError:
            pass
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv"],
                     capture_output=True, text=True).stdout
cur, h, agg = None, None, []
for r in csv.reader(io.StringIO(src)):
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]
    elif len(r) > 2 and r[0] == "Line No":
        h = r
        ie, ism = h.index("Instructions Executed"), h.index("# Samples")
    elif h and len(r) > ie and r[0].strip().isdigit():
        try:
            agg.append((cur, int(r[0]), int(r[ie]), int(r[ism]) if r[ism].isdigit() else 0, r[1].strip()[:105]))
        except ValueError:
            pass
tot, ts = sum(a[2] for a in agg), max(1, sum(a[3] for a in agg))
print(f"instructions per unit: {tot / units:.1f}   (units = {units:g})")
for f, l, n, s, text in agg:
    if n / units >= thr:
        print(f"{f:20s}{l:5d} {n / units:7.1f} {100 * s / ts:5.1f}%  {text}")
