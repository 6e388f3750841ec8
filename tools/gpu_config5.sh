#!/bin/bash
# config-5 shapes (262 144 x 500, half Simulation_1-shaped, half Simulation_2-shaped): kernel time and light ncu metrics
mkdir -p gpurun_out
timeout 200 python tools/stream_sweep.py --n-inst 131072 --n-steps 500 --steps 20 2>&1 | grep -o '"ms": [0-9.]*' | sed 's/^/sim2-shaped 131072x500 /'
timeout 200 python tools/stream_sweep.py --n-inst 131072 --n-steps 500 --S 1 --A 30 --live 11 --steps 20 2>&1 | grep -o '"ms": [0-9.]*' | sed 's/^/sim1-shaped 131072x500 /'
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__block_size,launch__registers_per_thread,launch__occupancy_limit_shared_mem,launch__occupancy_limit_registers,launch__occupancy_limit_warps,launch__occupancy_limit_blocks
timeout 300 ncu --metrics $M --clock-control none -k regex:k_stream_rows -c 2 --csv --log-file gpurun_out/c5_ncu_a.csv python tools/stream_sweep.py --n-inst 131072 --n-steps 500 --once > /dev/null 2>&1
timeout 300 ncu --metrics $M --clock-control none -k regex:k_stream_rows -c 2 --csv --log-file gpurun_out/c5_ncu_b.csv python tools/stream_sweep.py --n-inst 131072 --n-steps 500 --S 1 --A 30 --live 11 --once > /dev/null 2>&1
python - <<'PY'
import csv
for f in ("gpurun_out/c5_ncu_a.csv", "gpurun_out/c5_ncu_b.csv"):
    rows = [r for r in csv.reader(open(f)) if len(r) > 10]
    h = rows[0]; im, iv, iid = h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
    print(f, {r[im].split("__")[-1][:40]: r[iv] for r in rows[1:] if r[iid] == "0"})
PY
