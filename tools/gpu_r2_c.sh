#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_box_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout 200 > gpurun_out/pytest_box_r2c.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_box_r2c.log
timeout 200 python tools/box_q64_probe.py 2>&1 | tail -6
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_box_moments -c 12 --csv --log-file gpurun_out/box_ncu_r2c.csv python tools/box_q64_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/box_ncu_r2c.csv")) if len(r) > 10]
h = rows[0]; ik, im, iv = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
iid = h.index("ID")
last = {}
for r in rows[1:]:
    print(r[iid], r[ik][:60], r[im], r[iv]) if r[iid] in ("3", "11") else None
PY
