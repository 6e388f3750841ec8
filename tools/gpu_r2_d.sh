#!/bin/bash
# filter-kernel variants (warps per CTA x ring stages) on the config-4 buffer, then ncu of the default build
mkdir -p gpurun_out
echo "default (8 warps x 3 stages)"; timeout 200 python tools/box_q64_probe.py 2>&1 | grep "^64"
for v in ab_variants/f*/; do
  echo $v; DCARL_B200_LIB=$PWD/$v/libdcarl_b200.so timeout 200 python tools/box_q64_probe.py 2>&1 | grep "^64"
done
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum
for s in barrier short_scoreboard wait long_scoreboard math_pipe_throttle mio_throttle not_selected lg_throttle membar; do M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio; done
timeout 300 ncu --metrics $M --clock-control none -k regex:k_box_filter -c 4 --csv --log-file gpurun_out/box_ncu_r2d.csv python tools/box_q64_probe.py > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(open("gpurun_out/box_ncu_r2d.csv")) if len(r) > 10]
h = rows[0]; ik, im, iv, iid = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
for r in rows[1:]:
    if r[iid] in ("3",): print(r[ik][:40], r[im], r[iv])
PY
