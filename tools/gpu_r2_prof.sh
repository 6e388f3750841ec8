#!/bin/bash
# round-2 ncu passes (1 GPU): launch list of a short bench run; full capture of the filtered box scan (Q = 64)
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 400 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box_filter -s 3 -c 1 \
    -f -o gpurun_out/box_filter_r2 python tools/box_q64_probe.py > gpurun_out/ncu_box_filter.log 2>&1
echo "box filter capture exit $?"; tail -2 gpurun_out/ncu_box_filter.log
