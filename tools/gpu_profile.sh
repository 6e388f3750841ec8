#!/bin/bash
# ncu passes (1 GPU): launch list of a short bench run, then full captures of the dominant kernels.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 80 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_stream_rows -s 3 -c 1 \
    -f -o gpurun_out/stream_rows python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_full.log 2>&1
echo "stream capture exit $?"
if [ -n "$BOX" ]; then
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_box_moments -s 2 -c 1 \
    -f -o gpurun_out/box_moments python bench.py --steps 1 --warmup 3 --no-cpu > gpurun_out/ncu_box.log 2>&1
echo "box capture exit $?"
fi
grep -o '"kernel_ms": [0-9.]*' gpurun_out/ncu_full.log
ls -la gpurun_out | head -30
