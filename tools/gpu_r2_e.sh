#!/bin/bash
mkdir -p gpurun_out
for rep in 1 2; do timeout 120 python tools/stream_sweep.py --steps 30 2>&1 | grep -o '"ms": [0-9.]*' | sed 's/^/current      /'; done
timeout 300 python -m pytest tests/test_stream_gpu.py -m gpu -q -p no:cacheprovider --timeout 200 2>&1 | tail -2
timeout 400 ncu --set full --import-source on --clock-control none -k regex:k_stream_rows -s 2 -c 1 -o gpurun_out/rows_r2e -f python tools/stream_sweep.py --steps 2 --warmup 1 > gpurun_out/ncu_rows_r2e.log 2>&1; tail -2 gpurun_out/ncu_rows_r2e.log
