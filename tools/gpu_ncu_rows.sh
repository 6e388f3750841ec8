#!/bin/bash
# Full ncu capture (source-level) of one k_stream_rows launch on config 2.
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_rows -s 3 -c 1 \
    -f -o gpurun_out/stream_rows python tools/stream_sweep.py --steps 1 --warmup 3 > gpurun_out/ncu_rows.log 2>&1
echo "capture exit $?"; ls -la gpurun_out/stream_rows.ncu-rep
