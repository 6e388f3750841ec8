#!/bin/bash
# ncu passes (1 GPU): launch list of a short bench run, then one full capture of the dominant kernel.
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
    --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-secondary > gpurun_out/launches_bench.log 2>&1
echo "launch list exit $?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:${KERNEL:-k_stream_rows} -s 2 -c 1 \
    -f -o gpurun_out/${OUT:-stream_rows} python bench.py --steps 1 --warmup 3 --no-cpu --no-secondary > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
tail -3 gpurun_out/ncu_full.log
ls -la gpurun_out
