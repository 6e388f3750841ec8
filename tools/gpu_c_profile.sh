#!/bin/bash
# full ncu capture of the stream kernel + a short bench run (new legs)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_rows -c 1 -f -o gpurun_out/rows_${TAG:-r2c} \
    python tools/stream_sweep.py --once > gpurun_out/ncu_rows_full.log 2>&1
echo "full capture exit $?"
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_quick.json 2> gpurun_out/bench_quick.err
echo "bench exit $?"
tail -c 6000 gpurun_out/bench_quick.json; tail -20 gpurun_out/bench_quick.err
