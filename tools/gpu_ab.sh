#!/bin/bash
# A/B on one box: the current build and every ab_variants/<name>/libdcarl_b200.so (builds of the same library with
# other -D switches), interleaved on the config-2 stream workload.
mkdir -p gpurun_out
for rep in 1 2; do
  timeout 120 python tools/stream_sweep.py --steps 30 2>&1 | grep -o '"ms": [0-9.]*' | sed 's/^/current      /'
  for v in ab_variants/*/; do
    [ -f $v/libdcarl_b200.so ] || continue
    n=$(basename $v)
    DCARL_B200_LIB=$PWD/$v/libdcarl_b200.so timeout 120 python tools/stream_sweep.py --steps 30 2>&1 | grep -o '"ms": [0-9.]*\|error.*' | sed "s/^/$n  /"
  done
done
timeout 120 python tools/stream_sweep.py --steps 30 --no-overall 2>&1 | grep -o '"ms": [0-9.]*' | sed 's/^/current no-overall /'
