#!/bin/bash
# A/B on one box: builds of the same library with different -DDCARL_V_* switches (ab_variants/<name>/), and the
# round-1 kernel (git worktree ab_r1) if present, interleaved on the config-2 workload.
mkdir -p gpurun_out
for rep in 1 2; do
  if [ -d ab_r1 ]; then (cd ab_r1 && timeout 120 python tools/stream_sweep.py --steps 30 2>&1 | grep -o '"ms": [0-9.]*' | sed 's/^/r1           /'); fi
  timeout 120 python tools/stream_sweep.py --steps 30 2>&1 | grep -o '"ms": [0-9.]*' | sed 's/^/current      /'
  for v in ab_variants/*/; do
    n=$(basename $v)
    DCARL_B200_LIB=$PWD/$v/libdcarl_b200.so timeout 120 python tools/stream_sweep.py --steps 30 2>&1 | grep -o '"ms": [0-9.]*\|error.*' | sed "s/^/$n  /"
  done
done
