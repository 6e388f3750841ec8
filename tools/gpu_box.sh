#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_box_gpu.py tests/test_misc_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout 200 > gpurun_out/pytest_box_r2c.log 2>&1
echo "pytest exit $?"; tail -4 gpurun_out/pytest_box_r2c.log
timeout 200 python tools/box_q64_probe.py 2>&1 | tail -6
for v in ab_variants/*/; do
  [ -f $v/libdcarl_b200.so ] || continue
  echo $v; DCARL_B200_LIB=$PWD/$v/libdcarl_b200.so timeout 200 python tools/box_q64_probe.py 2>&1 | grep "^64"
done
