#!/bin/bash
# round 2: bitmap-filter box scan — tests, timing at config-4 size, full bench line
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_box_gpu.py -m gpu -q --maxfail=20 -p no:cacheprovider --timeout 200 > gpurun_out/pytest_box_r2b.log 2>&1
echo "pytest exit $?"; tail -8 gpurun_out/pytest_box_r2b.log
timeout 200 python tools/box_q64_probe.py 2>&1 | tail -4
timeout 900 python bench.py > gpurun_out/bench_r2b.json 2> gpurun_out/bench_r2b.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r2b.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2b.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"], d["e2e"]["value"])
print(d["config4"]["Q1"], d["config4"]["Q64"], d["config4"]["parity"])
PY
