#!/bin/bash
# round 2: whole GPU suite, smoke, bench line (N = 1)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_r2f.log 2>&1
echo "pytest exit $?"; tail -5 gpurun_out/pytest_gpu_r2f.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/bench_r2f.json 2> gpurun_out/bench_r2f.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r2f.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/bench_r2f.json"))
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"], d["e2e"]["value"])
print(d["config4"]["Q1"], d["config4"]["Q64"], d["config4"]["parity"])
print({k: v for k, v in d["config5"].items() if k in ("ms_per_batch", "value", "frac_of_aggregate_peak", "parity")})
PY
