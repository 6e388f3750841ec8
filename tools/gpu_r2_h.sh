#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_r2h.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_r2h.log
timeout 900 python bench.py > gpurun_out/bench_r2h.json 2> gpurun_out/bench_r2h.err; echo "bench exit $?"; tail -3 gpurun_out/bench_r2h.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2h.json").read().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"], d["stats_check"])
print(d["config4"]["Q1"]["ms_per_query_pass"], d["config4"]["Q64"]["ms_per_query_pass"], d["config4"]["parity"]["count_mismatches"], d["config5"]["ms_per_batch"], d["config5"]["parity"]["argmax_mismatches"])
print(d.get("cpu_baseline"), d.get("parity_check", None))
PY
