#!/bin/bash
# round-2 closing pass (1 GPU): whole suite, smoke, sanitizers on every kernel, ncu launch list + full capture of the
# filtered box scan, the bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -p no:cacheprovider --timeout 300 > gpurun_out/pytest_gpu_final.log 2>&1
echo "pytest exit $?"; tail -3 gpurun_out/pytest_gpu_final.log
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -1
for tool in memcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 python tools/sanitize_cases.py > gpurun_out/sanitize_$tool.log 2>&1
  echo "$tool exit $?"; tail -3 gpurun_out/sanitize_$tool.log
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:^k_ -c 600 --csv \
    --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --no-cpu > gpurun_out/r02_launches_bench.log 2>&1
echo "launch list exit $?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_box_filter -s 3 -c 1 \
    -f -o gpurun_out/box_filter_r2 python tools/box_q64_probe.py > gpurun_out/ncu_box_filter.log 2>&1
echo "box filter capture exit $?"
timeout 900 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err; echo "bench exit $?"; tail -2 gpurun_out/bench_final.err
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_final.json").read().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"], d["e2e"]["value"])
print(d["config4"]["Q1"]["ms_per_query_pass"], d["config4"]["Q64"]["ms_per_query_pass"], d["config4"]["Q64"]["frac_of_aggregate_peak"], d["config4"]["parity"]["count_mismatches"], d["config5"]["ms_per_batch"], d["config5"]["parity"]["argmax_mismatches"])
PY
