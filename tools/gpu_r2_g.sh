#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_stream_gpu.py tests/test_misc_gpu.py -m gpu -q -p no:cacheprovider --timeout 200 2>&1 | tail -3
bash tools/gpu_r2_c5.sh
timeout 600 python bench.py --steps 20 --no-cpu --no-secondary > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err; echo "bench exit $?"
python - <<'PY'
import json
d = json.loads(open("gpurun_out/bench_r2g.json").read().splitlines()[-1])
print({k: d[k] for k in ("value", "ms_per_step")}, d["roofline"]["kernel_ms"], d["roofline"]["frac"])
print({k: v for k, v in d["config5"].items() if k in ("ms_per_batch", "value", "frac_of_aggregate_peak", "parity")})
PY
