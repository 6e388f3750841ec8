#!/bin/bash
# N ranks: host-logic check of every sharded path, then the bench line (configs 2, 4, 5 at this world size)
N=${N:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    tools/multi_gpu_check.py > gpurun_out/multi_gpu_check_r2_n$N.json 2> gpurun_out/multi_r2_n$N.err
echo "multi_gpu_check exit $?"; grep -v "^NCCL" gpurun_out/multi_gpu_check_r2_n$N.json | tail -1 | cut -c1-1500
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/bench_r2_n$N.json 2> gpurun_out/bench_r2_n$N.err
echo "bench exit $?"
python - <<PY
import json
d = json.loads(open("gpurun_out/bench_r2_n$N.json").read().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["kernel_ms"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["stats_check"])
print(d["config4"].get("Q1"), d["config4"].get("Q64"), d["config4"].get("parity"))
print({k: v for k, v in d["config5"].items() if k in ("ms_per_batch", "value", "frac_of_aggregate_peak", "parity", "collective", "error")})
PY
