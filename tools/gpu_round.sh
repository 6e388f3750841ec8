#!/bin/bash
# One gpurun visit: GPU tests, smoke, bench.  Each stage has its own timeout; logs land in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.used --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --durations=10 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1
echo "smoke exit $?" | tee -a gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
timeout 900 python bench.py ${BENCH_ARGS:---steps 20} > gpurun_out/bench.json 2> gpurun_out/bench.err
echo "bench exit $?"
cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
