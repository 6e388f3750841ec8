#!/bin/bash
# One gpurun visit (1 GPU): all GPU tests, stream sweep, light ncu metrics.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --maxfail=20 --durations=6 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?"; tail -25 gpurun_out/pytest_gpu.log
rm -f gpurun_out/stream_sweep.jsonl
timeout 240 python tools/stream_sweep.py --steps 20 2>&1 | cut -c1-300
timeout 240 python tools/stream_sweep.py --steps 20 --n-inst 32768 --n-steps 500 2>&1 | cut -c1-300
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
timeout 300 ncu --metrics $M --clock-control none -k regex:k_stream_rows --csv --log-file gpurun_out/sweep_ncu.csv \
    python tools/stream_sweep.py --once > gpurun_out/sweep_ncu.log 2>&1
grep -E "inst_executed|time_duration|wavefronts" gpurun_out/sweep_ncu.csv | cut -d, -f13- | head
