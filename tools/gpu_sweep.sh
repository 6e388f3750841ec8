#!/bin/bash
# One gpurun visit: stream-kernel GPU tests, kernel timings on three shapes, a light ncu pass.
mkdir -p gpurun_out
rm -f gpurun_out/stream_sweep.jsonl
timeout 420 python -m pytest tests/test_stream_gpu.py -m gpu -q --maxfail=60 --durations=8 > gpurun_out/pytest_stream.log 2>&1
echo "pytest exit $?" | tee -a gpurun_out/pytest_stream.log
tail -5 gpurun_out/pytest_stream.log
timeout 240 python tools/stream_sweep.py --steps 20 2>&1 | cut -c1-300                                   # config 2
timeout 240 python tools/stream_sweep.py --steps 20 --n-inst 32768 --n-steps 500 2>&1 | cut -c1-300    # config 5 shape
timeout 240 python tools/stream_sweep.py --steps 20 --n-inst 1024 2>&1 | cut -c1-300                   # one e2e chunk
M=gpu__time_duration.sum,smsp__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active
M=$M,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum
M=$M,sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,dram__bytes_write.sum
for s in short_scoreboard wait long_scoreboard math_pipe_throttle mio_throttle not_selected branch_resolving dispatch_stall no_instruction; do
  M=$M,smsp__average_warps_issue_stalled_${s}_per_issue_active.ratio
done
timeout 300 ncu --metrics $M --clock-control none -k regex:k_stream_rows --csv --log-file gpurun_out/sweep_ncu.csv \
    python tools/stream_sweep.py --once > gpurun_out/sweep_ncu.log 2>&1
echo "ncu exit $?"; tail -3 gpurun_out/sweep_ncu.log | cut -c1-300
if [ -n "$FULL" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_stream_rows -c 1 -f -o gpurun_out/rows_${FULL} \
    python tools/stream_sweep.py --once > gpurun_out/ncu_rows_full.log 2>&1
echo "full capture exit $?"
fi
