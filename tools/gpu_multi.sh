#!/bin/bash
# Multi-GPU bench (one process per GPU, NCCL): usage  N=2 bash tools/gpu_multi.sh
N=${N:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt
NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT,COLL timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N \
    --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps ${STEPS:-20} --warmup 3 \
    > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "bench N=$N exit $?"
grep -h '"metric"' gpurun_out/bench_n$N.json | tail -1
grep -hE "NVLS|Connected all|Channel 00|via P2P" gpurun_out/bench_n$N.json gpurun_out/bench_n$N.err | head -8
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --impl reference --gpus $N --steps 1 --warmup 0 > gpurun_out/bench_ref_n$N.json 2>&1
echo "reference arm N=$N exit $?"; tail -1 gpurun_out/bench_ref_n$N.json | cut -c1-300
