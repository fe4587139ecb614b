#!/bin/bash
# usage: gpu_multi.sh N
N=$1
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
echo "exit $?"; tail -5 gpurun_out/bench_n$N.err; cut -c1-1800 gpurun_out/bench_n$N.json
