#!/bin/bash
# usage: gpu_multi.sh N [extra bench args]   — N-rank bench under torchrun, output in gpurun_out/
N=$1; shift
mkdir -p gpurun_out
tag=${TAG:-n$N}
timeout ${TIMEOUT:-420} python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
   bench.py --gpus $N "$@" > gpurun_out/r2_bench_$tag.json 2> gpurun_out/r2_bench_$tag.err
echo "exit $? ($tag)"; grep -v "^W\|^$" gpurun_out/r2_bench_$tag.err | tail -4; cut -c1-2500 gpurun_out/r2_bench_$tag.json
