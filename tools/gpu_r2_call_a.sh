#!/bin/bash
# 2-GPU box: full GPU test tier, paired-kernel sweep, 2-rank bench with verify (token split on / off)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest6.log | cut -c1-300
for poly in 0 2 3 4; do
  TF_EXT_ATTN_POLY_PAIR=$poly timeout 200 python tools/attn_bench.py --inject 1 --tag "q4d-poly$poly" 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/r2_attn_pairs.jsonl
done
TF_EXT_ATTN_DEDUP=0 timeout 200 python tools/attn_bench.py --inject 1 --tag "inject-unpaired" 2>&1 | tail -1 | cut -c1-200 | tee -a gpurun_out/r2_attn_pairs.jsonl
TAG=n2_tokensplit timeout 400 tools/gpu_multi.sh 2 --steps 4 --warmup 3 --verify --no-gpu-reference | cut -c1-1200
