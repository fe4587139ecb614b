#!/bin/bash
# 1-GPU box: previously failing tests, then q4 turn-taking / one-tile sweeps
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_round2.py -m gpu -q -k "table_matches or sharded_cuda or paired or row_ranges or vs_oracle" > gpurun_out/r2_pytest7.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2_pytest7.log | cut -c1-300
rm -f gpurun_out/r2_attn_turn.jsonl
ab() { tag=$1; shift; env "$@" timeout 120 python tools/attn_bench.py --tag "$tag" $EXTRA 2>&1 | tail -1 | cut -c1-230 | tee -a gpurun_out/r2_attn_turn.jsonl; }
EXTRA=""
ab q4-turn0-poly3 TF_EXT_ATTN_TURN=0
for poly in 0 2 3 4 5; do ab q4-turn1-poly$poly TF_EXT_ATTN_TURN=1 TF_EXT_ATTN_POLY=$poly; done
for poly in 2 3 4; do ab q4-1tile-2cta-poly$poly TF_EXT_ATTN_TILES=1 TF_EXT_ATTN_POLY=$poly; done
EXTRA="--S 2304 --dim 640 --heads 10"
ab sd21-turn0 TF_EXT_ATTN_TURN=0
ab sd21-turn1-poly4 TF_EXT_ATTN_TURN=1
ab sd21-turn1-poly3 TF_EXT_ATTN_TURN=1 TF_EXT_ATTN_POLY=3
ab sd21-1tile-poly4 TF_EXT_ATTN_TILES=1
