#!/bin/bash
# 1-GPU box: full GPU test tier, paired-kernel sweep, one-tile (two CTAs per SM) sweep, d=80 two-half kernel sweep
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py smoke 2>&1 | tail -3
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_pytest6.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2_pytest6.log | cut -c1-300
rm -f gpurun_out/r2_attn_pairs.jsonl
ab() { tag=$1; shift; env "$@" timeout 200 python tools/attn_bench.py --tag "$tag" $EXTRA 2>&1 | tail -1 | cut -c1-220 | tee -a gpurun_out/r2_attn_pairs.jsonl; }
EXTRA="--inject 1"
for poly in 0 2 3 4; do ab q4d-poly$poly TF_EXT_ATTN_POLY_PAIR=$poly; done
ab inject-unpaired TF_EXT_ATTN_DEDUP=0
EXTRA=""
ab q4-2tile-default TF_X=1
for poly in 2 3 4; do ab q4-1tile-2cta-poly$poly TF_EXT_ATTN_TILES=1 TF_EXT_ATTN_POLY=$poly; done
EXTRA="--S 2304 --dim 640 --heads 10"
ab sd21-q4-2tile TF_X=1
ab sd21-q4-1tile-poly4 TF_EXT_ATTN_TILES=1 TF_EXT_ATTN_POLY=4
EXTRA="--S 1024 --dim 640 --heads 8"
for poly in 0 2 3 4; do ab h2-d80-poly$poly TF_EXT_ATTN_POLY_H2=$poly; done
ab v1-d80 TF_EXT_ATTN_MODE=v1
EXTRA="--S 256 --dim 1280 --heads 8"
ab v1-d160 TF_X=1
