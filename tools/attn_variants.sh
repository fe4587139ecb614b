#!/bin/bash
# Sweep the extended-attention kernel variants at the C2 top-level shape (and the d=64 SD2.1 shapes).
# usage: tools/attn_variants.sh > gpurun_out/attn_variants.jsonl
run() { tag=$1; shift; env "$@" timeout 300 python tools/attn_bench.py --tag "$tag" $EXTRA || echo "{\"tag\": \"$tag\", \"failed\": $?}"; }
EXTRA=""
run pp-ones1-poly0 TF_EXT_ATTN_MODE=pp TF_EXT_ATTN_ONES=1 TF_EXT_ATTN_POLY=0
run q4-ones1-poly3 TF_EXT_ATTN_MODE=q4 TF_EXT_ATTN_ONES=1 TF_EXT_ATTN_POLY=3
for poly in 0 2 3 4 5 6 8; do run q4s-ones1-poly$poly TF_EXT_ATTN_ONES=1 TF_EXT_ATTN_POLY=$poly; done
for poly in 0 3 4 5; do run q4s-ones0-poly$poly TF_EXT_ATTN_ONES=0 TF_EXT_ATTN_POLY=$poly; done
EXTRA="--inject 1"; run q4s-default-inject TF_X=1
EXTRA="--video-like 0"; run q4s-default-iid TF_X=1
EXTRA="--S 2304 --dim 640 --heads 10"
run sd21-q4-poly4 TF_EXT_ATTN_MODE=q4 TF_EXT_ATTN_POLY=4
for poly in 0 3 4 5 6; do run sd21-q4s-poly$poly TF_EXT_ATTN_POLY=$poly; done
EXTRA="--S 9216 --dim 320 --heads 5"; run sd21-top-q4s TF_X=1
EXTRA="--S 576 --dim 1280 --heads 20"; run sd21-ragged-q4s TF_X=1
