#!/bin/bash
# 8-GPU box: C2 (verify), C3, C4 — one bench line each under gpurun_out/
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
TAG=n8_c2 TIMEOUT=330 tools/gpu_multi.sh 8 --steps 5 --warmup 3 --verify | cut -c1-1500
TAG=n8_c3 TIMEOUT=300 tools/gpu_multi.sh 8 --config C3 --steps 4 --warmup 3 | cut -c1-700
TAG=n8_c4 TIMEOUT=400 tools/gpu_multi.sh 8 --config C4 --steps 4 --warmup 3 | cut -c1-700
