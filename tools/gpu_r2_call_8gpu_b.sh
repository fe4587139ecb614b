#!/bin/bash
# 8-GPU box: C2 with the dual-stream schedule (verify) and with the fused schedule, back to back
mkdir -p gpurun_out
TAG=n8_c2_dual TIMEOUT=300 tools/gpu_multi.sh 8 --steps 5 --warmup 3 --verify --dual-stream 1 | cut -c1-900
TAG=n8_c2_fused2 TIMEOUT=240 tools/gpu_multi.sh 8 --steps 5 --warmup 3 --dual-stream 0 --skip-e2e | cut -c1-500
