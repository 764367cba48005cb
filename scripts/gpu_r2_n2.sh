#!/bin/bash
# 2-GPU job: single-GPU parity tests on GPU 0, N-GPU agreement under both transports, bench at N=2 (P2P and NCCL forms)
mkdir -p gpurun_out
CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
bash scripts/gpu_r2_multi.sh 2 all
