#!/bin/bash
# 2-GPU job: the multi-GPU pytest file as the driver would run it on a multi-GPU box, then the same on ONE visible GPU
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q -x --timeout 900 2>&1 | tail -8
CUDA_VISIBLE_DEVICES=0 timeout 1200 python -m pytest tests/test_gpu_multi.py -m gpu -q --timeout 900 2>&1 | tail -8
