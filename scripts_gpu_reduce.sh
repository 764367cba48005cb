#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_index.py tests/test_gpu_fullsize.py -m gpu -q -x --timeout 600 2>&1 | tail -8
timeout 600 python benchmarks/bench_kernels.py --only "axis=2" 2>&1 | cut -c1-200
timeout 600 python benchmarks/bench_kernels.py --only "gather" 2>&1 | cut -c1-200
