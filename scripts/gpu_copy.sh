#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_index.py -m gpu -q --timeout 300 -x 2>&1 | tail -2
timeout 300 python benchmarks/bench_kernels.py --only "transposed" 2>&1 | cut -c1-200
timeout 300 python -m pytest tests/test_gpu_fullsize.py -m gpu -q --timeout 300 -x -k config2 2>&1 | tail -2
