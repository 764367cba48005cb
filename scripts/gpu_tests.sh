#!/bin/bash
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit|Error" gpurun_out/pytest_gpu.log | head -40
timeout 600 python benchmarks/bench_kernels.py --only "axis=2" 2>&1 | cut -c1-200
timeout 600 python benchmarks/bench_kernels.py --only "ewise" 2>&1 | cut -c1-200
