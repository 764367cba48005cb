#!/bin/bash
mkdir -p gpurun_out
timeout 60 python -m pytest tests/test_gpu_index.py -m gpu -q -x -k "slices" 2>&1 | tail -3
timeout 30 ./tests/cpp/device_narray_spec 2>&1 | grep -E "FAIL|expectations"
timeout 30 ./tests/cpp/device_narray_bench > gpurun_out/cpp_bench.jsonl 2>&1; grep slices gpurun_out/cpp_bench.jsonl | cut -c1-200
timeout 60 python benchmarks/bench_kernels.py --only slices 2>&1 | cut -c1-230
