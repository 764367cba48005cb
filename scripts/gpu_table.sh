#!/bin/bash
# Per-kernel table of the round (every BASELINE config + the 8(f) rows), one GPU.
mkdir -p gpurun_out
timeout 900 python benchmarks/bench_kernels.py --big-heat > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; echo "exit $?"; wc -l gpurun_out/kernels.jsonl; tail -5 gpurun_out/kernels.err
cut -c1-260 gpurun_out/kernels.jsonl
timeout 300 python -m pytest tests/test_gpu_heat.py -m gpu -q -k "reference_operators" 2>&1 | tail -3
timeout 120 ./tests/cpp/device_narray_spec 2>&1 | grep -E "FAIL|expectations"
