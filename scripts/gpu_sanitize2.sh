#!/bin/bash
# memcheck over the kernels changed late in the round (axis_strip plans, rows-kernel block order)
mkdir -p gpurun_out
timeout 115 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_reduce.py tests/test_gpu_index.py -m gpu -q -x -k "axis_reductions or axis_sum_float or short_last or slices_tile or remaining_crystal" > gpurun_out/sanitizer_memcheck2.log 2>&1; echo "exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid|exit" gpurun_out/sanitizer_memcheck2.log | head
