#!/bin/bash
mkdir -p gpurun_out
# memcheck over the small-shape parity tests (skip the full-size ones: sanitizer is ~20-50x slower)
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_index.py tests/test_gpu_heat.py tests/test_gpu_reduce.py -m gpu -q -x --timeout 1200 \
  -k "not large and not 1e9 and not 100_steps and not tolerance" > gpurun_out/sanitizer_pytest.log 2>&1
echo "sanitizer exit $?" >> gpurun_out/sanitizer_pytest.log
tail -5 gpurun_out/sanitizer_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds" gpurun_out/sanitizer_memcheck.log | head -10
