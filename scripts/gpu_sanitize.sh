#!/bin/bash
mkdir -p gpurun_out
# memcheck + racecheck over the small-shape parity tests (skip the full-size ones: sanitizer is ~20-50x slower)
timeout 2400 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file gpurun_out/sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_index.py tests/test_gpu_heat.py tests/test_gpu_reduce.py tests/test_gpu_ewise.py -m gpu -q -x --timeout 1200 \
  -k "not large and not 1e9 and not 100_steps and not tolerance and not 1_000_003" > gpurun_out/sanitizer_pytest.log 2>&1
echo "memcheck exit $?" >> gpurun_out/sanitizer_pytest.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file gpurun_out/sanitizer_racecheck.log \
  python -m pytest tests/test_gpu_heat.py -m gpu -q -x --timeout 1200 -k "two_step or 3d_one_step or slab" >> gpurun_out/sanitizer_pytest.log 2>&1
echo "racecheck exit $?" >> gpurun_out/sanitizer_pytest.log
grep -E "passed|failed|exit" gpurun_out/sanitizer_pytest.log; grep -E "ERROR SUMMARY|Invalid|out of bounds|hazard" gpurun_out/sanitizer_memcheck.log gpurun_out/sanitizer_racecheck.log | head -10
