#!/bin/bash
# Lean verification job: GPU parity tests, C++ spec, smoke(), default bench (both arms), launch list.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit|Error" gpurun_out/pytest_gpu.log | head -40
timeout 120 ./tests/cpp/device_narray_spec > gpurun_out/cpp_spec.log 2>&1; echo "cpp spec exit $?"; grep -E "FAIL|expectations" gpurun_out/cpp_spec.log | head -30
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3
timeout 400 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>&1; cut -c1-300 gpurun_out/bench_ref.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; tail -2 gpurun_out/ncu_bench.log | cut -c1-200
