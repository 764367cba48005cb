#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | head -40
timeout 900 python benchmarks/bench_kernels.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; cut -c1-220 gpurun_out/kernels.jsonl; tail -5 gpurun_out/kernels.err
PH_HEAT_GROUP_BYTES=16 timeout 300 python benchmarks/bench_kernels.py --only heat > gpurun_out/kernels_heat16.jsonl 2>&1; cut -c1-220 gpurun_out/kernels_heat16.jsonl
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"heat_march|copy_rows|axis_row|ext_partial" -c 12 -o gpurun_out/prof_r1b -f python benchmarks/bench_kernels.py --quick --reps 1 > gpurun_out/ncu_full.log 2>&1; tail -3 gpurun_out/ncu_full.log
