#!/bin/bash
# ncu --set full of the kernels written / changed this round (one launch each), summarised with benchmarks/ncu_summary.py
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"map_flat_kernel" -s 10 -c 2 -o /tmp/r02_bench_kernels -f \
  python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_full_bench.log 2>&1
python benchmarks/ncu_summary.py /tmp/r02_bench_kernels.ncu-rep gpurun_out/r02_ncu_bench_kernels.csv; cut -c1-400 gpurun_out/r02_ncu_bench_kernels.csv
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"axis_strip_staged|axis_row_kernel|sum_partial_kernel|ext_partial_kernel" -c 12 -o /tmp/r02_reduce_kernels -f \
  python benchmarks/bench_kernels.py --only "reduce" --reps 1 --warm 0 --inner 1 > gpurun_out/ncu_full_reduce.log 2>&1
python benchmarks/ncu_summary.py /tmp/r02_reduce_kernels.ncu-rep gpurun_out/r02_ncu_reduce_kernels.csv; cut -c1-330 gpurun_out/r02_ncu_reduce_kernels.csv
