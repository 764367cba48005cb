#!/bin/bash
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none -k regex:"axis_rowreg|heat2d|heat_tma2|transpose|axis_strip_kernel<float, 8, 3>" -c 16 -o /tmp/prof_kernels2 -f python benchmarks/bench_kernels.py --quick --reps 1 --warm 0 --inner 1 > gpurun_out/ncu_full3.log 2>&1
python benchmarks/ncu_summary.py /tmp/prof_kernels2.ncu-rep gpurun_out/ncu_all_kernels2.csv
cut -c1-330 gpurun_out/ncu_all_kernels2.csv
timeout 900 ncu --set full --clock-control none -k regex:"heat_tma2" -s 1 -c 1 -o /tmp/prof_tma2_final -f python benchmarks/bench_kernels.py --heat-shape 256,2048,2048 --reps 1 --warm 0 --inner 1 > gpurun_out/ncu_tma2.log 2>&1
python benchmarks/ncu_summary.py /tmp/prof_tma2_final.ncu-rep gpurun_out/ncu_heat_tma2_final.csv; tail -1 gpurun_out/ncu_heat_tma2_final.csv
