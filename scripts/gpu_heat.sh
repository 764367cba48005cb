#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_heat.py -m gpu -q --timeout 300 -x 2>&1 | tail -2
for s in 1024,1024,1024 2048,2048,2048 256,2048,2048; do
echo "$s: $(timeout 300 python benchmarks/bench_kernels.py --heat-shape $s 2>&1 | grep -o 'gcell_per_s.*')"
done
timeout 900 ncu --set full --import-source on --clock-control none -k regex:"heat_tma2" -s 2 -c 1 -o gpurun_out/prof_tma2_v3 -f python benchmarks/bench_kernels.py --heat-shape 256,2048,2048 --reps 1 > gpurun_out/ncu_tma2.log 2>&1
python benchmarks/ncu_summary.py gpurun_out/prof_tma2_v3.ncu-rep gpurun_out/ncu_heat_tma2_v3.csv; tail -1 gpurun_out/ncu_heat_tma2_v3.csv
