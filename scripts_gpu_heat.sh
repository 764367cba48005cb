#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_heat.py tests/test_gpu_example.py -m gpu -q --timeout 300 -x 2>&1 | tail -6
for cfg in 0 1 2 3 4; do
echo "TB cfg $cfg"
PH_HEAT_TB_CFG=$cfg timeout 300 python benchmarks/bench_kernels.py --heat-shape 1024,1024,1024 2>&1 | cut -c1-220
PH_HEAT_TB_CFG=$cfg timeout 300 python benchmarks/bench_kernels.py --heat-shape 2048,2048,2048 2>&1 | cut -c1-220
done
PH_HEAT_NO_FUSE2=1 timeout 300 python benchmarks/bench_kernels.py --heat-shape 2048,2048,2048 2>&1 | cut -c1-220
timeout 900 ncu --set full --clock-control none -k regex:"heat_tma2" -s 2 -c 1 -o gpurun_out/prof_tma2_v2 -f python benchmarks/bench_kernels.py --heat-shape 256,2048,2048 --reps 1 > gpurun_out/ncu_tma2.log 2>&1
python benchmarks/ncu_summary.py gpurun_out/prof_tma2_v2.ncu-rep gpurun_out/ncu_heat_tma2_v2.csv; cat gpurun_out/ncu_heat_tma2_v2.csv
