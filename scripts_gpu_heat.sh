#!/bin/bash
mkdir -p gpurun_out
for ty in 16 32; do
echo "TMA rows $ty"
PH_HEAT_TMA_ROWS=$ty timeout 300 python -m pytest tests/test_gpu_heat.py -m gpu -q --timeout 120 -x 2>&1 | tail -1
PH_HEAT_TMA_ROWS=$ty timeout 300 python benchmarks/bench_kernels.py --heat-shape 1024,1024,1024 2>&1 | cut -c1-200
PH_HEAT_TMA_ROWS=$ty timeout 300 python benchmarks/bench_kernels.py --heat-shape 2048,2048,2048 2>&1 | cut -c1-200
done
PH_HEAT_TMA_ROWS=16 timeout 900 ncu --set full --clock-control none -k regex:"heat_tma" -s 4 -c 1 -o gpurun_out/prof_heat_tma_2k16 -f python benchmarks/bench_kernels.py --heat-shape 256,2048,2048 --reps 1 > gpurun_out/ncu_heat_tma.log 2>&1
