#!/bin/bash
mkdir -p gpurun_out
for v in "1 16" "2 16" "2 32"; do set -- $v; echo "variant $1 bytes $2"; PH_HEAT_VARIANT=$1 PH_HEAT_GROUP_BYTES=$2 timeout 300 python benchmarks/bench_kernels.py --only "heat 3-D" 2>&1 | cut -c1-200; done
PH_HEAT_VARIANT=1 PH_HEAT_GROUP_BYTES=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"heat_march" -s 4 -c 2 -o gpurun_out/prof_heat_v1 -f python benchmarks/bench_kernels.py --quick --only "heat 3-D" --reps 1 > gpurun_out/ncu_heat1.log 2>&1
PH_HEAT_VARIANT=2 PH_HEAT_GROUP_BYTES=16 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"heat_march" -s 4 -c 2 -o gpurun_out/prof_heat_v2 -f python benchmarks/bench_kernels.py --quick --only "heat 3-D" --reps 1 > gpurun_out/ncu_heat2.log 2>&1
PH_HEAT_VARIANT=2 PH_HEAT_GROUP_BYTES=32 timeout 600 ncu --set full --clock-control none --import-source on -k regex:"heat_march" -s 4 -c 2 -o gpurun_out/prof_heat_v2w -f python benchmarks/bench_kernels.py --quick --only "heat 3-D" --reps 1 > gpurun_out/ncu_heat2w.log 2>&1
ls gpurun_out | grep heat
