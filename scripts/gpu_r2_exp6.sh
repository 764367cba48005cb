#!/bin/bash
# Round-2 experiment 6 (one GPU): ncu --set full of the staged strip kernel on [16384,16384] f32, one / two columns per lane
mkdir -p gpurun_out
for cpt in 1 2; do
  PH_AXIS_STAGED_CPT=$cpt timeout 300 ncu --set full --clock-control none --import-source on -k regex:"axis_strip_staged" -c 1 -o gpurun_out/exp6_staged_f32_cpt$cpt -f \
    python benchmarks/bench_kernels.py --only "reduce axis=0 sum f32" --reps 1 --warm 0 --inner 1 > gpurun_out/exp6_cpt$cpt.log 2>&1
  tail -2 gpurun_out/exp6_cpt$cpt.log
done
ls -la gpurun_out/*.ncu-rep
