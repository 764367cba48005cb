#!/bin/bash
# Evidence job: all GPU tests, kernel table, bench (both arms), ncu launch list + full captures condensed to CSV.
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit|Error" gpurun_out/pytest_gpu.log | head -20
timeout 900 python benchmarks/bench_kernels.py --big-heat > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; wc -l gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>&1; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
# DRAM bytes of the two bench kernels in their natural L2 state (one pass, caches NOT flushed between launches)
timeout 600 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -k regex:"map_flat" -s 8 -c 8 --csv --log-file gpurun_out/bench_dram_warm.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_bench2.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"map_flat|map_rows" -s 8 -c 4 -o /tmp/prof_bench -f python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_full1.log 2>&1
python benchmarks/ncu_summary.py /tmp/prof_bench.ncu-rep gpurun_out/ncu_bench_kernels.csv
timeout 1500 ncu --set full --clock-control none -k regex:"copy_rows|transpose|sum_partial|ext_partial|axis_strip|axis_row|heat_tma|heat2d|heat_march|mask_set" -c 60 -o /tmp/prof_kernels -f python benchmarks/bench_kernels.py --quick --reps 1 --warm 0 --inner 1 > gpurun_out/ncu_full2.log 2>&1
python benchmarks/ncu_summary.py /tmp/prof_kernels.ncu-rep gpurun_out/ncu_all_kernels.csv
ls -la gpurun_out | tail -14
