#!/bin/bash
# Evidence job: kernel table, bench (both arms), ncu launch list + full captures condensed to CSV on the box.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_example.py tests/test_gpu_io.py -m gpu -q 2>&1 | tail -2
timeout 900 python benchmarks/bench_kernels.py --big-heat > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; wc -l gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-200 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 20 --warmup 3 > gpurun_out/bench_ref.json 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1
timeout 1200 ncu --set full --clock-control none -k regex:"map_flat|map_rows" -s 8 -c 4 -o /tmp/prof_bench -f python bench.py --steps 3 --warmup 3 --no-extras > gpurun_out/ncu_full1.log 2>&1
python benchmarks/ncu_summary.py /tmp/prof_bench.ncu-rep gpurun_out/ncu_bench_kernels.csv
timeout 1500 ncu --set full --clock-control none -k regex:"copy_rows|transpose|sum_partial|ext_partial|axis_strip|axis_row|heat_tma|heat_march|mask_set|map_" -c 60 -o /tmp/prof_kernels -f python benchmarks/bench_kernels.py --quick --reps 1 > gpurun_out/ncu_full2.log 2>&1
python benchmarks/ncu_summary.py /tmp/prof_kernels.ncu-rep gpurun_out/ncu_all_kernels.csv
ls -la gpurun_out
