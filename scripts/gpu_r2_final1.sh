#!/bin/bash
# one GPU: sanitizer over the kernels written this round, then the evidence captures of the bench command
mkdir -p gpurun_out
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_reduce.py -m gpu -q -x \
  -k "staged or strided_views or axis_reductions or short_last or sharded_entry or full_sum_bit_exact" > gpurun_out/sanitizer_memcheck_r02.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitizer_memcheck_r02.log | head -5
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_gpu_reduce.py -m gpu -q -x \
  -k "staged or strided_views" > gpurun_out/sanitizer_racecheck_r02.log 2>&1; echo "racecheck exit $?"
grep -E "RACECHECK SUMMARY|passed|failed|hazard" gpurun_out/sanitizer_racecheck_r02.log | head -5
timeout 300 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_gpu_heat.py -m gpu -q -x > gpurun_out/sanitizer_memcheck_heat_r02.log 2>&1; echo "memcheck heat exit $?"
grep -E "ERROR SUMMARY|passed|failed|Invalid" gpurun_out/sanitizer_memcheck_heat_r02.log | head -5
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>&1; cut -c1-400 gpurun_out/bench_ref.json
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; grep -c map_flat gpurun_out/launches.csv
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -c 40 --csv --log-file gpurun_out/bench_dram_warm.csv python bench.py --steps 5 --warmup 3 --no-extras > /dev/null 2>&1; grep -c map_flat gpurun_out/bench_dram_warm.csv
