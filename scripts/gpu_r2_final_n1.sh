#!/bin/bash
# Final evidence job of round 2 (one GPU): parity suite, C++ specs + host-layer bench, smoke, both bench arms,
# per-kernel table, ncu launch list and warm DRAM bytes of the bench command.
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | head -20
timeout 100 ./tests/cpp/device_narray_spec > gpurun_out/cpp_spec.log 2>&1; echo "cpp spec exit $?"; grep -E "FAIL|expectations" gpurun_out/cpp_spec.log | head
timeout 100 ./tests/cpp/device_narray_bench > gpurun_out/cpp_bench.jsonl 2>&1; cut -c1-160 gpurun_out/cpp_bench.jsonl
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 600 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/bench_ref.json 2>&1; cut -c1-200 gpurun_out/bench_ref.json
timeout 600 python benchmarks/bench_kernels.py --big-heat > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; wc -l gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; grep -c map_flat gpurun_out/launches.csv
timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -c 40 --csv --log-file gpurun_out/bench_dram_warm.csv python bench.py --steps 5 --warmup 3 --no-extras > /dev/null 2>&1; grep -c map_flat gpurun_out/bench_dram_warm.csv
python - <<'PY'
import json
for l in open("gpurun_out/kernels.jsonl"):
    d = json.loads(l)
    if "gbs" in d: print(f'{d["gbs"]:8.1f} {d["frac_measured"]:.3f}  {d["kernel"][:110]}')
    else: print(l[:220].rstrip())
PY
