#!/bin/bash
# One gpurun job: GPU parity tests, smoke, kernel table, bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; tail -2 gpurun_out/smoke.log
timeout 900 python benchmarks/bench_kernels.py > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; cat gpurun_out/kernels.jsonl; tail -5 gpurun_out/kernels.err
timeout 600 python bench.py --steps 50 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-400 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
