#!/bin/bash
# Final evidence job of the round (one GPU): parity tests, C++ spec + host-layer bench, smoke, both bench arms,
# per-kernel table, sharded ops at world size 1, ncu launch list of the bench command.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --timeout 500 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | head -20
timeout 100 ./tests/cpp/device_narray_spec > gpurun_out/cpp_spec.log 2>&1; echo "cpp spec exit $?"; grep -E "FAIL|expectations" gpurun_out/cpp_spec.log | head
timeout 100 ./tests/cpp/device_narray_bench > gpurun_out/cpp_bench.jsonl 2>&1; cut -c1-200 gpurun_out/cpp_bench.jsonl
timeout 100 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 300 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; cut -c1-300 gpurun_out/bench.json; tail -3 gpurun_out/bench.err
timeout 200 python bench.py --impl reference --steps 10 --warmup 3 > gpurun_out/bench_ref.json 2>&1; cut -c1-200 gpurun_out/bench_ref.json
timeout 400 python benchmarks/bench_kernels.py --big-heat > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; wc -l gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 1 --master-addr 127.0.0.1 --master-port 29713 benchmarks/bench_sharded_ops.py --quick --reps 2 2>&1 | grep "^{" | cut -c1-200
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/ncu_bench.log 2>&1; grep -c map_flat gpurun_out/launches.csv
