#!/bin/bash
# one GPU: full parity suite, C++ spec incl. the device-side I/O block, host-layer bench, the whole per-kernel table
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
grep -E "^(FAILED|ERROR)|passed|failed|pytest exit" gpurun_out/pytest_gpu.log | head -20
PH_SPEC_DEVICE_IO=1 timeout 100 ./tests/cpp/device_narray_spec > gpurun_out/cpp_spec.log 2>&1; echo "cpp spec exit $?"; grep -E "FAIL|expectations|I/O" gpurun_out/cpp_spec.log | head
timeout 100 ./tests/cpp/device_narray_bench > gpurun_out/cpp_bench.jsonl 2>&1; cut -c1-200 gpurun_out/cpp_bench.jsonl
timeout 600 python benchmarks/bench_kernels.py --big-heat > gpurun_out/kernels.jsonl 2> gpurun_out/kernels.err; wc -l gpurun_out/kernels.jsonl; tail -3 gpurun_out/kernels.err
python - <<'PY'
import json
for l in open("gpurun_out/kernels.jsonl"):
    d = json.loads(l)
    if "gbs" in d: print(f'{d["gbs"]:8.1f} {d["frac_measured"]:.3f}  {d["kernel"][:110]}')
    else: print(l[:200])
PY
