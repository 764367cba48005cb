#!/bin/bash
# 1-GPU job: parity tests, the view / few-column reduction rows (new kernels vs PH_REDUCE_GATHER=1 / PH_AXIS_STAGED=0),
# their DRAM bytes under ncu, full reductions through the record path
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q --timeout 600 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/pytest_gpu.log
timeout 600 python benchmarks/bench_kernels.py --only "reduce" > gpurun_out/view_reduce_inplace.jsonl 2> gpurun_out/view_reduce.err
echo "inplace rc=$?"; cut -c1-200 gpurun_out/view_reduce_inplace.jsonl; tail -3 gpurun_out/view_reduce.err
PH_REDUCE_GATHER=1 PH_AXIS_STAGED=0 timeout 600 python benchmarks/bench_kernels.py --only "reduce" > gpurun_out/view_reduce_gather.jsonl 2>> gpurun_out/view_reduce.err
echo "gather rc=$?"; cut -c1-200 gpurun_out/view_reduce_gather.jsonl
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv \
  -k regex:"sum_partial|ext_partial|axis_" --log-file gpurun_out/ncu_view_reduce.csv python benchmarks/bench_kernels.py --only "reduce" --reps 1 --warm 0 --inner 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/ncu_view_reduce.csv") if not l.startswith("=="))]
h = rows[0]; ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
acc = {}
for r in rows[1:]:
    acc.setdefault((int(r[idi]), r[ki][:70]), {})[r[mi]] = r[vi]
for (i, k), m in sorted(acc.items()):
    print(i, k, m.get("gpu__time_duration.sum"), "ns read", m.get("dram__bytes_read.sum"), "write", m.get("dram__bytes_write.sum"))
PY
