#!/bin/bash
# A/B of the L2 eviction-priority hints in chained elementwise launches (bench headline: t = a*b ; out = t + c)
mkdir -p gpurun_out
for h in 0 1 2 0 2; do
  PH_FLAT_L2_HINT=$h timeout 300 python bench.py --steps 200 --warmup 20 --no-extras 2>/dev/null | python -c "
import sys, json
d = json.loads([l for l in sys.stdin if l.startswith('{')][0])
r = d['roofline']
print('hint $h value', d['value'], 'ms', d['ms_per_step'], 'add', r['avg_launch_ms'], 'mul', list(r['other_kernels'].values())[0]['avg_launch_ms'], 'fused', d['fused_single_pass']['gbs'], 'parity', d['parity_spot_check'])"
done
for h in 0 2; do
  PH_FLAT_L2_HINT=$h timeout 200 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --cache-control none --clock-control none -c 40 --csv --log-file gpurun_out/l2hint_$h.csv python bench.py --steps 5 --warmup 3 --no-extras > /dev/null 2>&1
  python - <<PY
import csv
rows = [r for r in csv.reader(l for l in open("gpurun_out/l2hint_$h.csv") if not l.startswith("=="))]
h = rows[0]; ki, mi, vi, idi = h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value"), h.index("ID")
acc = {}
for r in rows[1:]:
    acc.setdefault((int(r[idi]), r[ki][:60]), {})[r[mi]] = r[vi]
for (i, k), m in sorted(acc.items())[-6:]:
    print("hint $h", i, k, m.get("gpu__time_duration.sum"), "ns read", m.get("dram__bytes_read.sum"), "write", m.get("dram__bytes_write.sum"))
PY
done
PH_FLAT_L2_HINT=2 timeout 600 python -m pytest tests/test_gpu_ewise.py -q -x 2>&1 | tail -2
