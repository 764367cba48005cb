#!/bin/bash
# Round-2 experiment 10 (one GPU): consumed input lines demoted to normal L2 priority (applypriority) in the flat
# elementwise kernel, PH_FLAT_L2_HINT=3 vs the default 2 -- bench.py headline
mkdir -p gpurun_out
for h in 2 3 2 3; do
  PH_FLAT_L2_HINT=$h timeout 200 python bench.py --no-extras --steps 40 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('hint=$h', d['value'], 'add_ms', r['avg_launch_ms'], 'mul_ms', list(r['other_kernels'].values())[0]['avg_launch_ms'], 'fused', d['fused_single_pass']['gbs'], d['parity_spot_check'])
"
done
