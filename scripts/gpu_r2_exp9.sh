#!/bin/bash
# Round-2 experiment 9 (one GPU): does an L2 set-aside for persisting lines change how much of the temporary of
# a*b+c survives until its consumer?  bench.py headline at several set-aside sizes.
mkdir -p gpurun_out
for mb in none 0 32 64 128; do
  if [ $mb = none ]; then unset PH_L2_PERSIST_MB; else export PH_L2_PERSIST_MB=$mb; fi
  timeout 200 python bench.py --no-extras --steps 40 --warmup 5 2>/dev/null | python -c "
import sys,json
d=json.loads(sys.stdin.read().strip().splitlines()[-1])
r=d['roofline']
print('persist_mb=$mb', d['value'], 'add_ms', r['avg_launch_ms'], 'mul_ms', list(r['other_kernels'].values())[0]['avg_launch_ms'], 'fused', d['fused_single_pass']['gbs'])
"
done
