#!/bin/bash
# Final 8-GPU job of round 2: bench line with the peer-store forms, then the sharded-ops table (ordered all-reduce in three launches)
mkdir -p gpurun_out
N=${1:-8}
bash scripts/gpu_r2_multi.sh $N p2ponly
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2964$N \
  benchmarks/bench_sharded_ops.py 2> gpurun_out/sharded_ops_p2p.err | grep "^{" > gpurun_out/sharded_ops_n${N}_p2p.jsonl
python -c "
import json
for l in open('gpurun_out/sharded_ops_n${N}_p2p.jsonl'):
    d=json.loads(l); print(d['ms'], d['aggregate_gbs'], d['op'][:80], d.get('result_ok',''))
for l in open('gpurun_out/bench_n${N}_p2p.json'):
    if l.startswith('{'):
        d=json.loads(l); p=d['extras'].get('multi_gpu_parity',{}); print('parity', {k:p.get(k) for k in ('ok','failed','checks','seconds')})"
