#!/bin/bash
# 2-GPU job: the e2e leg with the tapered schedule, with and without the ramp at the front
mkdir -p gpurun_out
for sch in 4,7,0 4,7,5 16,0,0; do
  PH_E2E_SCHEDULE=$sch timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29655 bench.py --gpus 2 --steps 10 --warmup 3 --no-extras 2>/dev/null | python -c "
import sys,json
d=json.loads([l for l in sys.stdin.read().splitlines() if l.startswith('{')][-1])
print('schedule $sch', d['value'], 'e2e', d['e2e']['value'], 'ceiling', d['e2e']['host_link_ceiling']['value'], d['e2e']['frac_of_host_link_ceiling'])
"
done
