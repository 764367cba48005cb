#!/bin/bash
# N-GPU job: agreement script under both transports, then the bench line at N (P2P), optionally with the NCCL forms
mkdir -p gpurun_out
N=${1:-2}
WHAT=${2:-all}
if [ "$WHAT" = all ] || [ "$WHAT" = check ]; then
for mode in p2p nccl; do
  if [ $mode = nccl ]; then export PH_NO_P2P=1; else unset PH_NO_P2P; fi
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2963$N \
    tests/mgpu_check.py > gpurun_out/mgpu_$mode.out 2> gpurun_out/mgpu_$mode.err
  echo "== $mode rc=$?"; tail -2 gpurun_out/mgpu_$mode.out; grep "^\[rank0\]" gpurun_out/mgpu_$mode.err | grep -v Warning | tail -12
done
unset PH_NO_P2P
fi
for mode in p2p nccl; do
  if [ $mode = nccl ]; then
    if [ "$WHAT" != all ] && [ "$WHAT" != ab ]; then continue; fi
    export PH_NO_P2P=1
  else unset PH_NO_P2P; fi
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2973$N \
    bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_n${N}_$mode.json 2> gpurun_out/bench_n${N}_$mode.err
  echo "== bench $mode rc=$?"; grep -v "^W0\|^\*\*\*\*\|OMP_NUM" gpurun_out/bench_n${N}_$mode.err | tail -8
  python - <<PY
import json
for l in open("gpurun_out/bench_n${N}_$mode.json"):
    if l.startswith("{"):
        d = json.loads(l)
        print("value", d["value"], "e2e", d["e2e"]["value"], "naive", d["e2e"]["naive"]["value"], "link", d["e2e"]["host_link_ceiling"]["value"])
        ex = d["extras"]
        if "error" in ex: print(ex)
        else:
            for k, v in ex.items():
                print(k, {kk: vv for kk, vv in v.items() if kk in ("gcell_updates_per_s", "ms_per_step", "field_hash", "subcube_vs_oracle", "gbs", "ms", "result_ok", "result", "expected", "gbs_aggregate", "checksum", "clocks")})
PY
done
