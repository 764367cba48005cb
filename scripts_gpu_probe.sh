#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --timeout 600 -x 2>&1 | tail -2
timeout 300 python benchmarks/bench_kernels.py --only tile 2>&1 | cut -c1-200
timeout 300 python benchmarks/bench_kernels.py --only gather 2>&1 | cut -c1-200
timeout 300 ncu --metrics gpu__time_duration.sum,launch__grid_size,launch__registers_per_thread,dram__bytes_read.sum,dram__bytes_write.sum,smsp__inst_executed.sum --clock-control none --csv --log-file gpurun_out/tile_launches.csv python benchmarks/bench_kernels.py --only "tile [4096" --reps 1 --warm 0 --inner 1 > /dev/null 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/tile_launches.csv')))
hdr=[i for i,r in enumerate(rows) if r and r[0]=='ID'][0]
H=rows[hdr]; out={}
for r in rows[hdr+1:]:
    d=dict(zip(H,r)); out.setdefault((d['ID'],d['Kernel Name'][:90]),{})[d['Metric Name']]=d['Metric Value']
for k,v in list(out.items())[-2:]: print(k, v)
PY
