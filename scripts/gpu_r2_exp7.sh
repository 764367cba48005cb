#!/bin/bash
# Round-2 experiment 7 (one GPU): the arithmetic flag word read through a pinned record (one tiny launch + host poll)
# instead of memcpy + memset + stream synchronisation -- parity suite, then the per-axis rows both ways.
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --timeout 500 2>&1 | tail -4
timeout 100 ./tests/cpp/device_narray_spec 2>&1 | grep -E "FAIL|expectations" | head
for form in memcpy record; do
  echo "== flags by $form"
  if [ $form = memcpy ]; then export PH_FLAGS_MEMCPY=1; else unset PH_FLAGS_MEMCPY; fi
  timeout 200 python benchmarks/bench_kernels.py --only "reduce axis" 2>gpurun_out/exp7.err | cut -c1-200 | tee gpurun_out/exp7_$form.jsonl | python -c "
import sys,json
for l in sys.stdin:
    d=json.loads(l[:l.rindex('}')+1]) if l.rstrip().endswith('}') else None
    if d: print(d['kernel'][:70], d['ms_median'], d['gbs'], d['frac_measured'])
"
done
tail -3 gpurun_out/exp7.err
