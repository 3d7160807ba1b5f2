#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --no-batched --no-cpu-baseline > gpurun_out/bench_r3s.json 2> gpurun_out/r3s_err.log; tail -3 gpurun_out/r3s_err.log
python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r3s.json').read().strip().splitlines()[-1])
print(d['value'], d['roofline'].get('jacobi_sweeps'))
PY
