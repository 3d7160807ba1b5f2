#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/exp12.log
timeout 900 python bench.py > gpurun_out/bench_r1_l.json 2> gpurun_out/bench_r1_l.err
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/exp12.log 2>&1
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/bench_r1_ref.json 2> gpurun_out/bench_r1_ref.err
tail -8 gpurun_out/exp12.log; python - <<'PY'
import json
d=json.loads(open('gpurun_out/bench_r1_l.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], d['e2e'], d['roofline'])
print(open('gpurun_out/bench_r1_ref.json').read()[:600])
PY
