#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/final2.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/final2.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1_final.json 2> gpurun_out/bench_r1_final.err
timeout 300 python scratch/site_ops.py all 3 >> gpurun_out/final2.log 2>&1
TNB_JACOBI_FIXED_SWEEPS=5 timeout 120 python scratch/jac_phases.py >> gpurun_out/final2.log 2>&1
tail -22 gpurun_out/final2.log; cut -c1-260 gpurun_out/bench_r1_final.json
