#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -5 > gpurun_out/final1.log
python -c "import __graft_entry__ as g; g.smoke()" >> gpurun_out/final1.log 2>&1
timeout 900 python bench.py > gpurun_out/bench_r1_o.json 2> gpurun_out/bench_r1_o.err
timeout 200 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r01h python scratch/one_op.py svd > gpurun_out/ncu_final_j.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 150 -c 40 -f -o gpurun_out/prof_gemm_r01d python scratch/one_op.py qr > gpurun_out/ncu_final_g.log 2>&1
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -c 5000 --csv --log-file gpurun_out/launches_r01c.csv python bench.py --sites 14 --steps 1 --warmup 1 --no-cpu-baseline > gpurun_out/ncu_bench_c.log 2>&1
tail -6 gpurun_out/final1.log; cut -c1-300 gpurun_out/bench_r1_o.json
