#!/bin/bash
mkdir -p gpurun_out
timeout 200 python scratch/hbm_ops.py time > gpurun_out/hbm_time_r02.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"permute|mps_mpo|scale_inplace|norm2" -c 14 -f -o gpurun_out/prof_hbm_r02 python scratch/hbm_ops.py > gpurun_out/ncu_hbm.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r02b python scratch/one_op.py svd > gpurun_out/ncu_j.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:gemm_kernel -s 150 -c 20 -f -o gpurun_out/prof_gemm_r02 python scratch/one_op.py qr > gpurun_out/ncu_g.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/launches_r02.csv python bench.py --sites 14 --steps 1 --warmup 1 --no-cpu-baseline --no-batched > gpurun_out/ncu_bench.log 2>&1
timeout 600 compute-sanitizer --tool racecheck python scratch/sanity_small.py > gpurun_out/racecheck_r02.log 2>&1
cat gpurun_out/hbm_time_r02.log; tail -3 gpurun_out/ncu_hbm.log gpurun_out/ncu_j.log gpurun_out/ncu_g.log gpurun_out/ncu_bench.log; tail -5 gpurun_out/racecheck_r02.log; wc -l gpurun_out/launches_r02.csv
