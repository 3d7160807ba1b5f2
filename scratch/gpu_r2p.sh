#!/bin/bash
# ncu captures of the round-2 (second session) kernels: split-tournament Jacobi round, TMA-fed GEMM, HBM-bound kernels,
# launch list of a short bench; .ncu-rep files stay in gpurun_out/, summaries go to profiles/
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 600 -c 2 -f -o gpurun_out/prof_jacobi_r02c python scratch/one_op.py svd > gpurun_out/ncu_j2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma -s 2 -c 4 -f -o gpurun_out/prof_gemm_tma_r02 python scratch/one_op.py absorb > gpurun_out/ncu_g2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"permute|mps_mpo" -c 8 -f -o gpurun_out/prof_hbm_r02b python scratch/hbm_ops.py > gpurun_out/ncu_h2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --sites 14 --steps 1 --warmup 1 --no-cpu-baseline --no-batched > gpurun_out/ncu_bench2.log 2>&1
tail -3 gpurun_out/ncu_j2.log gpurun_out/ncu_g2.log gpurun_out/ncu_h2.log gpurun_out/ncu_bench2.log; wc -l gpurun_out/launches_r02b.csv
for f in prof_jacobi_r02c prof_gemm_tma_r02 prof_hbm_r02b; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
ls -la gpurun_out | head -40
