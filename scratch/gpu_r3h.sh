#!/bin/bash
# final captures of the session: launch list of a short bench, ncu of the Jacobi round / chol / panel_scale, sanitizer runs
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/launches_r02c.csv python bench.py --sites 14 --steps 1 --warmup 1 --no-cpu-baseline --no-batched > gpurun_out/ncu_bench3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r02e python scratch/one_op.py svd > gpurun_out/ncu_j4.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"chol_inv|panel_scale" -s 20 -c 2 -f -o gpurun_out/prof_qrsmall_r02 python scratch/one_op.py qr > gpurun_out/ncu_q4.log 2>&1
for f in prof_jacobi_r02e prof_qrsmall_r02; do ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null; done
rm -f gpurun_out/*.ncu-rep
timeout 900 compute-sanitizer --tool memcheck python scratch/sanity_small.py > gpurun_out/memcheck_r02b.log 2>&1; tail -4 gpurun_out/memcheck_r02b.log
timeout 1500 compute-sanitizer --tool racecheck python scratch/sanity_small.py > gpurun_out/racecheck_r02b.log 2>&1; tail -4 gpurun_out/racecheck_r02b.log
wc -l gpurun_out/launches_r02c.csv
