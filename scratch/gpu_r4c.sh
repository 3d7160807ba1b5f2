#!/bin/bash
# final build: launch list of one QR 3072 x 1536 (per-kernel times, serialised), ncu --set full of the two new QR kernels
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"gram_correct|skip_predicate" -s 12 -c 12 -f -o gpurun_out/prof_qrnew_r02h python scratch/one_op.py qr > gpurun_out/ncu_qrnew.log 2>&1; tail -2 gpurun_out/ncu_qrnew.log
ncu -i gpurun_out/prof_qrnew_r02h.ncu-rep --page raw --csv > gpurun_out/prof_qrnew_r02h.raw.csv 2>/dev/null; wc -c gpurun_out/prof_qrnew_r02h.raw.csv
