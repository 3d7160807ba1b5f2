#!/bin/bash
# validation of the promoted kernels (TMA GEMM, mps_mpo, permute, one-CTA-per-SM padding) + timelines + ncu captures
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2q_pytest.log
tail -6 gpurun_out/r2q_pytest.log
TNB_LIB_PATH=scratch/exp/libtnb_nopad.so timeout 300 python scratch/jac_time.py > gpurun_out/r2q_jac_nopad.log 2>&1
timeout 300 python scratch/jac_time.py > gpurun_out/r2q_jac.log 2>&1
TNB_LIB_PATH=scratch/exp/libtnb_rotconv.so timeout 300 python scratch/jac_time.py > gpurun_out/r2q_jac_rotconv.log 2>&1
cat gpurun_out/r2q_jac_nopad.log gpurun_out/r2q_jac.log gpurun_out/r2q_jac_rotconv.log
TNB_LIB_PATH=scratch/exp/libtnb_rotconv.so timeout 600 python -m pytest tests -m gpu -q -k "svd or fullsize or batched" > gpurun_out/r2q_pytest_rotconv.log 2>&1; tail -3 gpurun_out/r2q_pytest_rotconv.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2q_gemm.log 2>&1; cat gpurun_out/r2q_gemm.log
for v in qrpdl qrgrp qrboth; do
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2q_gemm_$v.log 2>&1; tail -2 gpurun_out/r2q_gemm_$v.log
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 600 python -m pytest tests -m gpu -q -k "qr or svd or fullsize" > gpurun_out/r2q_pytest_$v.log 2>&1; tail -3 gpurun_out/r2q_pytest_$v.log
done
TNB_LIB_PATH=scratch/exp/libtnb_qrboth.so timeout 600 python bench.py --no-batched --no-cpu-baseline > gpurun_out/bench_r2q_qrboth.json 2> gpurun_out/r2q_bench_qrboth_err.log; cut -c1-200 gpurun_out/bench_r2q_qrboth.json
timeout 600 python bench.py > gpurun_out/bench_r2q.json 2> gpurun_out/r2q_bench_err.log
cut -c1-300 gpurun_out/bench_r2q.json; tail -3 gpurun_out/r2q_bench_err.log
# launch timelines (serialised, cold): one QR 3072x1536, one projection SVD 1024x1536
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_qr_r02.csv python scratch/one_op.py qr > gpurun_out/ncu_qr.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 1200 --csv --log-file gpurun_out/launches_svd_r02.csv python scratch/one_op.py svd > gpurun_out/ncu_svd.log 2>&1
# ncu --set full: TMA GEMM (absorb shapes), skinny split-K GEMMs inside the QR, Jacobi round, HBM-bound kernels
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm_tma -c 3 -f -o gpurun_out/prof_gemm_tma_r02 python scratch/one_op.py absorb > gpurun_out/ncu_g2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 400 -c 12 -f -o gpurun_out/prof_gemm_qr_r02 python scratch/one_op.py qr > gpurun_out/ncu_g3.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:jacobi_round -s 300 -c 1 -f -o gpurun_out/prof_jacobi_r02c python scratch/one_op.py svd > gpurun_out/ncu_j2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"permute|mps_mpo" -c 8 -f -o gpurun_out/prof_hbm_r02b python scratch/hbm_ops.py > gpurun_out/ncu_h2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/launches_r02b.csv python bench.py --sites 14 --steps 1 --warmup 1 --no-cpu-baseline --no-batched > gpurun_out/ncu_bench2.log 2>&1
for f in prof_gemm_tma_r02 prof_gemm_qr_r02 prof_jacobi_r02c prof_hbm_r02b; do
  ncu -i gpurun_out/$f.ncu-rep --page raw --csv > gpurun_out/$f.raw.csv 2>/dev/null
done
rm -f gpurun_out/*.ncu-rep   # the raw csv pages are what is read here; the reports are tens of MB each
ls -la gpurun_out | tail -30
