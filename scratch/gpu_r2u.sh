#!/bin/bash
# blocked DMMA Cholesky + inverse (chol_inv_kernel): tests, QR / SVD timing, bench, QR timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2u_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2u_pytest.log; tail -8 gpurun_out/r2u_pytest.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2u_gemm.log 2>&1; tail -2 gpurun_out/r2u_gemm.log
timeout 600 python bench.py > gpurun_out/bench_r2u.json 2> gpurun_out/r2u_bench_err.log; cut -c1-260 gpurun_out/bench_r2u.json; tail -3 gpurun_out/r2u_bench_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_qr_r02c.csv python scratch/one_op.py qr > gpurun_out/ncu_qr.log 2>&1
grep -c chol_inv gpurun_out/launches_qr_r02c.csv; grep chol_inv gpurun_out/launches_qr_r02c.csv | head -3 | cut -c1-300
