#!/bin/bash
# R update on the side stream: tests, QR timing, bench, QR timeline
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2w_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2w_pytest.log; tail -8 gpurun_out/r2w_pytest.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2w_gemm.log 2>&1; tail -2 gpurun_out/r2w_gemm.log
timeout 600 python bench.py > gpurun_out/bench_r2w.json 2> gpurun_out/r2w_bench_err.log; cut -c1-260 gpurun_out/bench_r2v.json; tail -3 gpurun_out/r2w_bench_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_qr_r02e.csv python scratch/one_op.py qr > gpurun_out/ncu_qr.log 2>&1
grep panel_scale gpurun_out/launches_qr_r02e.csv | head -2 | cut -c150-330
