#!/bin/bash
# one-wave split-K planner + fused in-place panel scale: tests, QR-shape GEMMs, QR / SVD timing, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r2t_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2t_pytest.log; tail -5 gpurun_out/r2t_pytest.log
timeout 200 python scratch/gemm_qr_shapes.py > gpurun_out/r2t_shapes.log 2>&1; cat gpurun_out/r2t_shapes.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2t_gemm.log 2>&1; tail -4 gpurun_out/r2t_gemm.log
timeout 600 python bench.py > gpurun_out/bench_r2t.json 2> gpurun_out/r2t_bench_err.log; cut -c1-260 gpurun_out/bench_r2t.json; tail -3 gpurun_out/r2t_bench_err.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/launches_qr_r02b.csv python scratch/one_op.py qr > gpurun_out/ncu_qr.log 2>&1
