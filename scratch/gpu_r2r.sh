#!/bin/bash
# default library (conjugation fold, launch_k) + experiment variants: rotconv, QR PDL, QR group projection
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2r_pytest.log
tail -6 gpurun_out/r2r_pytest.log
TNB_LIB_PATH=scratch/exp/libtnb_nopad.so timeout 300 python scratch/jac_time.py > gpurun_out/r2r_jac_nopad.log 2>&1
timeout 300 python scratch/jac_time.py > gpurun_out/r2r_jac.log 2>&1
TNB_LIB_PATH=scratch/exp/libtnb_rotconv.so timeout 300 python scratch/jac_time.py > gpurun_out/r2r_jac_rotconv.log 2>&1
cat gpurun_out/r2r_jac_nopad.log gpurun_out/r2r_jac.log gpurun_out/r2r_jac_rotconv.log
TNB_LIB_PATH=scratch/exp/libtnb_rotconv.so timeout 600 python -m pytest tests -m gpu -q -k "svd or fullsize or batched" > gpurun_out/r2r_pytest_rotconv.log 2>&1; tail -3 gpurun_out/r2r_pytest_rotconv.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2r_gemm.log 2>&1; cat gpurun_out/r2r_gemm.log
for v in qrpdl qrgrp qrboth; do
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2r_gemm_$v.log 2>&1; tail -2 gpurun_out/r2r_gemm_$v.log
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 600 python -m pytest tests -m gpu -q -k "qr or svd or fullsize" > gpurun_out/r2r_pytest_$v.log 2>&1; tail -3 gpurun_out/r2r_pytest_$v.log
done
TNB_LIB_PATH=scratch/exp/libtnb_qrboth.so timeout 600 python bench.py --no-batched --no-cpu-baseline > gpurun_out/bench_r2r_qrboth.json 2> gpurun_out/r2r_bench_qrboth_err.log; cut -c1-200 gpurun_out/bench_r2r_qrboth.json
timeout 600 python bench.py > gpurun_out/bench_r2r.json 2> gpurun_out/r2r_bench_err.log
cut -c1-300 gpurun_out/bench_r2r.json; tail -3 gpurun_out/r2r_bench_err.log
