#!/bin/bash
# GEMM planner variants at the QR's shapes (tnb_gemm_ws micro-benchmark + whole QR), new split-K test
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cabi_gpu.py -m gpu -q -k "gemm" > gpurun_out/r2s_pytest.log 2>&1; tail -3 gpurun_out/r2s_pytest.log
timeout 200 python scratch/gemm_qr_shapes.py > gpurun_out/r2s_shapes_default.log 2>&1; cat gpurun_out/r2s_shapes_default.log
for v in sgs sk2 sk3 sk6; do
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 200 python scratch/gemm_qr_shapes.py > gpurun_out/r2s_shapes_$v.log 2>&1; cat gpurun_out/r2s_shapes_$v.log
  TNB_LIB_PATH=scratch/exp/libtnb_$v.so timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
done
