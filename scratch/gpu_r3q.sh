#!/bin/bash
# float64 TMA-fed GEMM (experiment build): direct tests, timing against the cp.async kernel, whole GPU suite
mkdir -p gpurun_out
TNB_LIB_PATH=scratch/exp/libtnb_realtma.so timeout 600 python -m pytest tests/test_cabi_gpu.py -m gpu -q -k "gemm or tensordot" > gpurun_out/r3q_pytest_gemm.log 2>&1; tail -4 gpurun_out/r3q_pytest_gemm.log
timeout 300 python scratch/gemm_shapes_real.py > gpurun_out/r3q_real_default.log 2>&1; cat gpurun_out/r3q_real_default.log
TNB_LIB_PATH=scratch/exp/libtnb_realtma.so timeout 300 python scratch/gemm_shapes_real.py > gpurun_out/r3q_real_tma.log 2>&1; cat gpurun_out/r3q_real_tma.log
TNB_LIB_PATH=scratch/exp/libtnb_realtma.so timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3q_pytest.log 2>&1; tail -4 gpurun_out/r3q_pytest.log
