#!/bin/bash
# G <- G0 - C^H C of a first-pass block as one kernel (gram_correct_kernel) instead of split-K GEMM + reduce
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r4a_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r4a_pytest.log; tail -5 gpurun_out/r4a_pytest.log
echo "--- gram_correct_kernel (default)"; timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
echo "--- split-K GEMM + reduce (previous)"; TNB_LIB_PATH=scratch/exp/libtnb_gemmgcorr.so timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r4a.json 2> gpurun_out/r4a_bench_err.log; cut -c1-260 gpurun_out/bench_r4a.json; tail -3 gpurun_out/r4a_bench_err.log
TNB_LIB_PATH=scratch/exp/libtnb_gemmgcorr.so timeout 600 python bench.py --no-cpu-baseline --no-batched > gpurun_out/bench_r4a_prev.json 2> gpurun_out/r4a_bench2_err.log; cut -c1-260 gpurun_out/bench_r4a_prev.json
