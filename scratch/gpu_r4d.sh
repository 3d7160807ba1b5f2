#!/bin/bash
# look-ahead of the first-pass S product on a third stream
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r4d_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r4d_pytest.log; tail -5 gpurun_out/r4d_pytest.log
echo "--- look-ahead (default)"; timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
echo "--- no look-ahead"; TNB_LIB_PATH=scratch/exp/libtnb_nolook.so timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
timeout 600 python bench.py --no-cpu-baseline --no-batched > gpurun_out/bench_r4d.json 2> gpurun_out/r4d_bench_err.log; cut -c1-260 gpurun_out/bench_r4d.json; tail -3 gpurun_out/r4d_bench_err.log
TNB_LIB_PATH=scratch/exp/libtnb_nolook.so timeout 600 python bench.py --no-cpu-baseline --no-batched > gpurun_out/bench_r4d_prev.json 2> gpurun_out/r4d_bench2_err.log; cut -c1-260 gpurun_out/bench_r4d_prev.json
