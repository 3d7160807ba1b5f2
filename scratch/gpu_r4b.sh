#!/bin/bash
# skip predicate ahead of the second-pass correction (the correction GEMM of a group is skipped with the rest)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x > gpurun_out/r4b_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r4b_pytest.log; tail -5 gpurun_out/r4b_pytest.log
echo "--- predicate first (default)"; timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
echo "--- previous commit"; TNB_LIB_PATH=scratch/exp/libtnb_prev.so timeout 200 python scratch/gemm_shapes.py 2>&1 | tail -2
timeout 300 compute-sanitizer --tool memcheck python scratch/sanity_small.py > gpurun_out/memcheck_r02h.log 2>&1; tail -3 gpurun_out/memcheck_r02h.log
timeout 600 python bench.py --no-cpu-baseline --no-batched > gpurun_out/bench_r4b.json 2> gpurun_out/r4b_bench_err.log; cut -c1-260 gpurun_out/bench_r4b.json; tail -3 gpurun_out/r4b_bench_err.log
TNB_LIB_PATH=scratch/exp/libtnb_prev.so timeout 600 python bench.py --no-cpu-baseline --no-batched > gpurun_out/bench_r4b_prev.json 2> gpurun_out/r4b_bench2_err.log; cut -c1-260 gpurun_out/bench_r4b_prev.json
