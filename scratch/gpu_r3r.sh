#!/bin/bash
# float64 TMA GEMM as the default: whole GPU suite, cfg 5 at its stated size, batched bench, main bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3r_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3r_pytest.log; tail -4 gpurun_out/r3r_pytest.log
timeout 600 python scratch/cfg5_full.py > gpurun_out/cfg5_full_r03.json 2> gpurun_out/r3r_cfg5_err.log; cut -c1-200 gpurun_out/cfg5_full_r03.json; tail -2 gpurun_out/r3r_cfg5_err.log
timeout 600 python bench.py > gpurun_out/bench_r3r.json 2> gpurun_out/r3r_bench_err.log; cut -c1-260 gpurun_out/bench_r3r.json; tail -3 gpurun_out/r3r_bench_err.log
