#!/bin/bash
# in-block steps folded into cross round 0: Jacobi timing, tests, bench
mkdir -p gpurun_out
timeout 300 python scratch/jac_time.py > gpurun_out/r3n_jac.log 2>&1; cat gpurun_out/r3n_jac.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3n_pytest.log; tail -4 gpurun_out/r3n_pytest.log
timeout 600 python bench.py > gpurun_out/bench_r3n.json 2> gpurun_out/r3n_bench_err.log; cut -c1-260 gpurun_out/bench_r3n.json; tail -3 gpurun_out/r3n_bench_err.log
