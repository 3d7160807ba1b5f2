#!/bin/bash
# restored-state check: GPU test suite + default bench line
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2n_pytest.log
timeout 600 python bench.py > gpurun_out/bench_r2n.json 2> gpurun_out/r2n_bench_err.log
tail -4 gpurun_out/r2n_pytest.log; cut -c1-600 gpurun_out/bench_r2n.json; tail -3 gpurun_out/r2n_bench_err.log
