#!/bin/bash
# second QR (R^H = Q2 R2, Jacobi on R2^H) for tall / square projection SVDs: tests, A/B timing, cfg 5, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3z_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3z_pytest.log; tail -6 gpurun_out/r3z_pytest.log
echo "--- second QR (default)"; timeout 300 python scratch/svd_tall.py 2>&1 | tee gpurun_out/svd_tall_r02f.txt
echo "--- Jacobi on the columns of R (previous)"; TNB_LIB_PATH=scratch/exp/libtnb_noqr2.so timeout 300 python scratch/svd_tall.py 2>&1 | tee gpurun_out/svd_tall_r02f_noqr2.txt
timeout 300 python scratch/cfg5_profile.py > gpurun_out/cfg5_profile_r02f.json 2> gpurun_out/r3z_cfg5_err.log; cut -c1-1500 gpurun_out/cfg5_profile_r02f.json; tail -3 gpurun_out/r3z_cfg5_err.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_r3z.json 2> gpurun_out/r3z_bench_err.log; cut -c1-260 gpurun_out/bench_r3z.json; tail -3 gpurun_out/r3z_bench_err.log
