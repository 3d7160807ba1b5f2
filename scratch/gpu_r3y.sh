#!/bin/bash
# null-column floor of the projection SVD: tests, cfg 5 by kernel class with the floor at sqrt(k) eps (default) and at eps
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3y_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3y_pytest.log; tail -6 gpurun_out/r3y_pytest.log
timeout 300 python scratch/cfg5_profile.py > gpurun_out/cfg5_profile_r02e.json 2> gpurun_out/r3y_cfg5_err.log; cut -c1-1500 gpurun_out/cfg5_profile_r02e.json; tail -3 gpurun_out/r3y_cfg5_err.log
TNB_LIB_PATH=scratch/exp/libtnb_eps.so timeout 300 python scratch/cfg5_profile.py > gpurun_out/cfg5_profile_r02e_eps.json 2> gpurun_out/r3y_cfg5b_err.log; cut -c1-1500 gpurun_out/cfg5_profile_r02e_eps.json; tail -3 gpurun_out/r3y_cfg5b_err.log
