#!/bin/bash
# NEXT kernels (csrc_next: split tournament, TMA GEMM, new mps_mpo / permute) against their A/B variants, then the GPU
# test suite and the bench line with the next library; finally the bench line of the committed library
mkdir -p gpurun_out
N=scratch/exp/libtnb_next.so
TNB_LIB_PATH=scratch/exp/libtnb_flat2.so timeout 300 python scratch/jac_time.py > gpurun_out/r2o_jac_flat.log 2>&1
TNB_LIB_PATH=$N timeout 300 python scratch/jac_time.py > gpurun_out/r2o_jac_split.log 2>&1
cat gpurun_out/r2o_jac_flat.log gpurun_out/r2o_jac_split.log
TNB_LIB_PATH=scratch/exp/libtnb_notma.so timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2o_gemm_notma.log 2>&1
TNB_LIB_PATH=$N timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2o_gemm_tma.log 2>&1
TNB_LIB_PATH=scratch/exp/libtnb_sk1.so timeout 200 python scratch/gemm_shapes.py > gpurun_out/r2o_gemm_sk1.log 2>&1
cat gpurun_out/r2o_gemm_notma.log gpurun_out/r2o_gemm_tma.log gpurun_out/r2o_gemm_sk1.log
TNB_LIB_PATH=scratch/exp/libtnb_stamps.so timeout 120 python scratch/jac_stamps.py > gpurun_out/r2o_stamps.log 2>&1; tail -8 gpurun_out/r2o_stamps.log
timeout 200 python scratch/hbm_ops.py time > gpurun_out/r2o_hbm_time_old.log 2>&1
TNB_LIB_PATH=$N timeout 200 python scratch/hbm_ops.py time > gpurun_out/r2o_hbm_time.log 2>&1; cat gpurun_out/r2o_hbm_time_old.log gpurun_out/r2o_hbm_time.log
TNB_LIB_PATH=$N timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2o_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r2o_pytest.log
tail -25 gpurun_out/r2o_pytest.log
TNB_LIB_PATH=$N timeout 600 python bench.py > gpurun_out/bench_r2o.json 2> gpurun_out/r2o_bench_err.log
cut -c1-300 gpurun_out/bench_r2o.json; tail -3 gpurun_out/r2o_bench_err.log
timeout 600 python bench.py --no-batched > gpurun_out/bench_r2o_committed.json 2> gpurun_out/r2o_bench2_err.log
cut -c1-300 gpurun_out/bench_r2o_committed.json
