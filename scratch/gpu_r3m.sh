#!/bin/bash
# final validation of HEAD: whole GPU suite, shapes, bench
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3m_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3m_pytest.log; tail -4 gpurun_out/r3m_pytest.log
timeout 200 python scratch/gemm_shapes.py > gpurun_out/r3m_gemm.log 2>&1; tail -3 gpurun_out/r3m_gemm.log
timeout 600 python bench.py > gpurun_out/bench_r3m.json 2> gpurun_out/r3m_bench_err.log; cut -c1-260 gpurun_out/bench_r3m.json; tail -3 gpurun_out/r3m_bench_err.log
timeout 900 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/bench_r3m_reference.json 2> gpurun_out/r3m_ref_err.log; cut -c1-400 gpurun_out/bench_r3m_reference.json
