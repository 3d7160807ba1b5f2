#!/bin/bash
# third session, final build: tests, memcheck of the QR skip path (skipped and corrected groups), cfg 5 by kernel class, bench, launch list
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3x_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3x_pytest.log; tail -4 gpurun_out/r3x_pytest.log
timeout 400 compute-sanitizer --tool memcheck python scratch/sanity_small.py > gpurun_out/memcheck_r02d.log 2>&1; tail -12 gpurun_out/memcheck_r02d.log
timeout 300 python scratch/cfg5_profile.py > gpurun_out/cfg5_profile_r02d.json 2> gpurun_out/r3x_cfg5_err.log; cut -c1-1500 gpurun_out/cfg5_profile_r02d.json; tail -3 gpurun_out/r3x_cfg5_err.log
timeout 600 python bench.py > gpurun_out/bench_r3x.json 2> gpurun_out/r3x_bench_err.log; cut -c1-260 gpurun_out/bench_r3x.json; tail -3 gpurun_out/r3x_bench_err.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 200000 --csv --log-file gpurun_out/launches_r02d.csv python bench.py --sites 14 --steps 1 --warmup 1 --no-cpu-baseline --no-batched > gpurun_out/ncu_bench_r3x.log 2>&1; tail -2 gpurun_out/ncu_bench_r3x.log | cut -c1-200
