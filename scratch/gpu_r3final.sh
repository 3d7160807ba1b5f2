#!/bin/bash
# final build of the round: smoke, tests, bench line of record, reference arm skipped (profiles/bench_r02b_reference_arm.json)
mkdir -p gpurun_out
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r3final_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/r3final_pytest.log; tail -4 gpurun_out/r3final_pytest.log
timeout 600 python bench.py > gpurun_out/bench_r3final.json 2> gpurun_out/r3final_bench_err.log; cut -c1-260 gpurun_out/bench_r3final.json; tail -3 gpurun_out/r3final_bench_err.log
