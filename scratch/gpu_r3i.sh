#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cabi_gpu.py -m gpu -q -k "gemm" > gpurun_out/r3i_pytest.log 2>&1; tail -5 gpurun_out/r3i_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r3i_smoke.log 2>&1; tail -3 gpurun_out/r3i_smoke.log
