#!/bin/bash
# flakiness check: the whole GPU suite three times, once more with blocking launches; smoke
mkdir -p gpurun_out
for i in 1 2 3; do timeout 900 python -m pytest tests -m gpu -q 2>&1 | tail -1; done
CUDA_LAUNCH_BLOCKING=1 timeout 1200 python -m pytest tests -m gpu -q 2>&1 | tail -1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
