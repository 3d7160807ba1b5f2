#!/bin/bash
mkdir -p gpurun_out
{
timeout 900 python -m pytest tests/test_cabi_gpu.py -x -q -m gpu -k "qr" 2>&1 | tail -25
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -8
timeout 300 python scratch/site_ops.py qr 3
timeout 300 python scratch/site_ops.py qrprof
TNB_QR_BCGS=0 timeout 300 python scratch/site_ops.py qr 3
} > gpurun_out/exp5.log 2>&1
tail -70 gpurun_out/exp5.log
