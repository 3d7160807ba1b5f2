#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/qr_launches.csv python scratch/one_op.py qr > gpurun_out/exp6.log 2>&1
tail -3 gpurun_out/exp6.log
