#!/bin/bash
mkdir -p gpurun_out
timeout 600 python scratch/batch_prof.py 148 > gpurun_out/r3f_batch_prof.log 2>&1; cat gpurun_out/r3f_batch_prof.log | tail -12
