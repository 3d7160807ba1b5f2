#!/bin/bash
# sanitizer runs on the final build of the session
mkdir -p gpurun_out
timeout 900 compute-sanitizer --tool memcheck python scratch/sanity_small.py > gpurun_out/memcheck_r02c.log 2>&1; tail -3 gpurun_out/memcheck_r02c.log
timeout 1500 compute-sanitizer --tool racecheck python scratch/sanity_small.py > gpurun_out/racecheck_r02c.log 2>&1; tail -3 gpurun_out/racecheck_r02c.log
