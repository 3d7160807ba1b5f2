#!/bin/bash
mkdir -p gpurun_out
timeout 900 python bench.py > gpurun_out/bench_r2g.json 2> gpurun_out/bench_r2g.err
cut -c1-3000 gpurun_out/bench_r2g.json; tail -5 gpurun_out/bench_r2g.err
