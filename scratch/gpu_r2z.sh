#!/bin/bash
# multi-GPU check of the bench (replicated cfg 3 sweeps + sharded batched cfg 4 record) at N = 2
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 2 --warmup 3 > gpurun_out/bench_r2z_2gpu.json 2> gpurun_out/r2z_err.log
cut -c1-400 gpurun_out/bench_r2z_2gpu.json; tail -3 gpurun_out/r2z_err.log
