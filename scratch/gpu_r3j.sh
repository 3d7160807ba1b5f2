#!/bin/bash
# 8-GPU run of the bench: replicated cfg 3 sweeps (weak scaling) and the sharded batched cfg 4 record
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_r3j_8gpu.json 2> gpurun_out/r3j_err.log
cut -c1-300 gpurun_out/bench_r3j_8gpu.json; tail -3 gpurun_out/r3j_err.log
