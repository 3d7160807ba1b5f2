#!/bin/bash
mkdir -p gpurun_out; rm -f gpurun_out/r2m_batch.log
python bench_batch.py --networks-per-gpu 296 >> gpurun_out/r2m_batch.log 2> gpurun_out/r2m_err.log
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500+n)) bench_batch.py --gpus $n --networks-per-gpu 296 >> gpurun_out/r2m_batch.log 2>> gpurun_out/r2m_err.log
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 8 --steps 2 --warmup 3 > gpurun_out/bench_r2m_8gpu.json 2>> gpurun_out/r2m_err.log
cut -c90-260 gpurun_out/r2m_batch.log; cut -c1-400 gpurun_out/bench_r2m_8gpu.json; tail -5 gpurun_out/r2m_err.log
