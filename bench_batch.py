#!/usr/bin/env python
"""Secondary benchmark: the batched path (BASELINE.json config 4).

    python bench_batch.py [--gpus N] [--networks-per-gpu M] [--streams S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_batch.py --gpus N ...

Independent random MPS networks (N=64, d=4, chi=128, float64; per network <a|b>, |a| and a.svd_compress(chi=64))
are sharded over the ranks (one process per GPU, contiguous slices, inputs generated on the owning GPU from the
counter-based generator of tncontract_b200.batch); inside a rank the networks are spread over S host threads, each on
its own CUDA stream, because one chi=128 network cannot fill a B200.  No collective sits on the data path; the
per-network records are gathered at the end (one small all_gather).  Rank 0 prints ONE JSON line: networks/s over all
GPUs (device-timed with CUDA events, max over ranks, weak scaling: M networks per GPU).  bench.py stays the headline
benchmark (config 3); this script documents the batched row of the scope table."""
import argparse
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--networks-per-gpu", type=int, default=64)
    ap.add_argument("--streams", type=int, default=8)
    ap.add_argument("--sites", type=int, default=64)
    ap.add_argument("--d", type=int, default=4)
    ap.add_argument("--chi", type=int, default=128)
    ap.add_argument("--keep", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8, help="untimed networks per rank before the timed batch")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        if os.environ.get("NCCL_DEBUG", "").upper() in ("VERSION", "WARN"):
            del os.environ["NCCL_DEBUG"]  # both levels print the version banner on stdout
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    import tncontract_b200 as tn
    from tncontract_b200 import batch

    def unit(i):
        return batch.overlap_norm_compress(3, i, args.sites, args.d, args.chi, args.keep)

    total = args.networks_per_gpu * world
    # warm-up on networks outside the timed range (same shapes: allocator, kernel attributes, cluster queries)
    batch.run_sharded(args.warmup * world, lambda i: unit(total + i), rank, world, streams=args.streams)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = tn.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    recs = batch.run_sharded(total, unit, rank, world, streams=args.streams)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = tn.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    allrecs = batch.gather_results(recs, world)
    if rank == 0:
        bonds = sorted({int(max(r[4:])) for r in allrecs})
        line = {"metric": "batched MPS overlap+norm+svd_compress networks/s (N=%d,d=%d,chi=%d->%d,float64)" %
                          (args.sites, args.d, args.chi, args.keep),
                "value": total / (t.item() * 1e-3), "unit": "networks/s", "n_gpus": world,
                "networks": total, "networks_per_gpu": args.networks_per_gpu, "streams_per_gpu": args.streams,
                "ms": t.item(), "higher_is_better": True, "scaling": "weak", "dtype": "f64", "data": "synthetic",
                "gpu_launches_rank0": launches, "records_gathered": len(allrecs), "max_bond_after": bonds,
                "config": {"workload": "cfg4: independent random MPS, contiguous shards, no data-path collective"}}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
