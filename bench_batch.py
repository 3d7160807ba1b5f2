#!/usr/bin/env python
"""Secondary benchmark: the batched path (BASELINE.json config 4).

    python bench_batch.py [--gpus N] [--networks-per-gpu M] [--streams S]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 bench_batch.py --gpus N ...

Independent random MPS networks (N=64, d=4, chi=128, float64; per network <a|b>, |a| and a.svd_compress(chi=64))
are sharded over the ranks (one process per GPU, contiguous slices, inputs generated on the owning GPU from the
counter-based generator of tncontract_b200.batch); inside a rank the networks are spread over S host threads, each on
its own CUDA stream (--mode streams, the round-1 path) or -- the default, --mode batched -- processed in groups with
ONE launch per site for the whole group (tncontract_b200.batched: strided-batched GEMMs and the batched projection
SVD, grid row = network).  No collective sits on the data path; the per-network records are gathered at the end
(one small all_gather).  Rank 0 prints ONE JSON line: networks/s over all
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


def measure(rank, world, networks_per_gpu=148, batch_size=148, sites=64, d=4, chi=128, keep=64, mode="batched",
            streams=8, warmup=8):
    """Time the cfg 4 unit of work over this rank's shard (process group, if any, already initialised; the current
    CUDA device is the rank's).  Returns the record on rank 0 (None elsewhere).  Device-timed with CUDA events,
    max over ranks; weak scaling: networks_per_gpu networks on every GPU."""
    import torch
    import torch.distributed as dist
    import tncontract_b200 as tn
    from tncontract_b200 import batch, batched

    def unit(i):
        return batch.overlap_norm_compress(3, i, sites, d, chi, keep)

    fallbacks = [0]

    def group(idx):
        """one batched group; a ragged group (networks keeping different bond dimensions) goes network by network"""
        try:
            return batched.overlap_norm_compress_batched(3, idx, sites, d, chi, keep)[0]
        except batched.RaggedBatchError:
            fallbacks[0] += 1
            return [unit(i) for i in idx]

    def run(n_items, first):
        if mode == "streams":
            return batch.run_sharded(n_items, lambda i: unit(first + i), rank, world, streams=streams)
        lo, hi = batch.shard_range(n_items, rank, world)
        out = []
        for g0 in range(lo, hi, batch_size):
            out.extend(group([first + i for i in range(g0, min(hi, g0 + batch_size))]))
        return out

    total = networks_per_gpu * world
    # warm-up on networks outside the timed range (same shapes: allocator, kernel attributes, cluster queries)
    run((warmup if mode == "streams" else min(batch_size, networks_per_gpu)) * world, total)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    l0 = tn.launch_count()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    recs = run(total, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    launches = tn.launch_count() - l0
    t = torch.tensor([ms], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    allrecs = batch.gather_results(recs, world)
    if rank != 0:
        return None
    bonds = sorted({int(max(r[4:])) for r in allrecs})
    return {"metric": "batched MPS overlap+norm+svd_compress networks/s (N=%d,d=%d,chi=%d->%d,float64)" %
                      (sites, d, chi, keep),
            "value": total / (t.item() * 1e-3), "unit": "networks/s", "n_gpus": world,
            "networks": total, "networks_per_gpu": networks_per_gpu, "mode": mode,
            "batch": batch_size if mode == "batched" else None,
            "streams_per_gpu": streams if mode == "streams" else None,
            "ragged_group_fallbacks_rank0": fallbacks[0],
            "launches_per_network_rank0": launches / max(1, networks_per_gpu),
            "ms": t.item(), "higher_is_better": True, "scaling": "weak", "dtype": "f64", "data": "synthetic",
            "gpu_launches_rank0": launches, "records_gathered": len(allrecs), "max_bond_after": bonds,
            "config": {"workload": "cfg4: independent random MPS, contiguous shards, no data-path collective; "
                                   "one NCCL all_gather of the per-network records at the end"}}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--networks-per-gpu", type=int, default=296)
    ap.add_argument("--streams", type=int, default=8, help="--mode streams: host threads x CUDA streams per GPU")
    ap.add_argument("--mode", default="batched", choices=["batched", "streams"],
                    help="batched: one launch per site for a whole group of networks (tncontract_b200.batched); "
                         "streams: the round-1 path, one network per host thread / CUDA stream")
    ap.add_argument("--batch", type=int, default=148, help="--mode batched: networks per batched group "
                    "(4 block pairs per network and Jacobi round: 37 networks fill the 148 SMs once)")
    ap.add_argument("--sites", type=int, default=64)
    ap.add_argument("--d", type=int, default=4)
    ap.add_argument("--chi", type=int, default=128)
    ap.add_argument("--keep", type=int, default=64)
    ap.add_argument("--warmup", type=int, default=8, help="--mode streams: untimed networks per rank before the timed batch")
    args = ap.parse_args()

    import torch
    import torch.distributed as dist
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    line = measure(rank, world, args.networks_per_gpu, args.batch, args.sites, args.d, args.chi, args.keep, args.mode,
                   args.streams, args.warmup)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
