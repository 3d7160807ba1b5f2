"""Batched path: many independent networks, sharded over the GPUs of one box.

A sweep is sequential in the site index, so one network lives on one GPU; the
batch is what parallelises.  One process per GPU
(torch.distributed, NCCL on GPUs / gloo in the CPU tests); each rank owns a
contiguous slice of the network indices, creates its inputs on its own device
from a counter-based generator keyed by (seed, network, mps, site), runs the
per-network routine -- optionally on several CUDA streams driven by host
threads, because a chi~128 network cannot fill a B200 on its own -- and the
per-network scalars are gathered at the end.  No collective sits on the data
path.
"""
import concurrent.futures
import ctypes

import numpy as np

from . import _lib
from . import devarray as dv
from . import tensor as tsr
from .onedim import onedim_core as core

__all__ = ["shard_range", "host_uniform", "device_uniform", "random_mps", "network_key", "run_sharded",
           "gather_results", "overlap_norm_compress"]

_GOLDEN = np.uint64(0x9E3779B97F4A7C15)


def shard_range(n_items, rank, world):
    """Contiguous slice [lo, hi) of ``range(n_items)`` owned by ``rank``; sizes differ by at most one."""
    base, extra = divmod(int(n_items), int(world))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def network_key(seed, network, which, site):
    """64-bit stream key of one site tensor."""
    return (int(seed) * 0x100000001B3 + int(network) * 0x9E3779B1 + int(which) * 0x85EBCA6B + int(site) * 0xC2B2AE35 + 1) \
        & 0xFFFFFFFFFFFFFFFF


def host_uniform(n, key, offset=0):
    """The generator of tnb_fill_uniform, in NumPy (parity checks regenerate single networks with it)."""
    with np.errstate(over="ignore"):
        i = np.arange(n, dtype=np.uint64) + np.uint64(offset)
        z = np.uint64(key) + i * _GOLDEN
        z = (z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)
        z = (z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)
        z = z ^ (z >> np.uint64(31))
    return (z >> np.uint64(11)).astype(np.float64) * (1.0 / 9007199254740992.0)


def device_uniform(shape, key, offset=0):
    out = dv.DevArray.empty(shape, np.float64)
    n = out.size
    if n:
        _lib.check(_lib.load().tnb_fill_uniform(ctypes.c_void_p(out.ptr), n, ctypes.c_ulonglong(key),
                                                ctypes.c_ulonglong(offset), dv.stream_ptr()))
    return out


def random_mps(seed, network, which, nsites, physdim, bonddim, on_host=False):
    """U[0,1) MPS with bonds [1, chi, ..., chi, 1] (cfg 4 input); ``on_host`` builds the same state
    from host_uniform for parity checks."""
    bonds = [1] + [bonddim] * (nsites - 1) + [1]
    sites = []
    for i in range(nsites):
        shape = (physdim, bonds[i], bonds[i + 1])
        key = network_key(seed, network, which, i)
        if on_host:
            sites.append((host_uniform(int(np.prod(shape)), key).reshape(shape), ["phys", "left", "right"]))
        else:
            sites.append(tsr.Tensor._wrap(device_uniform(shape, key), ["phys", "left", "right"]))
    if on_host:
        return sites
    return core.MatrixProductState._adopt(sites, "left", "right", "phys")


def overlap_norm_compress(seed, network, nsites=64, physdim=4, bonddim=128, chi=64):
    """The cfg 4 unit of work: <a|b>, |a|, a.svd_compress(chi).  Returns a flat float64 record
    [network, overlap, norm, norm_after, bonds...]."""
    a = random_mps(seed, network, 0, nsites, physdim, bonddim)
    b = random_mps(seed, network, 1, nsites, physdim, bonddim)
    ov = core.inner_product_mps(a, b)
    nrm = a.norm()
    a.svd_compress(chi=chi)
    rec = [float(network), float(np.real(ov)), float(nrm), float(a.norm(canonical_form="right"))]
    return np.array(rec + [float(x) for x in a.bonddims()])


def run_sharded(n_networks, fn, rank=0, world=1, streams=1):
    """Run ``fn(network_index) -> 1-D float array`` over this rank's slice.  With ``streams`` > 1 the
    networks are spread over that many host threads, each on its own CUDA stream (the C ABI takes the
    stream per call; ctypes releases the GIL while a call runs)."""
    lo, hi = shard_range(n_networks, rank, world)
    idx = list(range(lo, hi))
    if streams <= 1 or len(idx) <= 1:
        return [fn(i) for i in idx]
    import torch
    cur = torch.cuda.current_stream()
    pool = [torch.cuda.Stream() for _ in range(streams)]
    for s in pool:
        s.wait_stream(cur)

    def work(k):
        out = []
        with torch.cuda.stream(pool[k]):
            for i in idx[k::streams]:
                out.append((i, fn(i)))
            pool[k].synchronize()
        return out

    with concurrent.futures.ThreadPoolExecutor(max_workers=streams) as ex:
        parts = list(ex.map(work, range(streams)))
    merged = dict(p for part in parts for p in part)
    for s in pool:
        cur.wait_stream(s)
    return [merged[i] for i in idx]


def gather_results(local, world=1):
    """All ranks' per-network records, in network order (rank slices are contiguous).  Uses the default
    process group (NCCL on GPUs: one all_gather of a few KB; gloo in the CPU tests)."""
    if world <= 1:
        return list(local)
    import torch
    import torch.distributed as dist
    width = max([len(r) for r in local] + [0])
    meta = torch.tensor([len(local), width], dtype=torch.int64)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend() == "nccl" else torch.device("cpu")
    meta = meta.to(dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta)
    rows = max(int(m[0]) for m in metas)
    width = max(int(m[1]) for m in metas)
    buf = torch.full((rows, width), float("nan"), dtype=torch.float64)
    for i, r in enumerate(local):
        buf[i, :len(r)] = torch.as_tensor(np.asarray(r, dtype=np.float64))
    buf = buf.to(dev)
    bufs = [torch.empty_like(buf) for _ in range(world)]
    dist.all_gather(bufs, buf)
    out = []
    for m, b in zip(metas, bufs):
        b = b.cpu().numpy()
        for i in range(int(m[0])):
            row = b[i]
            out.append(row[~np.isnan(row)] if np.isnan(row).any() else row)
    return out
