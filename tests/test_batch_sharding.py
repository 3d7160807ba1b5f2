"""CPU: host logic of the multi-GPU batched path -- contiguous sharding, the
counter-based input generator, and the result gather over a world_size-2 gloo
process group (the data path has no collective; only the per-network records
are gathered)."""
import os
import socket

import numpy as np
import pytest


def test_shard_range_partitions():
    from tncontract_b200.batch import shard_range
    for n in (0, 1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, r, world) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(spans[i][1] == spans[i + 1][0] for i in range(world - 1))
            sizes = [hi - lo for lo, hi in spans]
            assert max(sizes) - min(sizes) <= 1


def test_host_uniform_is_counter_based():
    from tncontract_b200.batch import host_uniform, network_key
    a = host_uniform(1000, 12345)
    assert a.min() >= 0.0 and a.max() < 1.0 and abs(a.mean() - 0.5) < 0.05
    assert np.array_equal(host_uniform(100, 12345, offset=900), a[900:])        # random access
    assert not np.array_equal(host_uniform(1000, 12346), a)
    keys = {network_key(3, n, w, s) for n in range(50) for w in range(2) for s in range(64)}
    assert len(keys) == 50 * 2 * 64


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    import torch.distributed as dist
    from tncontract_b200 import batch
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n = 11

    def fn(i):  # stand-in for one network's record: [index, f(index), bonds...]
        return np.array([float(i), i * 0.5] + [float(b) for b in range(1 + i % 3)])

    local = batch.run_sharded(n, fn, rank=rank, world=world)
    allr = batch.gather_results(local, world)
    dist.barrier()
    dist.destroy_process_group()
    q.put((rank, [r.tolist() for r in allr]))


def test_gather_results_gloo_world2():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    want = [[float(i), i * 0.5] + [float(b) for b in range(1 + i % 3)] for i in range(11)]
    assert got[0] == want and got[1] == want


@pytest.mark.gpu
def test_device_generator_matches_host_and_batch_record():
    """GPU: tnb_fill_uniform reproduces host_uniform bit for bit; one cfg-4-like network (reduced
    size) through the batched path equals the same network rebuilt on the host and run through the oracle;
    several streams give the same records as one."""
    from tncontract_b200 import batch
    from oracle import tn_oracle as o
    d = batch.device_uniform((3, 5, 7), 777, offset=4)
    assert np.array_equal(np.asarray(d).ravel(), batch.host_uniform(105, 777, offset=4))
    fn = lambda i: batch.overlap_norm_compress(3, i, nsites=10, physdim=3, bonddim=12, chi=6)
    recs = batch.run_sharded(6, fn, streams=1)
    recs3 = batch.run_sharded(6, fn, streams=3)
    for r1, r3 in zip(recs, recs3):
        assert np.array_equal(r1, r3)
    for net in (0, 5):
        ha = batch.random_mps(3, net, 0, 10, 3, 12, on_host=True)
        hb = batch.random_mps(3, net, 1, 10, 3, 12, on_host=True)
        ca = o.Chain([o.OT(x, l) for x, l in ha], "left", "right", "phys")
        cb = o.Chain([o.OT(x, l) for x, l in hb], "left", "right", "phys")
        ov, nrm = o.inner_product_mps(ca, cb), o.chain_norm(ca)
        o.svd_compress(ca, chi=6)
        r = recs[net]
        assert r[0] == net
        assert abs(r[1] - ov) <= 1e-10 * abs(ov) and abs(r[2] - nrm) <= 1e-10 * nrm
        assert abs(r[3] - o.chain_norm(ca, "right")) <= 1e-10 * nrm
        assert [int(b) for b in r[4:]] == ca.bonddims()
