"""CPU, world_size 2 over gloo: the N>1 host logic (shard the rows, every rank searches its shard, all-gather
the per-rank top-k, merge by (distance, id)) == one exact search over the whole set.  On GPUs the same class
runs over NCCL with the CUDA search and merge kernels (bench.py --gpus N)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import cpu_ref as R


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dim, nq, k, metric, out):
    import dvdb_b200 as vdb
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = vdb.sharding.contiguous_range(n, rank, world)
        stored = R.prepare_rows(R.synth_rows(R.SEED_DB, lo, hi - lo, dim), metric)
        labels = np.arange(lo, hi)
        # a few exact duplicates across shards -> cross-rank distance ties, broken by id
        if rank == 1:
            stored[:3] = R.prepare_rows(R.synth_rows(R.SEED_DB, 0, 3, dim), metric)
        q = R.synth_rows(R.SEED_DB, 0, nq, dim)

        def local_search(queries, kk):
            l, d, _ = R.knn_exact(queries.numpy(), stored, labels, kk, metric)
            return torch.from_numpy(l), torch.from_numpy(d)

        def merge(g_d, g_i, kk):
            d, i = R.merge_topk_by_id(g_d.numpy(), g_i.numpy(), kk)
            return torch.from_numpy(d), torch.from_numpy(i)

        s = vdb.ShardedSearcher(local_search, merge)
        d, i = s.search(torch.from_numpy(q), k)
        if rank == 0:
            out.put((d.numpy(), i.numpy()))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("metric,nq", [("l2", 5), ("cosine", 5), ("l2", 6)])
def test_sharded_search_equals_single_shard(metric, nq):
    """nq = 5: does not divide by 2 -> all-gather + merge everywhere; nq = 6: all-to-all by query slice."""
    n, dim, k, world = 1200, 64, 10, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dim, nq, k, metric, out)) for r in range(world)]
    for p in procs:
        p.start()
    got_d, got_i = out.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    full = R.prepare_rows(R.synth_rows(R.SEED_DB, 0, n, dim), metric)
    full[600:603] = full[0:3]                                        # the duplicates rank 1 planted
    q = R.synth_rows(R.SEED_DB, 0, nq, dim)
    want_i, want_d, _ = R.knn_exact(q, full, np.arange(n), k, metric)
    assert np.array_equal(got_i, want_i) and np.array_equal(got_d, want_d)
    assert got_i[0, 0] == 0 and got_i[0, 1] == 600                   # equal distance: smaller id first
