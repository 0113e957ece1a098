"""Two GPUs, one process each: the sharded search with the fused NVLink exchange + merge kernel (K5x,
vdb_xchg_*) against (a) the same search with NCCL all-to-all + merge kernel and (b) the CPU oracle over the
whole set.  Skipped on boxes with a single GPU (run with `gpurun --gpus 2`)."""
import os
import socket

import numpy as np
import pytest

from oracle import cpu_ref as R

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, n, dim, nq, k, metric, steps, out):
    import torch
    import torch.distributed as dist
    import dvdb_b200 as vdb
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lib = vdb._ffi.lib()
        lo, hi = vdb.sharding.contiguous_range(n, rank, world)
        ix = vdb.Index(metric, dim, device=rank)
        ix.init_index(hi - lo)
        ix.add_items(R.synth_rows(R.SEED_DB, lo, hi - lo, dim), np.arange(lo, hi))
        stream = torch.cuda.current_stream().cuda_stream

        def gather_handles(mine: bytes):
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
            o = torch.empty((world, 64), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(o, t)
            return [bytes(o[r].cpu().numpy().tobytes()) for r in range(world)]

        px = vdb.PeerExchange(rank, rank, world, max_slice=nq // world, max_k=k, exchange_handles=gather_handles)
        sl = nq // world
        results = []
        for step in range(steps):
            q = torch.from_numpy(R.synth_rows(R.SEED_QUERY, 1000 * step, nq, dim)).to(dev)
            ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
            dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
            ix.search_device(q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), 0, stream)
            # fused NVLink exchange + merge
            o_ids = torch.empty((sl, k), dtype=torch.int64, device=dev)
            o_dd = torch.empty((sl, k), dtype=torch.float32, device=dev)
            px.merge(dd.data_ptr(), ids.data_ptr(), nq, k, o_dd.data_ptr(), o_ids.data_ptr(), stream)
            # NCCL all-to-all + merge kernel
            g_ids, g_dd = torch.empty_like(ids), torch.empty_like(dd)
            dist.all_to_all_single(g_ids, ids)
            dist.all_to_all_single(g_dd, dd)
            n_ids = torch.empty((sl, k), dtype=torch.int64, device=dev)
            n_dd = torch.empty((sl, k), dtype=torch.float32, device=dev)
            vdb._ffi.check(lib.vdb_merge_topk(g_dd.data_ptr(), g_ids.data_ptr(), world, sl, k, k, n_dd.data_ptr(),
                                              n_ids.data_ptr(), 1, rank, stream), "merge")
            torch.cuda.synchronize()
            assert torch.equal(o_ids, n_ids) and torch.equal(o_dd, n_dd), f"rank {rank} step {step}: p2p != nccl"
            results.append((o_ids.cpu().numpy(), o_dd.cpu().numpy()))
        out.put((rank, results))
        dist.barrier()
        px.close()
    finally:
        dist.destroy_process_group()


def _worker_ragged(rank, world, port, n, dim, k, out):
    """a single query (and 3 queries) through K5x: rank 0 owns query 0, ragged slices elsewhere"""
    import torch
    import torch.distributed as dist
    import dvdb_b200 as vdb
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, hi = vdb.sharding.contiguous_range(n, rank, world)
        ix = vdb.Index("l2", dim, device=rank)
        ix.init_index(hi - lo)
        ix.add_items(R.synth_rows(R.SEED_DB, lo, hi - lo, dim), np.arange(lo, hi))
        stream = torch.cuda.current_stream().cuda_stream

        def gather_handles(mine: bytes):
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
            o = torch.empty((world, 64), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(o, t)
            return [bytes(o[r].cpu().numpy().tobytes()) for r in range(world)]

        px = vdb.PeerExchange(rank, rank, world, max_slice=2, max_k=k, exchange_handles=gather_handles)
        res = {}
        for nq in (1, 3, 1):
            sl = (nq + world - 1) // world
            q = torch.from_numpy(R.synth_rows(R.SEED_QUERY, 7 * nq, nq, dim)).to(dev)
            ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
            dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
            ix.search_device(q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), 0, stream)
            o_ids = torch.full((sl, k), -7, dtype=torch.int64, device=dev)
            o_dd = torch.zeros((sl, k), dtype=torch.float32, device=dev)
            px.merge(dd.data_ptr(), ids.data_ptr(), nq, k, o_dd.data_ptr(), o_ids.data_ptr(), stream)
            torch.cuda.synchronize()
            res[nq] = (o_ids.cpu().numpy(), o_dd.cpu().numpy())
        out.put((rank, res))
        dist.barrier()
        px.close()
    finally:
        dist.destroy_process_group()


def test_peer_exchange_ragged_batches():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, dim, k, world = 3000, 512, 10, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_ragged, args=(r, world, port, n, dim, k, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    stored = R.prepare_rows(R.synth_rows(R.SEED_DB, 0, n, dim), "l2")
    for nq in (1, 3):
        sl = (nq + world - 1) // world
        q = R.synth_rows(R.SEED_QUERY, 7 * nq, nq, dim)
        for r in range(world):
            ids, dd = got[r][nq]
            for i in range(sl):
                qi = r * sl + i
                if qi < nq:
                    msg = R.check_topk(ids[i], dd[i], q[qi], stored, np.arange(n), k, "l2", rtol=1e-5)
                    assert msg is None, f"nq {nq} rank {r} query {qi}: {msg}"
                else:
                    assert (ids[i] == -7).all()          # rows a rank does not own stay untouched


@pytest.mark.parametrize("metric,nq,k", [("cosine", 64, 10), ("l2", 6, 100)])
def test_peer_exchange_equals_nccl_and_oracle(metric, nq, k):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, dim, world, steps = 4000, 512, 2, 3
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, dim, nq, k, metric, steps, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    stored = R.prepare_rows(R.synth_rows(R.SEED_DB, 0, n, dim), metric)
    sl = nq // world
    for step in range(steps):
        q = R.synth_rows(R.SEED_QUERY, 1000 * step, nq, dim)
        for r in range(world):
            ids, dd = got[r][step]
            for i in range(sl):
                msg = R.check_topk(ids[i], dd[i], q[r * sl + i], stored, np.arange(n), k, metric, rtol=1e-5)
                assert msg is None, f"step {step} rank {r} query {i}: {msg}"


def _worker_sharded_index(rank, world, port, n, dim, k, exchange, out):
    """the client API (ShardedIndex.search_host): slices in, slices out; even and ragged batches"""
    import torch
    import torch.distributed as dist
    import dvdb_b200 as vdb
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, hi = vdb.sharding.contiguous_range(n, rank, world)
        ix = vdb.Index("cosine", dim, device=rank)
        ix.init_index(hi - lo)
        ix.add_items(R.synth_rows(R.SEED_DB, lo, hi - lo, dim), np.arange(lo, hi))
        sx = vdb.ShardedIndex(ix, max_batch=64, max_k=k, exchange=exchange)
        res = {}
        for nq in (64, 1, 3, 64):
            q = R.synth_rows(R.SEED_QUERY, 11 * nq, nq, dim)
            if nq % world == 0:
                sl = nq // world
                mine = vdb.pinned_empty((sl, dim), np.float32)
                mine[:] = q[rank * sl:(rank + 1) * sl]
                ids, dd = sx.search_host(mine, k)
            else:
                ids, dd = sx.search_host(torch.from_numpy(q).pin_memory(), k, whole_batch=True)
            assert (lo_hi := sx.slice_of(nq)) and ids.shape[0] == lo_hi[1] - lo_hi[0]
            res[nq] = (lo_hi, ids.numpy().copy(), dd.numpy().copy())
        # pipelined form (submit_host / collect, two batches in flight) == one batch at a time
        batches = []
        for b in range(9):             # both query slots are reused several times, with two batch sizes
            nqb = 64 if b % 3 else 32
            q = R.synth_rows(R.SEED_QUERY, 500 + 64 * b, nqb, dim)
            sl = nqb // world
            batches.append(torch.from_numpy(q[rank * sl:(rank + 1) * sl].copy()).pin_memory())
        want = []
        for qb in batches:
            i1, d1 = sx.search_host(qb, k)
            want.append((i1.numpy().copy(), d1.numpy().copy()))
        got, pending = [], []
        for qb in batches:
            pending.append(sx.submit_host(qb, k))
            if len(pending) == 2:
                i2, d2 = sx.collect(pending.pop(0))
                got.append((i2.numpy().copy(), d2.numpy().copy()))
        while pending:
            i2, d2 = sx.collect(pending.pop(0))
            got.append((i2.numpy().copy(), d2.numpy().copy()))
        assert len(got) == len(want)
        for (gi, gd), (wi, wd) in zip(got, want):
            assert np.array_equal(gi, wi) and np.array_equal(gd, wd)
        out.put((rank, res))
        dist.barrier()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("exchange", ["p2p", "nccl"])
def test_sharded_index_client_api(exchange):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, dim, k, world = 5000, 512, 10, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_sharded_index, args=(r, world, port, n, dim, k, exchange, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    stored = R.prepare_rows(R.synth_rows(R.SEED_DB, 0, n, dim), "cosine")
    for nq in (64, 1, 3):
        q = R.synth_rows(R.SEED_QUERY, 11 * nq, nq, dim)
        covered = []
        for r in range(world):
            (lo, hi), ids, dd = got[r][nq]
            covered += list(range(lo, hi))
            for i in range(hi - lo):
                msg = R.check_topk(ids[i], dd[i], q[lo + i], stored, np.arange(n), k, "cosine", rtol=1e-5)
                assert msg is None, f"{exchange} nq {nq} rank {r} query {lo + i}: {msg}"
        assert covered == list(range(nq))            # every query answered by exactly one rank


# ---------------------------------------------------------------------------------------------
# robustness: several devices in one process, K5x next to other work, K5x when a peer misbehaves
# ---------------------------------------------------------------------------------------------
def test_two_indexes_on_two_devices_in_one_process(vdb):
    """LocalCoordinator's documented use: one handler (index) per GPU inside ONE process.  Every kernel's
    per-device launch configuration (dynamic shared memory opt-in) must be made on both devices: scan path,
    tensor path, merge."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, dim, k = 6000, 512, 10
    raw = R.synth_rows(R.SEED_DB, 0, n, dim)
    stored = R.prepare_rows(raw, "cosine")
    ixs = []
    for d in (0, 1):
        ix = vdb.Index("cosine", dim, device=d)
        ix.init_index(n)
        ix.add_items(raw, np.arange(n))
        ixs.append(ix)
    for nq in (1, 40):                                  # scan kernel, then the tensor path, on device 0 THEN 1
        q = R.synth_rows(R.SEED_QUERY, 5 * nq, nq, dim)
        for d, ix in enumerate(ixs):
            labels, dd, cnt = ix.knn_query_padded(q, k)
            for i in range(nq):
                msg = R.check_topk(labels[i], dd[i], q[i], stored, np.arange(n), k, "cosine", rtol=1e-5)
                assert msg is None, f"device {d} nq {nq} query {i}: {msg}"
    assert all(ix.get_stat("tensor_batches") == 1 and ix.get_stat("fallback_queries") == 0 for ix in ixs)
    for ix in ixs:
        ix.close()


def _worker_concurrent(rank, world, port, n, dim, nq, k, out):
    """K5x while another stream of the same GPU runs searches back to back: the cooperative launch may be delayed
    by them but must complete, with the right answer."""
    import threading
    import torch
    import torch.distributed as dist
    import dvdb_b200 as vdb
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        lo, hi = vdb.sharding.contiguous_range(n, rank, world)
        ix = vdb.Index("cosine", dim, device=rank)
        ix.init_index(hi - lo)
        ix.add_items(R.synth_rows(R.SEED_DB, lo, hi - lo, dim), np.arange(lo, hi))
        sx = vdb.ShardedIndex(ix, max_batch=nq, max_k=k, exchange="p2p")
        stop = threading.Event()
        noise_q = R.synth_rows(R.SEED_QUERY, 999, 300, dim)

        def noise():                                    # host-buffer searches on the library's own streams
            while not stop.is_set():
                ix.knn_query_padded(noise_q, k)

        t = threading.Thread(target=noise)
        t.start()
        res = []
        try:
            for step in range(6):
                q = torch.from_numpy(R.synth_rows(R.SEED_QUERY, 100 * step, nq, dim)).to(dev)
                dd, ids = sx.search_device(q, k)
                torch.cuda.synchronize()
                sx.px.status()
                res.append((ids.cpu().numpy().copy(), dd.cpu().numpy().copy()))
        finally:
            stop.set()
            t.join()
        out.put((rank, res))
        dist.barrier()
        sx.close()
    finally:
        dist.destroy_process_group()


def test_peer_exchange_with_a_concurrent_search_stream():
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n, dim, nq, k, world = 20000, 512, 256, 10, 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_concurrent, args=(r, world, port, n, dim, nq, k, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=300) for _ in range(world))
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    stored = R.prepare_rows(R.synth_rows(R.SEED_DB, 0, n, dim), "cosine")
    sl = nq // world
    for step in range(6):
        q = R.synth_rows(R.SEED_QUERY, 100 * step, nq, dim)
        for r in range(world):
            ids, dd = got[r][step]
            for i in range(0, sl, 7):
                msg = R.check_topk(ids[i], dd[i], q[r * sl + i], stored, np.arange(n), k, "cosine", rtol=1e-5)
                assert msg is None, f"step {step} rank {r} query {i}: {msg}"


def _worker_misbehaving_peer(rank, world, port, mode, out):
    """mode 'absent': rank 1 never launches its step; mode 'shape': rank 1 calls with another batch size.  Rank 0
    must come back (bounded wait), report the failure and refuse further steps; nothing may hang."""
    import torch
    import torch.distributed as dist
    os.environ["VDB_XCHG_TIMEOUT_MS"] = "1500"
    import dvdb_b200 as vdb
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        def gather_handles(mine: bytes):
            t = torch.frombuffer(bytearray(mine), dtype=torch.uint8).to(dev)
            o = torch.empty((world, 64), dtype=torch.uint8, device=dev)
            dist.all_gather_into_tensor(o, t)
            return [bytes(o[r].cpu().numpy().tobytes()) for r in range(world)]

        k, nq = 10, 8
        px = vdb.PeerExchange(rank, rank, world, max_slice=8, max_k=k, exchange_handles=gather_handles)
        stream = torch.cuda.current_stream().cuda_stream
        my_nq = nq if (rank == 0 or mode != "shape") else nq + 2
        ids = torch.arange(my_nq * k, dtype=torch.int64, device=dev).view(my_nq, k)
        dd = torch.rand((my_nq, k), device=dev).sort(dim=1).values
        o_ids = torch.zeros((8, k), dtype=torch.int64, device=dev)
        o_dd = torch.zeros((8, k), dtype=torch.float32, device=dev)
        verdict = "ok"
        if not (mode == "absent" and rank == 1):
            px.merge(dd.data_ptr(), ids.data_ptr(), my_nq, k, o_dd.data_ptr(), o_ids.data_ptr(), stream)
            torch.cuda.synchronize()
            try:
                px.status()
            except RuntimeError as e:
                verdict = str(e)
                assert (o_ids[: (my_nq + world - 1) // world].cpu().numpy() == -1).all()     # padded, not garbage
                with pytest.raises(RuntimeError):                                             # and it stays failed
                    px.merge(dd.data_ptr(), ids.data_ptr(), my_nq, k, o_dd.data_ptr(), o_ids.data_ptr(), stream)
        out.put((rank, verdict))
        dist.barrier()
        px.close()
    finally:
        dist.destroy_process_group()


@pytest.mark.parametrize("mode", ["absent", "shape"])
def test_peer_exchange_reports_a_misbehaving_peer(mode):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    world = 2
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker_misbehaving_peer, args=(r, world, port, mode, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = dict(out.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    if mode == "absent":
        assert "timed out waiting for rank 1" in got[0] and got[1] == "ok"
    else:
        assert "differs from rank" in got[0] and "differs from rank" in got[1]
