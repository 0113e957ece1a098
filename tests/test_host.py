"""CPU: host-side mirror of the reference interface -- WAL format/replay, handler semantics (with a test-double
index; the GPU index runs the same tests in test_gpu_handler.py), coordinator merge and routing."""
import json
import os

import numpy as np
import pytest

from oracle import cpu_ref as R
from tests import fake_index

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def vec(x, dim=8):
    v = [0.0] * dim
    v[0] = float(x)
    return v


class Recorder:
    def __init__(self):
        self.ops = []

    def put(self, data, replay_mode=False):
        assert replay_mode
        self.ops.append(("PUT", data.key, data.vector[:3]))

    def delete(self, key, replay_mode=False):
        assert replay_mode
        self.ops.append(("DELETE", key))


def test_wal_layout_and_record_format(vdb, tmp_path):
    """<root>/data/wal_<ms>.log JSON lines + <root>/checkpoint/checkpoint_ts.txt (wal_manager.py:13-19,91-98)."""
    w = vdb.WALManager(str(tmp_path / "wal"), node_id="node_1", fsync=False)
    w.write_log("PUT", "a", [0.5, 0.25], {"tag": "t"}, timestamp=1000)
    w.write_log("DELETE", "a", timestamp=1001)
    files = os.listdir(tmp_path / "wal" / "data")
    assert len(files) == 1 and files[0].startswith("wal_") and files[0].endswith(".log")
    lines = open(tmp_path / "wal" / "data" / files[0], encoding="utf-8").read().splitlines()
    assert len(lines) == 2                                   # a real append: the reference's rename keeps only the last line
    rec = json.loads(lines[0])
    # the reference's six fields in its order (wal_manager.py:91-98), then the additive sequence number
    assert list(rec) == ["op_type", "key", "vector", "metadata", "timestamp", "node_id", "seq"]
    assert rec == {"op_type": "PUT", "key": "a", "vector": [0.5, 0.25], "metadata": {"tag": "t"}, "timestamp": 1000,
                   "node_id": "node_1", "seq": 1}
    assert json.loads(lines[1])["seq"] == 2
    assert json.loads(lines[1])["vector"] is None and os.path.isdir(tmp_path / "wal" / "checkpoint")


def test_wal_replays_reference_fixture(vdb, tmp_path):
    """The reference's own records, renamed to the name pattern its WALManager lists (wal_manager.py:33-36)."""
    fx = json.load(open(os.path.join(GOLD, "wal_node_1.json"), encoding="utf-8"))
    root = tmp_path / "wal"
    os.makedirs(root / "data")
    for r in fx["records"]:
        v = None if r["vector"] is None else r["vector"]["head"] + [0.0] * (r["vector"]["len"] - len(r["vector"]["head"]))
        line = json.dumps({"op_type": r["op_type"], "key": r["key"], "vector": v, "metadata": r["metadata"], "timestamp": r["timestamp"]})
        with open(root / "data" / ("wal_" + r["file"]), "a", encoding="utf-8") as f:     # no node_id, no trailing newline: legacy records
            f.write(line)
    w = vdb.WALManager(str(root), fsync=False)
    rec = Recorder()
    assert w.replay(rec) == 5
    assert [o[1] for o in rec.ops] == fx["replay_order"]
    assert [o[1] for o in rec.ops if o[0] == "PUT"] == fx["live_after_replay"]
    assert rec.ops[2] == ("PUT", "test_2", [0.1, 0.2, 0.0])
    assert w.checkpoint_ts == 1766758239137 and w.replay(rec) == 0          # second call is a no-op (:118-120)
    rec2 = Recorder()
    w.replay_incremental(rec2, 1766757478851)
    assert rec2.ops == [("DELETE", "test_8081"), ("DELETE", "test_1")]


def test_wal_rotation_group_commit_and_torn_tail(vdb, tmp_path):
    w = vdb.WALManager(str(tmp_path / "w"), max_log_size=4000, fsync=False)
    n = w.write_batch(("PUT", f"k{i}", vec(i, 64), {}, 10 + i) for i in range(40))
    assert n == 40 and len(os.listdir(tmp_path / "w" / "data")) >= 1
    for i in range(40, 60):
        w.write_log("PUT", f"k{i}", vec(i, 64), {}, timestamp=10 + i)
    assert len(os.listdir(tmp_path / "w" / "data")) >= 2                     # rotated at max_log_size (:108-109)
    with open(w.current_log_file, "a") as f:
        f.write('{"op_type": "PUT", "key": "torn", "vec')                    # crash mid-line
    rec = Recorder()
    assert vdb.WALManager(str(tmp_path / "w"), fsync=False).replay(rec) == 60
    assert [o[1] for o in rec.ops] == [f"k{i}" for i in range(60)]


def make_handler(vdb, tmp_path, **kw):
    return vdb.GpuVectorNodeHandler("node_1", storage_root=str(tmp_path), dim=8, max_elements=64, fsync=False,
                                    index_factory=fake_index.factory, **kw)


def test_handler_matches_datanode_model(vdb, tmp_path):
    """same edge cases as oracle datanode_search (reference handler.py:222-428)."""
    h = make_handler(vdb, tmp_path)
    model = R.DatanodeModel(dim=8)
    q = vdb.SearchRequest(query_vector=vec(1), top_k=0)
    r = h.search(q)
    assert r.success and r.search_result.keys == [] and r.search_result.vectors == []
    for i in range(12):
        assert h.put(vdb.VectorData(key=f"k{i}", vector=vec(i), metadata={"i": str(i)})).success
        model.put(f"k{i}", vec(i))
    bad = h.put(vdb.VectorData(key="bad", vector=[1.0, 2.0]))
    assert not bad.success and "dim mismatch" in bad.message
    r = h.search(q)                                                            # top_k 0 -> 5
    assert r.search_result.keys == R.datanode_search_exact_live(model, vec(1), 0)[1]
    assert r.search_result.scores == [0.0, 1.0, 1.0, 4.0, 9.0]
    assert r.search_result.vectors[0].metadata == {"i": "1"} and r.search_result.vectors[0].vector == vec(1)
    h.put(vdb.VectorData(key="k1", vector=vec(100)))                            # overwrite
    h.delete("k2")
    model.put("k1", vec(100)); model.delete("k2")
    assert not h.delete("k2").success and not h.get("k2").success
    assert h.get("k1").vector_data.vector == vec(100)
    r = h.search(vdb.SearchRequest(query_vector=vec(1), top_k=3))
    assert r.search_result.keys == ["k0", "k3", "k4"] == R.datanode_search_exact_live(model, vec(1), 3)[1]
    assert h.hnsw_index.get_current_count() == 13 and h.deleted_ids == {1, 2}
    ks, ss = h.search_batch([vec(1), vec(11)], 2)
    assert ks == [["k0", "k3"], ["k11", "k10"]] and ss[1] == [0.0, 1.0]


def test_handler_reference_quirk_mode(vdb, tmp_path):
    h = make_handler(vdb, tmp_path, reference_quirks=True)
    for i in range(8):
        h.put(vdb.VectorData(key=f"s{i}", vector=vec(i)))
    r = h.search(vdb.SearchRequest(query_vector=vec(1), top_k=5))              # 2k = 10 > 8 rows (handler.py:364-369)
    assert not r.success and r.message == "HNSW index corrupted, search aborted"
    assert h.search(vdb.SearchRequest(query_vector=vec(1), top_k=4)).success


def test_handler_recovers_from_wal_and_checkpoint(vdb, tmp_path):
    h = make_handler(vdb, tmp_path, checkpoint_every=5)
    for i in range(7):
        h.put(vdb.VectorData(key=f"k{i}", vector=vec(i), metadata={"n": str(i)}))
    h.delete("k3")
    h.put(vdb.VectorData(key="k0", vector=vec(50)))
    want = h.search(vdb.SearchRequest(query_vector=vec(2), top_k=10)).search_result
    cps = [d for d in os.listdir(tmp_path / "node_1" / "checkpoint") if d.startswith("checkpoint_")]
    assert len(cps) == 1                                                       # taken at id 5 (handler.py:316-317)
    cp = tmp_path / "node_1" / "checkpoint" / cps[0]
    # the reference's checkpoint layout (handler.py:156-178) + the exact WAL record number the checkpoint contains
    assert sorted(os.listdir(cp)) == ["deleted_ids.json", "index.bin", "index.bin.npz", "leveldb_data", "wal_pos.txt",
                                      "wal_seq.txt"]
    assert os.listdir(cp / "leveldb_data") == ["kv_pos.json"]                  # a position in the append-only store
    # crash: a new handler on the same directory = newest checkpoint + incremental WAL replay (handler.py:181-219)
    h2 = make_handler(vdb, tmp_path, checkpoint_every=5)
    got = h2.search(vdb.SearchRequest(query_vector=vec(2), top_k=10)).search_result
    assert got.keys == want.keys and got.scores == want.scores
    assert h2.get("k3").success is False and h2.get("k0").vector_data.vector == vec(50)
    # no checkpoint at all: full WAL replay
    h3 = vdb.GpuVectorNodeHandler("node_2", storage_root=str(tmp_path), dim=8, max_elements=64, fsync=False,
                                  index_factory=fake_index.factory, checkpoint_every=0)
    for i in range(4):
        h3.put(vdb.VectorData(key=f"z{i}", vector=vec(i)))
    h3.put(vdb.VectorData(key="z1", vector=vec(9)))
    h4 = vdb.GpuVectorNodeHandler("node_2", storage_root=str(tmp_path), dim=8, max_elements=64, fsync=False,
                                  index_factory=fake_index.factory, checkpoint_every=0)
    r = h4.search(vdb.SearchRequest(query_vector=vec(0), top_k=4)).search_result
    assert r.keys == ["z0", "z2", "z3", "z1"]                                   # z1 replayed with its LAST vector


def test_recovery_is_exact_at_the_checkpoint_boundary(vdb, tmp_path, monkeypatch):
    """Operations that share the checkpoint's millisecond are neither lost nor applied twice: the checkpoint records
    the WAL sequence number it contains (the reference compares wall-clock milliseconds, wal_manager.py:214-215)."""
    import time as _time
    monkeypatch.setattr(_time, "time", lambda: 1_700_000_000.000)              # every record, and the checkpoint: same ms
    h = make_handler(vdb, tmp_path, checkpoint_every=3)
    for i in range(3):
        h.put(vdb.VectorData(key=f"a{i}", vector=vec(i)))                      # checkpoint taken after the third
    h.put(vdb.VectorData(key="late", vector=vec(7)))                           # same millisecond, AFTER the checkpoint
    h.delete("a1")
    want = h.search(vdb.SearchRequest(query_vector=vec(0), top_k=10)).search_result
    h2 = make_handler(vdb, tmp_path, checkpoint_every=3)
    got = h2.search(vdb.SearchRequest(query_vector=vec(0), top_k=10)).search_result
    assert got.keys == want.keys == ["a0", "a2", "late"] and got.scores == want.scores
    assert h2.hnsw_index.get_current_count() == 4                              # 3 from the snapshot + 1 replayed put


def test_checkpoints_are_pruned_and_bulk_puts_log_by_reference(vdb, tmp_path):
    h = make_handler(vdb, tmp_path, checkpoint_every=4)
    keys = [f"b{i}" for i in range(20)]
    vecs = np.stack([np.array(vec(i), dtype=np.float32) for i in range(20)])
    for lo in range(0, 20, 5):
        assert h.put_arrays(keys[lo:lo + 5], vecs[lo:lo + 5], [{"i": str(i)} for i in range(lo, lo + 5)]).success
    cps = sorted(d for d in os.listdir(tmp_path / "node_1" / "checkpoint"))
    assert len(cps) == 2                                                       # keep_checkpoints = 2
    wal_dir = tmp_path / "node_1" / "wal" / "data"
    recs = [json.loads(l) for f in sorted(os.listdir(wal_dir)) for l in open(wal_dir / f, encoding="utf-8")]
    assert len(recs) == 20 and all(r["vector"] is None and r["hnsw_id"] == i for i, r in enumerate(recs))
    h.put(vdb.VectorData(key="b3", vector=vec(33)))                            # inline JSON vector, overwrites b3
    # recovery reads the by-reference vectors back from the raw-vector file
    h2 = make_handler(vdb, tmp_path, checkpoint_every=4)
    assert h2.get("b7").vector_data.vector == vec(7) and h2.get("b7").vector_data.metadata == {"i": "7"}
    assert h2.get("b3").vector_data.vector == vec(33)
    r = h2.search(vdb.SearchRequest(query_vector=vec(19), top_k=3)).search_result
    assert r.keys == ["b19", "b18", "b17"]
    # full replay (no checkpoint left): same state
    import shutil
    shutil.rmtree(tmp_path / "node_1" / "checkpoint")
    os.makedirs(tmp_path / "node_1" / "checkpoint")
    h3 = make_handler(vdb, tmp_path, checkpoint_every=0)
    assert h3.get("b7").vector_data.vector == vec(7) and h3.get("b3").vector_data.vector == vec(33)
    assert len(h3.store) == 20


def test_key_store_roundtrip_and_torn_tail(vdb, tmp_path):
    KeyStore = vdb.kvstore.KeyStore
    s = KeyStore(str(tmp_path / "kv"), 8, fsync=False)
    s.put(0, "x", np.arange(8, dtype=np.float32), {"a": "1"})
    s.put_batch([1, 2, 5], ["y", "z", "x"], np.ones((3, 8), np.float32) * np.array([[1], [2], [5]], np.float32), [None, {}, {"b": "2"}])
    assert s.id_of("x") == 5 and s.key_of(0) == "" and s.key_of(5) == "x" and s.metadata(5) == {"b": "2"} and s.metadata(1) == {}
    assert s.vector(2).tolist() == [2.0] * 8 and len(s) == 3
    pos = s.position()
    assert s.delete("y") == 1 and s.delete("y") == -1 and "y" not in s
    s.close()
    with open(tmp_path / "kv" / "keys.log", "ab") as f:
        f.write(b'{"i": 9, "k": "torn')                                        # a crash in the middle of an append
    s2 = KeyStore(str(tmp_path / "kv"), 8, fsync=False)
    assert sorted(s2.keys()) == ["x", "z"] and s2.id_of("torn") == -1 and s2.id_of("x") == 5
    s2.rollback(pos)                                                           # back to before the delete
    assert sorted(s2.keys()) == ["x", "y", "z"] and s2.id_of("y") == 1
    s2.close()


def test_put_batch_equals_sequential_puts(vdb, tmp_path):
    a = make_handler(vdb, tmp_path / "a")
    b = make_handler(vdb, tmp_path / "b")
    items = [vdb.VectorData(key=f"k{i % 9}", vector=vec(i), metadata={}) for i in range(14)]   # keys repeat
    for d in items:
        a.put(d)
    assert b.put_batch(items).success
    for x in (0.0, 6.5, 13.0):
        ra = a.search(vdb.SearchRequest(query_vector=vec(x), top_k=6)).search_result
        rb = b.search(vdb.SearchRequest(query_vector=vec(x), top_k=6)).search_result
        assert ra.keys == rb.keys and ra.scores == rb.scores


def test_coordinator_merge_and_routing(vdb, tmp_path):
    res = [vdb.SearchResult(keys=["x", "y", "z"], scores=[0.1, 0.5, 0.9], vectors=[1, 2, 3]), None,
           vdb.SearchResult(keys=["y", "w", "v"], scores=[0.2, 0.5, 0.05], vectors=[4, 5, 6])]
    m = vdb.merge_search_results(res, 4)
    want_k, want_s = R.coordinator_merge([(["x", "y", "z"], [0.1, 0.5, 0.9]), (["y", "w", "v"], [0.2, 0.5, 0.05])], 4)
    assert m.keys == want_k == ["v", "x", "y", "w"] and m.scores == want_s and m.vectors == [6, 1, 2, 5]
    assert vdb.merge_search_results([None, vdb.SearchResult([], [], [])], 3).keys == []
    assert vdb.get_shard_id("test_1", 4) == R.get_shard_id("test_1", 4) == 3
    assert vdb.assign_shards_to_nodes(["a", "b"], 4, 2) == R.assign_shards_to_nodes(["a", "b"], 4, 2)

    nodes = {f"n{i}": make_handler(vdb, tmp_path / f"n{i}") for i in range(4)}
    coord = vdb.LocalCoordinator(nodes)
    model = R.DatanodeModel(dim=8)
    for i in range(40):
        key = f"key_{i}"
        assert coord.put(vdb.VectorData(key=key, vector=vec(i * 0.5))).success
        model.put(key, vec(i * 0.5))
        owner = nodes[coord.mapping[R.get_shard_id(key, 4)]["master"]]
        assert owner.get(key).success
    assert sum(h.hnsw_index.get_current_count() for h in nodes.values()) == 40
    coord.delete("key_3")
    model.delete("key_3")
    r = coord.search(vdb.SearchRequest(query_vector=vec(1.4), top_k=5)).search_result
    assert r.keys == R.datanode_search_exact_live(model, vec(1.4), 5)[1]        # sharded == single node
    assert coord.get("key_3").success is False
    serial = vdb.LocalCoordinator(nodes, parallel=False)                       # the reference's serial loop
    for x in (0.0, 1.4, 7.3, 19.0):
        for k in (1, 5, 50):
            a = coord.search(vdb.SearchRequest(query_vector=vec(x), top_k=k)).search_result
            b = serial.search(vdb.SearchRequest(query_vector=vec(x), top_k=k)).search_result
            assert a.keys == b.keys == R.datanode_search_exact_live(model, vec(x), k)[1] and a.scores == b.scores
    ks, ss = coord.search_batch([vec(1.4), vec(7.3)], 5, gpu_merge=False)      # host merge (the GPU merge: tests/test_gpu_handler.py)
    assert ks == [R.datanode_search_exact_live(model, vec(x), 5)[1] for x in (1.4, 7.3)]
    assert ss[0] == coord.search(vdb.SearchRequest(query_vector=vec(1.4), top_k=5)).search_result.scores

    class Down:                                                                # a node that fails is skipped (:198-199)
        def search(self, req):
            raise ConnectionError("node down")
    flaky = vdb.LocalCoordinator({**nodes, "zz": Down()}, shard_count=4)
    assert flaky.search(vdb.SearchRequest(query_vector=vec(1.4), top_k=5)).search_result.keys == r.keys
    assert vdb.LocalCoordinator({}).search(vdb.SearchRequest(query_vector=vec(1), top_k=1)).success is False


# ---------------------------------------------------------------------------------------------
# micro-batching of concurrent single-query searches
# ---------------------------------------------------------------------------------------------
def test_micro_batcher_coalesces_concurrent_searches(vdb, tmp_path):
    """8 worker threads x 25 requests with different k: every answer equals the one-at-a-time handler's, and
    the index saw far fewer queries than requests."""
    import threading
    plain = make_handler(vdb, tmp_path / "a")
    fast = make_handler(vdb, tmp_path / "b", micro_batch_wait_s=2e-3)
    for h in (plain, fast):
        for i in range(40):
            assert h.put(vdb.VectorData(key=f"k{i}", vector=vec(i), metadata={"i": str(i)})).success
        h.delete("k7")
    reqs = [vdb.SearchRequest(query_vector=vec(3 * j % 41), top_k=(j % 6)) for j in range(25)]
    want = [plain.search(r).search_result for r in reqs]
    got, errs = {}, []

    def worker(t):
        try:
            for j, r in enumerate(reqs):
                got[(t, j)] = fast.search(r)
        except Exception as e:          # noqa
            errs.append(e)

    ts = [threading.Thread(target=worker, args=(t,)) for t in range(8)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert not errs, errs
    for (t, j), resp in got.items():
        assert resp.success
        assert resp.search_result.keys == want[j].keys and resp.search_result.scores == want[j].scores
        assert [v.metadata for v in resp.search_result.vectors] == [v.metadata for v in want[j].vectors]
    b = fast._batcher
    assert b.requests == 200 and b.batches < 120          # coalesced: 8 threads -> batches of several requests


def test_micro_batcher_edge_cases(vdb, tmp_path):
    """empty index, reference-quirk failure, an index error reaching every request of the batch, and batches
    larger than max_batch (leadership passes on)."""
    import threading
    h = make_handler(vdb, tmp_path, micro_batch_wait_s=0.0, micro_batch_max=2, reference_quirks=True)
    r = h.search(vdb.SearchRequest(query_vector=vec(1), top_k=3))
    assert r.success and r.search_result.keys == []
    for i in range(5):
        h.put(vdb.VectorData(key=f"k{i}", vector=vec(i)))
    assert not h.search(vdb.SearchRequest(query_vector=vec(1), top_k=3)).success          # 2k > count
    assert h.search(vdb.SearchRequest(query_vector=vec(1), top_k=2)).search_result.keys == ["k1", "k0"]
    out = []
    ts = [threading.Thread(target=lambda: out.append(h.search(vdb.SearchRequest(query_vector=vec(4), top_k=1))))
          for _ in range(7)]
    [t.start() for t in ts]
    [t.join() for t in ts]
    assert len(out) == 7 and all(o.success and o.search_result.keys == ["k4"] for o in out)

    def boom(q, k):
        raise RuntimeError("index gone")
    h.hnsw_index.knn_query_padded = boom
    bad = h.search(vdb.SearchRequest(query_vector=vec(1), top_k=2))
    assert not bad.success and "HNSW index corrupted" in bad.message


# ---------------------------------------------------------------------------------------------
# bench.py contract (the reference arm runs without a GPU)
# ---------------------------------------------------------------------------------------------
def test_bench_reference_arm_prints_the_contract_line():
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # torchrun exports OMP_NUM_THREADS=1 to its workers: the arm must still use every host core
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "1", "--steps", "2",
                          "--warmup", "1", "--rows", "20000", "--cpu-queries", "16"], capture_output=True, text=True,
                         timeout=600, cwd=root, env=dict(os.environ, OMP_NUM_THREADS="1"))
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.startswith("{")]
    assert len(lines) == 1                                   # ONE json line
    d = json.loads(lines[0])
    for key in ("metric", "value", "unit", "n_gpus", "steps", "warmup", "ms_per_step", "higher_is_better", "scaling",
                "vs_baseline", "dtype", "data", "config", "e2e", "cpu_baseline", "impl"):
        assert key in d, key
    assert d["impl"] == "reference" and d["unit"] == "queries/s" and d["higher_is_better"] is True
    assert d["steps"] == 2 and d["warmup"] == 1 and d["n_gpus"] == 1 and d["vs_baseline"] is None
    assert d["config"]["workload"].startswith("20000x512 f32 cosine top-10") and "model" not in d["config"]
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["sample"]
    assert d["cpu_baseline"]["cores"] == len(os.sched_getaffinity(0))
    assert "brute-force" in d["config"]["note"]              # labelled for what it is: not the HNSW walk
    assert d["cpu_baseline"]["value"] == d["value"] == d["e2e"]["value"] > 0
    assert d["e2e"]["h2d_bytes_per_step"] == 0 and d["e2e"]["d2h_bytes_per_step"] == 0
    # the approximate walk the reference really does (HNSW restatement), informative, beside the exact value
    hp = d["hnsw_port"]
    assert "error" not in hp, hp
    assert hp["rows"] == 20000 and hp["M"] == 32 and hp["ef_construction"] == 128 and hp["ef"] == 50 and hp["qps"] > 0
    assert 0.0 <= hp["recall_at_k_vs_exact"] <= 1.0 and hp["clustered"]["recall_at_k_vs_exact"] >= 0.95
    # a rank other than 0 prints nothing and exits 0
    env = dict(os.environ, RANK="1", WORLD_SIZE="2")
    out = subprocess.run([sys.executable, os.path.join(root, "bench.py"), "--impl", "reference", "--gpus", "2"],
                         capture_output=True, text=True, timeout=120, cwd=root, env=env)
    assert out.returncode == 0 and out.stdout.strip() == ""


def test_bench_checkers_on_cpu():
    """bench.py's parity helpers: the chunked whole-database oracle equals one exact scan, and compare_topk applies
    the parity bar (ids identical, swaps only inside distance ties within 1e-5 relative)."""
    import importlib.util
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("bench_mod", os.path.join(root, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    sys.modules["bench_mod"] = bench                         # dataclasses look the module up by name
    spec.loader.exec_module(bench)
    from oracle import c_ref
    for metric, store in (("cosine", "f32"), ("l2", "f32"), ("ip", "f16")):
        wl = bench.Workload("t", 3000, 96, 7, metric, store, 16)
        q_idx = np.array([0, 3, 11])
        ids, dd, info = bench.oracle_topk(wl, q_idx, budget_s=60.0, chunk=700)       # 5 ragged chunks
        assert info["rows_scanned"] == 3000
        stored = bench.stored_rows(c_ref, wl, 0, 3000)
        q = np.concatenate([c_ref.synth_rows(bench.SEED_QUERY, int(i), 1, 96) for i in q_idx])
        want_i, want_d, _ = c_ref.knn(q, stored, None, 7, metric)
        assert np.array_equal(ids, want_i) and np.array_equal(dd.view(np.uint32), want_d.view(np.uint32))
        assert np.array_equal(bench.oracle_distances_of(wl, q_idx, ids).view(np.uint32), want_d.view(np.uint32))
        ok = bench.compare_topk(ids, dd, want_i, want_d)
        assert ok["parity_ok"] and ok["recall_at_k"] == 1.0 and ok["ordered_ids_equal"] == 1.0
    # a wrong id at a clear gap fails; a swap inside a tie passes; a distance off by 1e-4 fails
    wi = np.array([[5, 9, 2]]); wd = np.array([[0.1, 0.2, 0.2000001]], dtype=np.float32)
    assert not bench.compare_topk(np.array([[5, 9, 7]]), wd, wi, np.array([[0.1, 0.2, 0.3]], dtype=np.float32))["parity_ok"]
    assert bench.compare_topk(np.array([[5, 2, 9]]), wd, wi, wd)["parity_ok"]
    assert not bench.compare_topk(wi, wd + np.float32(1e-4), wi, wd)["parity_ok"]
    # the workloads bench.py runs at 8 GPUs are BASELINE.json's configs 3 and 4 at full size
    class A: configs = "auto"
    w3, w4 = bench.extra_workloads(A, 8)
    assert (w3.rows, w3.dim, w3.k, w3.metric, w3.store, w3.batch) == (10_000_000, 512, 100, "l2", "f32", 4096)
    assert (w4.rows, w4.dim, w4.k, w4.metric, w4.store) == (100_000_000, 768, 10, "ip", "f16")
    assert bench.extra_workloads(A, 4) == []
