"""Generates the committed golden fixtures.  Run in the build container (needs /root/reference for the WAL
records); the GPU box only reads the outputs.

  knn_golden.npz      exact top-k (labels, fp32 distances) from the numpy oracle (oracle/cpu_ref.py) for a
                      grid of (metric, dim, store dtype, n, nq, k); inputs are NOT stored, they are
                      regenerated from the seeds with cpu_ref.synth_rows.
  wal_node_1.json     the ten WAL records the reference checked in under Static/wal/node_1/ (op, key,
                      timestamp, metadata, the non-zero head of the 512-d vector) -- the only fixture the
                      reference holds for this path -- plus the state a replay must end in.
"""
import glob
import json
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
from oracle import cpu_ref as R  # noqa: E402

CASES = [  # name, metric, dim, store, n, nq, k, scale, n_deleted
    ("l2_512", "l2", 512, "f32", 3000, 6, 10, 1.7, 0),
    ("ip_512", "ip", 512, "f32", 3000, 6, 10, 1.7, 0),
    ("cos_512", "cosine", 512, "f32", 3000, 6, 10, 1.7, 0),
    ("cos_512_k100_del", "cosine", 512, "f32", 2500, 4, 100, 1.0, 40),
    ("ip_768_f16", "ip", 768, "f16", 2000, 4, 10, 1.0, 0),
    ("l2_100_odd", "l2", 100, "f32", 777, 3, 7, 1.0, 5),
    ("cos_tiny", "cosine", 512, "f32", 7, 2, 10, 1.0, 0),
    # the widest rows the shard accepts, ALL-POSITIVE components (scale < 0 means |row| * |scale|): every product of
    # the contraction has the same sign, the worst case for the accumulation term of the tensor path's error bound
    ("ip_1408_pos", "ip", 1408, "f32", 1500, 9, 10, -1.0, 0),
    ("l2_1408_pos", "l2", 1408, "f32", 1500, 9, 10, -1.3, 0),
    ("cos_1024_pos_k100", "cosine", 1024, "f32", 1500, 5, 100, -1.0, 10),
    ("ip_2816_f16_pos", "ip", 2816, "f16", 1200, 9, 10, -1.0, 0),
]


def case_inputs(metric, dim, store, n, nq, scale, n_del):
    raw = R.synth_rows(R.SEED_DB, 0, n, dim) * np.float32(abs(scale))
    q = R.synth_rows(R.SEED_QUERY, 0, nq, dim) * np.float32(0.9)
    if scale < 0:
        raw, q = np.abs(raw), np.abs(q)
    deleted = list(range(3, 3 + 2 * n_del, 2))
    return raw, q, deleted


def main():
    out = {}
    for name, metric, dim, store, n, nq, k, scale, n_del in CASES:
        raw, q, deleted = case_inputs(metric, dim, store, n, nq, scale, n_del)
        stored = R.prepare_rows(raw, metric, store)
        labels, dist, cnt = R.knn_exact(q, stored, np.arange(n), k, metric, deleted=deleted)
        out[name + "/labels"] = labels
        out[name + "/dist"] = dist
        out[name + "/counts"] = cnt
    np.savez_compressed(os.path.join(HERE, "knn_golden.npz"), **out)

    ref = "/root/reference/Static/wal/node_1"
    if os.path.isdir(ref):
        recs = []
        for path in sorted(glob.glob(os.path.join(ref, "*.log"))):
            with open(path, "r", encoding="utf-8") as f:
                for line in f:
                    line = line.strip()
                    if not line:
                        continue
                    e = json.loads(line)
                    vec = e.get("vector")
                    head = None
                    if vec is not None:
                        nz = max((i for i, v in enumerate(vec) if v != 0.0), default=-1) + 1
                        head = {"len": len(vec), "head": vec[:nz]}
                    recs.append({"file": os.path.basename(path), "op_type": e["op_type"], "key": e["key"],
                                 "vector": head, "metadata": e.get("metadata"), "timestamp": e["timestamp"]})
        final = R.replay_wal_records(recs)
        live = [r["key"] for r in final if r["op_type"] == "PUT"]
        with open(os.path.join(HERE, "wal_node_1.json"), "w", encoding="utf-8") as f:
            json.dump({"source": "reference Static/wal/node_1/*.log", "records": recs,
                       "replay_order": [r["key"] for r in final], "live_after_replay": live}, f, indent=1)
    print("golden fixtures written")


if __name__ == "__main__":
    main()
