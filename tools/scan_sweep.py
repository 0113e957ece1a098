#!/usr/bin/env python
"""Sweep of the K1 scan kernel's launch shape (CTAs per SM x ring stages) on one GPU.
Prints: config, kernel us (profile callbacks, mean of `reps`), GB/s, fraction of the measured HBM peak.
usage: python tools/scan_sweep.py [rows] [dim] [f32|f16]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import dvdb_b200 as vdb

rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
dim = int(sys.argv[2]) if len(sys.argv) > 2 else 512
store = sys.argv[3] if len(sys.argv) > 3 else "f32"
nq = int(sys.argv[4]) if len(sys.argv) > 4 else 1
peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
lib = vdb._ffi.lib()
ix = vdb.Index("cosine", dim, store_dtype=store)
ix.init_index(rows)
ix.add_synthetic(0xD5B200, 0, rows)
dev = torch.device("cuda", 0)
st = torch.cuda.current_stream().cuda_stream
q = torch.empty((nq, dim), dtype=torch.float32, device=dev)
vdb._ffi.check(lib.vdb_synth_dev(0xC0FFEE, 0, nq, dim, q.data_ptr(), st), "synth")
ids = torch.empty((nq, 10), dtype=torch.int64, device=dev)
dd = torch.empty((nq, 10), dtype=torch.float32, device=dev)
nbytes = rows * ix.get_stat("ld") * (2 if store == "f16" else 4)
configs = [(c.split(":") + ["0"])[:3] for c in os.environ.get("SWEEP", "1:6,2:2,2:3,3:2").split(",")]
for rep in range(2):
    for ctas, stages, dbg in configs:
        os.environ["VDB_SCAN_CTAS"] = ctas
        os.environ["VDB_SCAN_STAGES"] = stages
        os.environ["VDB_SCAN_DBG"] = dbg
        for _ in range(5):
            ix.search_device(q.data_ptr(), nq, 10, ids.data_ptr(), dd.data_ptr(), 0, st)
        torch.cuda.synchronize()
        ix.set_option("profile", 1)
        for _ in range(30):
            ix.search_device(q.data_ptr(), nq, 10, ids.data_ptr(), dd.data_ptr(), 0, st)
        torch.cuda.synchronize()
        ix.set_option("profile", 0)
        cnt = ix.get_stat("profile_count")          # read before profile_ns (which drains the events)
        us = ix.get_stat("profile_ns") / max(cnt, 1) / 1e3
        gbs = nbytes / us / 1e3
        print(f"rows={rows} dim={dim} {store} nq={nq} ctas/sm={ctas} stages={stages} dbg={dbg}: {us:.1f} us {gbs:.0f} GB/s {gbs/peak:.4f}", flush=True)
