"""Times one search shape with CUDA events: python tools/time_search.py <nq> <path 0|1|2> [rows] [k] [metric] [dim]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dvdb_b200 as vdb
nq, path = int(sys.argv[1]), int(sys.argv[2])
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
metric = sys.argv[5] if len(sys.argv) > 5 else "cosine"
dim = int(sys.argv[6]) if len(sys.argv) > 6 else 512
ix = vdb.Index(metric, dim); ix.init_index(rows); ix.add_synthetic(0xD5B200, 0, rows)
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
q = torch.empty((nq, dim), dtype=torch.float32, device=dev)
vdb._ffi.check(vdb._ffi.lib().vdb_synth_dev(0xC0FFEE, 0, nq, dim, q.data_ptr(), st), "synth")
ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
ix.set_option("path", path)
for _ in range(5):
    ix.search_device(q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), 0, st)
torch.cuda.synchronize()
n = 50
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(n):
    ix.search_device(q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), 0, st)
e1.record(); torch.cuda.synchronize()
print(f"nq={nq} path={path} rows={rows} dim={dim} k={k}: {e0.elapsed_time(e1) / n * 1e3:8.1f} us  "
      f"(shadow scans {ix.get_stat('shadow_scans')}, tensor {ix.get_stat('tensor_batches')}, fallback {ix.get_stat('fallback_queries')})  env "
      + " ".join(f"{a}={os.environ[a]}" for a in os.environ if a.startswith("VDB_")))
