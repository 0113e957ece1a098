"""A few searches of one shape (for ncu launch lists): python tools/one_search.py <nq> <path 0|1|2> [rows] [k] [metric]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import dvdb_b200 as vdb
nq, path = int(sys.argv[1]), int(sys.argv[2])
rows = int(sys.argv[3]) if len(sys.argv) > 3 else 1_000_000
k = int(sys.argv[4]) if len(sys.argv) > 4 else 10
metric = sys.argv[5] if len(sys.argv) > 5 else "cosine"
ix = vdb.Index(metric, 512); ix.init_index(rows); ix.add_synthetic(0xD5B200, 0, rows)
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
q = torch.empty((nq, 512), dtype=torch.float32, device=dev)
vdb._ffi.check(vdb._ffi.lib().vdb_synth_dev(0xC0FFEE, 0, nq, 512, q.data_ptr(), st), "synth")
ids = torch.empty((nq, k), dtype=torch.int64, device=dev); dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
ix.set_option("path", path)
for _ in range(4):
    ix.search_device(q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), 0, st)
torch.cuda.synchronize()
print("ok", ix.get_stat("shadow_scans"), ix.get_stat("tensor_batches"), ix.get_stat("fallback_queries"))
