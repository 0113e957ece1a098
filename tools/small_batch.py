"""Search time against batch size: forced fp32 scan (passes of <= 8 queries), forced tensor path, and what the
library picks itself (one or two queries: scan of the fp16 shadow plane + exact re-rank), 1M x 512."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import dvdb_b200 as vdb
rows = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
store = sys.argv[2] if len(sys.argv) > 2 else "f32"
ix = vdb.Index("cosine", 512, store_dtype=store); ix.init_index(rows); ix.add_synthetic(0xD5B200, 0, rows)
dev = torch.device("cuda", 0); st = torch.cuda.current_stream().cuda_stream
lib = vdb._ffi.lib()
for nq in (1, 2, 3, 4, 5, 6, 8, 12, 16, 32, 64, 128, 256):
    q = torch.empty((nq, 512), dtype=torch.float32, device=dev)
    vdb._ffi.check(lib.vdb_synth_dev(0xC0FFEE, 0, nq, 512, q.data_ptr(), st), "synth")
    ids = torch.empty((nq, 10), dtype=torch.int64, device=dev); dd = torch.empty((nq, 10), dtype=torch.float32, device=dev)
    res = {}
    for path in (1, 2, 0):
        ix.set_option("path", path)
        for _ in range(3):
            ix.search_device(q.data_ptr(), nq, 10, ids.data_ptr(), dd.data_ptr(), 0, st)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n = 20 if (path == 2 or nq <= 16) else 5
        e0.record()
        for _ in range(n):
            ix.search_device(q.data_ptr(), nq, 10, ids.data_ptr(), dd.data_ptr(), 0, st)
        e1.record(); torch.cuda.synchronize()
        res[path] = e0.elapsed_time(e1) / n * 1e3
    print(f"nq={nq:4d}  scan {res[1]:9.1f} us   tensor {res[2]:9.1f} us   auto {res[0]:9.1f} us (shadow scans so far {ix.get_stat('shadow_scans')})", flush=True)
