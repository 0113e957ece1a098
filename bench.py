#!/usr/bin/env python
"""bench.py -- exact top-k search throughput of the B200 shard (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the CPU path, same metric/config)

A "step" is one search of one batch of synthetic queries over the whole (sharded) database.
Workload at N=1 = BASELINE config[1]: 1M x 512 fp32 unit-norm rows, cosine, top-10.  `value` is
the batched (batch 1024) queries/s with queries and results resident in HBM; the same line also
carries the single-query (batch 1) scan numbers under "single_query", because config[1] names
both.  N > 1 shards the SAME database by contiguous row ranges; every rank searches its shard for
the whole batch, the per-rank top-k lists are exchanged by query slice (all-to-all over
NCCL/NVLink) and rank r merges the lists of its slice on the GPU (the coordinator's
scatter-gather, src/coordinator/handler.py:191-216).  Default --scaling weak: the global batch is
1024 x N, so the contraction work per GPU (batch x rows/N) is fixed as N grows -- the shape of
BASELINE config[2] (10M rows over 8 GPUs, batch 4096); --scaling strong keeps the batch at 1024.

Prints ONE JSON line (rank 0).  Inputs are larger than L2 (2 GB shard vs 126 MB), so no flush.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_DB, SEED_QUERY = 0xD5B200, 0xC0FFEE


def measured_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            p = json.load(f)
        return {"hbm_gbs": float(p["hbm_gbs"]), "bf16_tflops": float(p["bf16_tflops"]),
                "bf16_tflops_sustained": float(p.get("bf16_tflops_sustained", p["bf16_tflops"])), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    """Samples SM clock / throttle reasons through NVML while the timed region runs."""

    def __init__(self, index: int):
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop = threading.Event()
        self._thr = None
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.nv = None

    def _run(self):
        nv = self.nv
        names = {
            getattr(nv, "nvmlClocksEventReasonHwSlowdown", 0x8): "hw_slowdown",
            getattr(nv, "nvmlClocksEventReasonHwThermalSlowdown", 0x40): "hw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwThermalSlowdown", 0x20): "sw_thermal_slowdown",
            getattr(nv, "nvmlClocksEventReasonSwPowerCap", 0x4): "sw_power_cap",
        }
        while not self._stop.is_set():
            try:
                self.samples.append(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
                try:
                    r = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h)
                except Exception:
                    r = nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                for bit, name in names.items():
                    if r & bit:
                        self.reasons.add(name)
            except Exception:
                pass
            self._stop.wait(0.01)

    def __enter__(self):
        if self.nv:
            self._thr = threading.Thread(target=self._run, daemon=True)
            self._thr.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        if self._thr:
            self._thr.join()

    def summary(self):
        return {"sm_mhz": float(np.median(self.samples)) if self.samples else None,
                "sm_max_mhz": self.max_mhz, "reasons": sorted(self.reasons), "samples": len(self.samples)}


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--rows", type=int, default=1_000_000, help="total database rows (all GPUs)")
    ap.add_argument("--dim", type=int, default=512)
    ap.add_argument("--k", type=int, default=10)
    ap.add_argument("--metric", default="cosine", choices=["l2", "ip", "cosine"])
    ap.add_argument("--store", default="f32", choices=["f32", "f16"])
    ap.add_argument("--batch", type=int, default=1024, help="queries per step (per GPU under --scaling weak)")
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"],
                    help="N>1: weak = global batch is batch x N (work per GPU fixed); strong = batch fixed")
    ap.add_argument("--single-steps", type=int, default=0, help="steps for the batch-1 leg (default: 5*steps)")
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="N>1: p2p = fused NVLink exchange+merge kernel (vdb_xchg_*), nccl = all-to-all + merge kernel")
    ap.add_argument("--no-single", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--cpu-queries", type=int, default=0)
    return ap.parse_args()


def global_batch(a, world):
    return a.batch * world if a.scaling == "weak" else (a.batch + world - 1) // world * world


def workload_name(a, world=1):
    return f"{a.rows}x{a.dim} {a.store} {a.metric} top-{a.k}, batch {global_batch(a, world)} (+ batch 1)"


# --------------------------------------------------------------------------------------------
# CPU legs (oracle port: the reference's own path -- hnswlib/plyvel/thrift -- is not installable)
# --------------------------------------------------------------------------------------------
_cpu_cache = {}


def cpu_knn_qps(a, nq: int, rows_cap: int = 1_000_000, return_ids: bool = False):
    """Times the oracle's exact scan (oracle/knn_ref.c, OpenMP, all host threads) for nq queries
    over min(rows, rows_cap) rows and scales linearly to the full row count.  The database is built once
    per process (outside the timed region)."""
    from oracle import c_ref
    n = min(a.rows, rows_cap)
    key = (n, a.dim, a.metric, a.store, nq)
    if key not in _cpu_cache:
        rows = c_ref.synth_rows(SEED_DB, 0, n, a.dim)
        stored = c_ref.normalize(rows) if a.metric == "cosine" else rows
        if a.store == "f16":
            stored = stored.astype(np.float16).astype(np.float32)
        q = c_ref.synth_rows(SEED_QUERY, 0, nq, a.dim)
        c_ref.knn(q[:1], stored[:1000], None, a.k, a.metric)        # warm the OpenMP pool
        _cpu_cache.clear()
        _cpu_cache[key] = (stored, q)
    stored, q = _cpu_cache[key]
    t0 = time.perf_counter()
    ids, _, _ = c_ref.knn(q, stored, None, a.k, a.metric)
    dt = time.perf_counter() - t0
    dt_full = dt * (a.rows / n)
    out = (nq / dt_full, c_ref.num_threads(), f"{nq} queries x {n} rows in {dt:.2f}s" + (
        f", scaled x{a.rows / n:.1f} to {a.rows} rows" if n != a.rows else ""))
    return out + ((ids if n == a.rows else None),) if return_ids else out


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    nq = a.cpu_queries or 256         # per step: ~1 s of CPU work on all host threads (x (warmup + steps) steps)
    vals, desc, cores = [], "", 1
    for i in range(a.warmup + a.steps):
        qps, cores, desc = cpu_knn_qps(a, nq)
        if i >= a.warmup:
            vals.append(qps)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "queries/sec exact top-k", "value": v, "unit": "queries/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * nq / v,
        "higher_is_better": True, "scaling": a.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": workload_name(a, world), "global_batch": global_batch(a, world), "rows_total": a.rows,
                   "inputs": "larger than L2/LLC (2 GB)",
                   "note": "CPU exact scan of the same database; throughput does not depend on the batch size"},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": "per step: " + desc + " (oracle/knn_ref.c exact scan, OpenMP)"},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
def run_ours(a):
    import torch
    import torch.distributed as dist
    import dvdb_b200 as vdb

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if a.gpus != world:
        if world == 1 and a.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (python -m torch.distributed.run --nproc-per-node N ...)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = vdb._ffi.lib()

    # ---- shard: contiguous row range of the same synthetic database --------------------------
    lo = a.rows * rank // world
    hi = a.rows * (rank + 1) // world
    ix = vdb.Index(a.metric, a.dim, store_dtype=a.store, device=local)
    ix.init_index(hi - lo)
    ix.add_synthetic(SEED_DB, lo, hi - lo, label_start=lo)
    elem = 2 if a.store == "f16" else 4
    ld = ix.get_stat("ld")
    shard_bytes = (hi - lo) * ld * elem

    stream = torch.cuda.current_stream().cuda_stream

    def make_queries(nq):
        q = torch.empty((nq, a.dim), dtype=torch.float32, device=dev)
        vdb._ffi.check(lib.vdb_synth_dev(SEED_QUERY, 0, nq, a.dim, q.data_ptr(), stream), "synth")
        return q

    # N > 1: the sharded client API (ShardedIndex: slice upload + NVLink all-gather of the queries, local search of
    # the whole batch, exchange by query slice + merge on the owner)
    sx = None
    if world > 1:
        sx = vdb.ShardedIndex(ix, max_batch=global_batch(a, world), max_k=a.k, exchange=a.exchange)

    def device_leg(nq, steps, warmup):
        """queries + results resident in HBM; returns (seconds for `steps`, dominant-kernel ns, launches)"""
        q = make_queries(nq)
        ids = torch.empty((nq, a.k), dtype=torch.int64, device=dev)
        dd = torch.empty((nq, a.k), dtype=torch.float32, device=dev)
        cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
        res = [None]

        def step():
            if sx is None:
                ix.search_device(q.data_ptr(), nq, a.k, ids.data_ptr(), dd.data_ptr(), cnt.data_ptr(), stream)
            else:
                res[0] = sx.search_device(q, a.k)           # (dist, ids) of this rank's slice

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ix.set_option("profile", 1)
        l0 = vdb.launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            step()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ix.set_option("profile", 0)
        launches = vdb.launch_count() - l0
        nprof = ix.get_stat("profile_count")
        kern_ns = ix.get_stat("profile_ns")
        ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        final = (ids if sx is None else res[0][1])[:4096].cpu().numpy()   # rank 0: the first queries of the batch (its slice when sharded)
        return float(ms.item()) * 1e-3, kern_ns, nprof, launches, final

    def e2e_leg(nq, steps, warmup):
        """through the public host-buffer API: H2D of the queries and D2H of the results inside the timed region"""
        qh = vdb.pinned_empty((nq, a.dim), np.float32)   # page-locked host buffers, filled outside the timed region
        qh[:] = make_queries(nq).cpu().numpy()
        if sx is None:
            outs = (vdb.pinned_empty((nq, a.k), np.int64), vdb.pinned_empty((nq, a.k), np.float32),
                    vdb.pinned_empty((nq,), np.int32))

            def step():
                return ix.knn_query_padded(qh, a.k, out=outs)
        else:
            # the batch arrives split over the ranks' hosts: each rank passes ITS slice and gets that slice's results;
            # a batch that does not divide by N (the single query) is passed whole by every rank
            even = nq % world == 0
            sl = nq // world
            pin_q = torch.from_numpy(qh[rank * sl:(rank + 1) * sl] if even else qh).pin_memory()
            out = [None]

            def step():
                out[0] = sx.search_host(pin_q, a.k, whole_batch=not even, out=out[0])
                return out[0]

        for _ in range(warmup):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        torch.cuda.synchronize()
        dt = torch.tensor([time.perf_counter() - t0], device=dev)
        if world > 1:
            dist.all_reduce(dt, op=dist.ReduceOp.MAX)
        return float(dt.item())

    peaks = measured_peaks()

    def cublas_tf32_tflops():
        """fp32 rows go through kind::tf32: the fair tensor denominator is the TF32 dense rate this box
        sustains, measured here the way MEASURED_PEAKS.json measures bf16 (torch.matmul 8192^3, best of 5)."""
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = True
        try:
            x = torch.randn(8192, 8192, device=dev)
            y = torch.randn(8192, 8192, device=dev)
            torch.matmul(x, y)
            best = 1e9
            for _ in range(5):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record(); torch.matmul(x, y); e1.record(); torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1))
            return 2 * 8192 ** 3 / (best * 1e-3) / 1e12
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old
            del x, y

    shadow = a.store == "f32" and ix.get_stat("shadow") == 1       # fp32 rows contracted through their fp16 plane
    tf32_path = a.store == "f32" and not shadow
    tf32_peak = cublas_tf32_tflops() if tf32_path else None
    dtype_name = ("f16 operands (shadow plane of f32 rows) / f32 accumulate, exact f32 re-rank" if shadow else
                  "tf32 operands / f32 accumulate, exact f32 re-rank" if tf32_path else
                  "f16-stored / f32 accumulate")
    traffic = {}
    tpath = os.path.join(ROOT, "profiles", "traffic.json")       # dram bytes per launch from the committed ncu captures
    if os.path.exists(tpath):
        with open(tpath) as f:
            traffic = json.load(f)
    with ClockSampler(local) as clk:
        B = global_batch(a, world)
        sec, kern_ns, nprof, launches, got_ids = device_leg(B, a.steps, a.warmup)
        e2e_sec = e2e_leg(B, a.steps, a.warmup)
        single = None
        if not a.no_single:
            ss = a.single_steps or max(5 * a.steps, 50)
            s_sec, s_kern_ns, s_nprof, s_launch, _ = device_leg(1, ss, a.warmup)
            s_e2e = e2e_leg(1, ss, a.warmup)
            t_k = s_kern_ns * 1e-9 / max(s_nprof, 1)
            ach = shard_bytes / t_k / 1e9
            single = {
                "value": ss / s_sec, "unit": "queries/s", "ms_per_query": 1e3 * s_sec / ss, "steps": ss,
                "e2e": {"value": ss / s_e2e, "unit": "queries/s", "h2d_bytes_per_step": a.dim * 4,
                        "d2h_bytes_per_step": a.k * 12 + 4},
                "roofline": {"bound": "hbm", "kernel": "scan_topk_kernel", "achieved": ach, "peak": peaks["hbm_gbs"],
                             "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "peak_source": peaks["source"],
                             "algorithmic_bytes_per_launch": shard_bytes, "kernel_us": t_k * 1e6,
                             "traffic": traffic.get(f"scan_topk_kernel|{a.rows}x{a.dim} {a.store}") if world == 1 else None},
                "gpu_launches": s_launch,
            }

    # roofline of the batched leg's dominant kernel
    passes_per_step = nprof / max(a.steps, 1)
    t_kernel = kern_ns * 1e-9 / max(nprof, 1)
    tensor_batches = ix.get_stat("tensor_batches")
    if tensor_batches > 0:
        # one search = `passes_per_step` launches of gemm_filter_kernel (one per threshold level) that together
        # contract every query with every row once: algorithmic flops per search / summed launch time
        flops = 2.0 * B * (hi - lo) * a.dim          # this rank's share: the whole batch against its rows
        t_step = kern_ns * 1e-9 / max(a.steps, 1)
        ach = flops / t_step / 1e12
        # kind::f16 against the measured bf16 burst peak (same tensor rate); kind::tf32 runs at half that rate
        peak = peaks["bf16_tflops"] * (0.5 if tf32_path else 1.0)
        roof = {"bound": "tensor", "kernel": "gemm_filter_kernel", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak,
                "peak_source": peaks["source"] + (" bf16 burst x 0.5 (kind::tf32 runs at half the bf16 rate)" if tf32_path
                                                  else " bf16 burst (kind::f16, same tensor rate)"),
                "frac_of_measured_bf16_sustained": ach / (peaks["bf16_tflops_sustained"] * (0.5 if tf32_path else 1.0)),
                "algorithmic_flops_per_step": flops, "kernel_us_per_step": t_step * 1e6,
                "launches_per_step": passes_per_step,
                "traffic": traffic.get(f"gemm_filter_kernel|{workload_name(a, world)}")}
        if tf32_peak:
            roof["cublas_tf32_8192_tflops_this_run"] = tf32_peak
    else:
        ach = shard_bytes / t_kernel / 1e9
        roof = {"bound": "hbm", "kernel": "scan_topk_kernel", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": ach / peaks["hbm_gbs"], "peak_source": peaks["source"],
                "algorithmic_bytes_per_launch": shard_bytes, "kernel_us": t_kernel * 1e6,
                "launches_per_step": passes_per_step,
                "traffic": traffic.get(f"scan_topk_kernel|{a.rows}x{a.dim} {a.store}")}

    if rank == 0:
        cpu = None
        recall = None
        if not a.no_cpu and world == 1:
            # bounded sample of the same workload, sized from a short (cold) probe to 10-20 s of CPU work on all host threads
            nq_cpu = a.cpu_queries
            if not nq_cpu:
                probe_qps = cpu_knn_qps(a, 64)[0]
                nq_cpu = int(min(B * 8, max(64, 64 * round(20.0 * probe_qps / 64))))
            qps, cores, desc, want_ids = cpu_knn_qps(a, nq_cpu, return_ids=True)
            cpu = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                   "sample": desc + " (oracle/knn_ref.c exact scan, OpenMP)"}
            if want_ids is not None and a.rows <= 1_000_000:
                m = min(len(want_ids), len(got_ids))
                hits = sum(len(set(got_ids[i].tolist()) & set(want_ids[i].tolist())) for i in range(m))
                recall = hits / float(m * a.k)
        line = {
            "metric": "queries/sec exact top-k", "value": B * a.steps / sec, "unit": "queries/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * sec / a.steps,
            "higher_is_better": True, "scaling": a.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": dtype_name, "data": "synthetic",
            "config": {"workload": workload_name(a, world), "global_batch": B, "rows_total": a.rows,
                       "rows_per_gpu": hi - lo, "sharding": f"contiguous rows x{world}",
                       "exchange": "none" if world == 1 else (
                           "fused NVLink peer-store exchange + merge kernel, by query slice" if a.exchange == "p2p"
                           else "all-to-all by query slice (NCCL) + GPU merge"),
                       "l2": "inputs larger than L2 (no flush needed)", "path": "tensor" if tensor_batches > 0 else "scan"},
            "e2e": {"value": B * a.steps / e2e_sec, "unit": "queries/s",
                    "h2d_bytes_per_step": B * a.dim * 4,
                    "d2h_bytes_per_step": B * a.k * 12 + (B * 4 if world == 1 else 0)},
            "gpu_launches": launches,
            "roofline": roof,
            "cpu_baseline": cpu,
            "single_query": single,
            "clocks": clk.summary(),
            "recall_at_k": recall,
            "fallback_queries": ix.get_stat("fallback_queries"),
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
