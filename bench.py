#!/usr/bin/env python
"""bench.py -- exact top-k search throughput of the B200 shard (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched under torchrun)
    python bench.py --impl reference ...                     (the CPU path, same metric/config)

A "step" is one search of one batch of synthetic queries over the whole (sharded) database.

Headline workload = BASELINE config[1]: 1M x 512 fp32 unit-norm rows, cosine, top-10.  `value` is the
batched (batch 1024) queries/s with queries and results resident in HBM; the same line carries the
single-query (batch 1) scan numbers under "single_query", because config[1] names both.  N > 1 shards the
SAME database by contiguous row ranges; every rank searches its shard for the whole batch, the per-rank
top-k lists are exchanged by query slice and rank r merges the lists of its slice on the GPU (the
coordinator's scatter-gather, src/coordinator/handler.py:191-216).  Default --scaling weak: the global
batch is 1024 x N, so the contraction work per GPU (batch x rows/N) is fixed as N grows.

Beside the headline the same line holds
  "sustained"  the headline step repeated back to back for >= 2 s (power-capped steady state), against the
               sustained tensor peak, with its own clock samples
  "two_streams" (N=1) the same step with two batches in flight on two streams, as two server threads produce them
  "check"      results of sampled queries against the CPU oracle over the WHOLE database at every N (rows
               regenerated from the counter RNG), ordered ids + distances, and at N > 1 one step of the fused
               NVLink exchange against the NCCL exchange (bitwise) -- all outside the timed regions
  "configs"    at --gpus 8 (or --configs 3,4): BASELINE config[2] (10M x 512 fp32 L2 top-100, batch 4096) and
               config[3] (100M x 768 fp16 inner-product top-10) at FULL size, 1.25M / 12.5M rows per GPU, each
               with its own value / e2e / roofline / single-query scan / oracle check

Prints ONE JSON line (rank 0).  Inputs are larger than L2 (>= 2 GB per shard vs 126 MB), so no flush.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import threading
import time
from dataclasses import dataclass

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SEED_DB, SEED_QUERY = 0xD5B200, 0xC0FFEE
RTOL = 1e-5                      # BASELINE.json: distance ties within 1e-5 relative for fp32


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
    ap.add_argument("--no-shadow-scan", action="store_true",
                    help="single queries scan the fp32 rows (K1 over the stored rows) instead of the fp16 shadow plane + re-rank")
    ap.add_argument("--no-cpu", action="store_true", help="skip the CPU baseline leg (N=1) -- the oracle check stays")
    ap.add_argument("--no-check", action="store_true", help="skip the oracle / exchange cross-checks")
    ap.add_argument("--cpu-queries", type=int, default=0)
    ap.add_argument("--sustain-s", type=float, default=2.0, help="seconds of back-to-back steps (0 = skip)")
    ap.add_argument("--check-queries", type=int, default=64, help="sampled queries checked against the CPU oracle")
    ap.add_argument("--oracle-budget-s", type=float, default=240.0,
                    help="CPU-oracle time allowed per workload; beyond it the check falls back to the GPU cross-check "
                         "+ exact CPU distances of the returned rows")
    ap.add_argument("--hnsw-rows", type=int, default=100_000,
                    help="--impl reference: rows of the HNSW-port leg (oracle/hnsw_ref.c; 0 = skip it)")
    ap.add_argument("--configs", default="auto",
                    help="'auto' (configs 3 and 4 at --gpus 8), 'none', or a list like '3,4' (rows per GPU as at 8 GPUs)")
    return ap.parse_args()


def global_batch(a, world):
    return a.batch * world if a.scaling == "weak" else (a.batch + world - 1) // world * world


@dataclass
class Workload:
    key: str
    rows: int          # total rows, all GPUs
    dim: int
    k: int
    metric: str
    store: str
    batch: int         # global batch

    def name(self):
        return f"{self.rows}x{self.dim} {self.store} {self.metric} top-{self.k}, batch {self.batch} (+ batch 1)"


def headline_workload(a, world):
    return Workload("config2", a.rows, a.dim, a.k, a.metric, a.store, global_batch(a, world))


def extra_workloads(a, world):
    want = a.configs
    if want == "auto":
        want = "3,4" if world == 8 else "none"
    if want == "none":
        return []
    out = []
    for c in want.split(","):
        c = c.strip()
        if c == "3":    # BASELINE config[2]: 10M x 512 fp32 L2 top-100 over 8 GPUs, batch 4096
            out.append(Workload("config3", 1_250_000 * world, 512, 100, "l2", "f32", 4096))
        elif c == "4":  # BASELINE config[3]: 100M x 768 fp16-stored inner product top-10 over 8 GPUs
            out.append(Workload("config4", 12_500_000 * world, 768, 10, "ip", "f16", 1024))
        elif c:
            raise SystemExit(f"--configs: unknown config {c!r}")
    return out


# --------------------------------------------------------------------------------------------
# CPU side (oracle port: the reference's own stack -- hnswlib/plyvel/thrift -- is not installable)
# --------------------------------------------------------------------------------------------
_cpu_cache = {}


def cpu_setup():
    """All host cores for every OpenMP region of the oracle, whatever OMP_NUM_THREADS the launcher exported
    (torchrun sets it to 1)."""
    from oracle import c_ref
    cores = c_ref.host_cores()
    c_ref.set_threads(cores)
    return c_ref, cores


def stored_rows(c_ref, wl: Workload, row0: int, n: int) -> np.ndarray:
    """Rows [row0, row0+n) of the synthetic database as the shard stores them (fp32 values)."""
    rows = c_ref.synth_rows(SEED_DB, row0, n, wl.dim)
    if wl.metric == "cosine":
        rows = c_ref.normalize(rows)
    if wl.store == "f16":
        c_ref.round_f16_(rows)
    return rows


def cpu_knn_qps(wl: Workload, nq: int, rows_cap: int = 1_000_000, return_results: bool = False):
    """Times the oracle's exact scan (oracle/knn_ref.c, OpenMP, all host cores) for nq queries over
    min(rows, rows_cap) rows and scales linearly to the full row count.  The database is built once per
    process (outside the timed region)."""
    c_ref, cores = cpu_setup()
    n = min(wl.rows, rows_cap)
    key = (n, wl.dim, wl.metric, wl.store, nq)
    if key not in _cpu_cache:
        stored = stored_rows(c_ref, wl, 0, n)
        q = c_ref.synth_rows(SEED_QUERY, 0, nq, wl.dim)
        c_ref.knn(q[:1], stored[:1000], None, wl.k, wl.metric, nthreads=cores)        # warm the OpenMP pool
        _cpu_cache.clear()
        _cpu_cache[key] = (stored, q)
    stored, q = _cpu_cache[key]
    t0 = time.perf_counter()
    ids, dd, _ = c_ref.knn(q, stored, None, wl.k, wl.metric, nthreads=cores)
    dt = time.perf_counter() - t0
    dt_full = dt * (wl.rows / n)
    out = (nq / dt_full, cores, f"{nq} queries x {n} rows in {dt:.2f}s" + (
        f", scaled x{wl.rows / n:.1f} to {wl.rows} rows" if n != wl.rows else ""))
    return out + (((ids, dd) if n == wl.rows else None),) if return_results else out


def oracle_topk(wl: Workload, q_idx: np.ndarray, budget_s: float, chunk: int = 500_000):
    """Exact top-k of the sampled queries over the WHOLE database on the CPU, chunk by chunk (rows regenerated from
    the counter RNG, never held all at once).  Returns (ids, dist, info) -- ids None when the projected time
    exceeds the budget."""
    c_ref, cores = cpu_setup()
    q = np.concatenate([c_ref.synth_rows(SEED_QUERY, int(i), 1, wl.dim) for i in q_idx])
    nq, k = len(q_idx), wl.k
    best_i = np.full((nq, 0), -1, dtype=np.int64)
    best_d = np.full((nq, 0), np.inf, dtype=np.float32)
    t0 = time.perf_counter()
    done = 0
    for r0 in range(0, wl.rows, chunk):
        n = min(chunk, wl.rows - r0)
        rows = stored_rows(c_ref, wl, r0, n)
        ids, dd, _ = c_ref.knn(q, rows, np.arange(r0, r0 + n, dtype=np.int64), min(k, n), wl.metric, nthreads=cores)
        ci = np.concatenate([best_i, ids], axis=1)
        cd = np.concatenate([best_d, dd], axis=1)
        cd_key = np.where(ci < 0, np.inf, cd)
        order = np.lexsort((ci, cd_key), axis=1)[:, :k]          # ascending (distance, id)
        best_i = np.take_along_axis(ci, order, axis=1)
        best_d = np.take_along_axis(cd, order, axis=1)
        done = r0 + n
        el = time.perf_counter() - t0
        if done < wl.rows and el * wl.rows / done > budget_s and el > 5.0:
            return None, None, {"rows_scanned": done, "seconds": el, "cores": cores,
                                "note": f"projected {el * wl.rows / done:.0f}s > budget {budget_s:.0f}s"}
    return best_i, best_d, {"rows_scanned": done, "seconds": time.perf_counter() - t0, "cores": cores}


def oracle_distances_of(wl: Workload, q_idx: np.ndarray, ids: np.ndarray) -> np.ndarray:
    """Exact oracle distances of the returned (query, id) pairs only (rows regenerated one by one)."""
    c_ref, _ = cpu_setup()
    out = np.full(ids.shape, np.inf, dtype=np.float32)
    for r, qi in enumerate(q_idx):
        q = c_ref.synth_rows(SEED_QUERY, int(qi), 1, wl.dim)
        for j, rid in enumerate(ids[r]):
            if rid >= 0:
                out[r, j] = c_ref.distances(q[0], stored_rows(c_ref, wl, int(rid), 1), wl.metric)[0]
    return out


def compare_topk(got_i, got_d, want_i, want_d):
    """Parity figures of got vs want ([nq,k] each): recall (set overlap), fraction of rows whose id lists are
    identical in order, worst relative distance error position by position, and the verdict of the parity bar
    (ids identical; rows whose distances lie within RTOL relative may swap)."""
    nq, k = want_i.shape
    hits = sum(len(set(got_i[r].tolist()) & set(want_i[r].tolist())) for r in range(nq))
    same = (got_i == want_i)
    scale = np.maximum(1.0, np.abs(want_d.astype(np.float64)))
    fin = np.isfinite(want_d) & np.isfinite(got_d)
    rel = np.where(fin, np.abs(got_d.astype(np.float64) - want_d.astype(np.float64)) / scale, 0.0)
    rel = np.where(np.isfinite(want_d) != np.isfinite(got_d), np.inf, rel)
    # an id mismatch is tolerated only inside a group of distances tied within RTOL (neighbouring ranks swapped)
    bad, tie_misses = 0, 0
    for r in range(nq):        # ids the other side does not have at all: fine only as a swap at the k-th place between tied distances
        miss = set(want_i[r].tolist()) - set(got_i[r].tolist())
        if miss:
            dk = float(want_d[r, k - 1])
            tie_misses += sum(1 for t in range(k) if int(want_i[r, t]) in miss and abs(float(want_d[r, t]) - dk) <= RTOL * max(1.0, abs(dk))
                              and abs(float(got_d[r, k - 1]) - dk) <= RTOL * max(1.0, abs(dk)))
    for r, j in zip(*np.nonzero(~same)):
        d = float(want_d[r, j])
        lo, hi = d - RTOL * max(1.0, abs(d)), d + RTOL * max(1.0, abs(d))
        tied = {int(want_i[r, t]) for t in range(k) if lo <= float(want_d[r, t]) <= hi}
        if int(got_i[r, j]) not in tied and not (j == k - 1 and abs(float(got_d[r, j]) - d) <= RTOL * max(1.0, abs(d))):
            bad += 1
    return {"queries": int(nq), "recall_at_k": hits / float(nq * k),
            "recall_at_k_ties_within_rtol": (hits + tie_misses) / float(nq * k),     # a k-th place tie may go either way
            "ordered_ids_equal": float(same.all(axis=1).mean()),
            "max_rel_dist_err": float(rel.max()) if rel.size else 0.0,
            "parity_ok": bool(bad == 0 and (rel.max() if rel.size else 0.0) <= RTOL)}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    wl = headline_workload(a, world)
    nq = a.cpu_queries or 256         # per step: ~0.5 s of CPU work on all host cores (x (warmup + steps) steps)
    vals, desc, cores = [], "", 1
    for i in range(a.warmup + a.steps):
        qps, cores, desc = cpu_knn_qps(wl, nq)
        if i >= a.warmup:
            vals.append(qps)
    v = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "queries/sec exact top-k", "value": v, "unit": "queries/s",
        "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup, "ms_per_step": 1e3 * nq / v,
        "higher_is_better": True, "scaling": a.scaling if world > 1 else "weak", "vs_baseline": None, "dtype": "f32",
        "data": "synthetic",
        "config": {"workload": wl.name(), "global_batch": wl.batch, "rows_total": wl.rows,
                   "inputs": "larger than L2/LLC (2 GB)",
                   "note": "CPU EXACT brute-force scan of the same database (the oracle's OpenMP port of hnswlib's "
                           "distance arithmetic): not the reference's approximate HNSW walk, whose dependency "
                           "(hnswlib) is not installable here; throughput does not depend on the batch size"},
        "cpu_baseline": {"value": v, "unit": "queries/s", "cores": cores, "kind": "port",
                         "sample": "per step: " + desc + " (oracle/knn_ref.c exact scan, OpenMP)"},
        "e2e": {"value": v, "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    hn = hnswlib_leg(wl)
    if hn is not None:
        line["hnswlib"] = hn
    if a.hnsw_rows > 0:
        line["hnsw_port"] = hnsw_port_leg(wl, rows_cap=a.hnsw_rows)
    print(json.dumps(line), flush=True)


def hnsw_port_leg(wl: Workload, rows_cap: int = 100_000, nq: int = 2048):
    """What the reference's APPROXIMATE walk does per second, and at what recall: oracle/hnsw_ref.c, a restatement of
    hnswlib v0.8.0's HNSW (the library itself is not installable here) with the reference's parameters -- M=32,
    ef_construction=128 (src/datanode/handler.py:86), ef=max(50, 2k) and 2k results asked for (:360-364) -- on a bounded
    prefix of the same database, all host cores.  Informative only: `value` stays the exact scan (the GPU path is exact,
    recall 1.0), and the synthetic rows are structureless (uniform), the hardest case for a graph index."""
    try:
        from oracle import hnsw_port
        c_ref, cores = cpu_setup()
        n = min(wl.rows, rows_cap)
        raw = c_ref.synth_rows(SEED_DB, 0, n, wl.dim)
        stored = c_ref.normalize(raw) if wl.metric == "cosine" else raw
        q = c_ref.synth_rows(SEED_QUERY, 0, nq, wl.dim)
        qn = c_ref.normalize(q) if wl.metric == "cosine" else q
        t0 = time.perf_counter()
        h = hnsw_port.HnswPort(stored, wl.metric, M=32, ef_construction=128, nthreads=cores)
        build_s = time.perf_counter() - t0
        k2, ef = 2 * wl.k, max(50, 2 * wl.k)                        # the handler asks for 2k and filters (:364,375-402)
        h.knn_query(qn[:64], k2, ef, cores)
        t0 = time.perf_counter()
        hl, _ = h.knn_query(qn, k2, ef, cores)
        dt = time.perf_counter() - t0
        el, _, _ = c_ref.knn(q, stored, None, wl.k, wl.metric, nthreads=cores)
        rec = float(np.mean([len(set(hl[i, :wl.k].tolist()) & set(el[i].tolist())) / wl.k for i in range(nq)]))
        out = {"kind": "port of the HNSW algorithm of hnswlib v0.8.0 (oracle/hnsw_ref.c), not the library",
               "rows": n, "queries": nq, "M": 32, "ef_construction": 128, "ef": ef, "k": wl.k, "qps": nq / dt,
               "recall_at_k_vs_exact": rec, "build_s": build_s, "cores": cores, "max_level": h.max_level,
               "note": "approximate search over the first %d rows of the same synthetic (uniform, structureless) database; "
                       "the exact arms scan all %d rows" % (n, wl.rows)}
        # what ef the same graph needs for a usable recall on these rows (search only; the reference fixes ef = max(50, 2k))
        sweep = []
        for ef2 in (200, 800):
            m = min(nq, 512)
            t0 = time.perf_counter()
            hl2, _ = h.knn_query(qn[:m], k2, ef2, cores)
            dt2 = time.perf_counter() - t0
            sweep.append({"ef": ef2, "qps": m / dt2,
                          "recall_at_k_vs_exact": float(np.mean([len(set(hl2[i, :wl.k].tolist()) & set(el[i].tolist())) / wl.k
                                                                 for i in range(m)]))})
        out["ef_sweep"] = sweep
        h.close()
        # the same index parameters on CLUSTERED rows (what embeddings look like): 200 Gaussian clusters, sigma 0.3
        rng = np.random.default_rng(7)
        nc = min(n, 50_000)
        cent = rng.normal(size=(200, wl.dim)).astype(np.float32)
        rows_c = (cent[rng.integers(0, 200, nc)] + 0.3 * rng.normal(size=(nc, wl.dim))).astype(np.float32)
        q_c = (cent[rng.integers(0, 200, nq)] + 0.3 * rng.normal(size=(nq, wl.dim))).astype(np.float32)
        stored_c = c_ref.normalize(rows_c) if wl.metric == "cosine" else rows_c
        qn_c = c_ref.normalize(q_c) if wl.metric == "cosine" else q_c
        h = hnsw_port.HnswPort(stored_c, wl.metric, M=32, ef_construction=128, nthreads=cores)
        h.knn_query(qn_c[:64], k2, ef, cores)
        t0 = time.perf_counter()
        hl, _ = h.knn_query(qn_c, k2, ef, cores)
        dt = time.perf_counter() - t0
        el, _, _ = c_ref.knn(q_c, stored_c, None, wl.k, wl.metric, nthreads=cores)
        out["clustered"] = {"rows": nc, "qps": nq / dt,
                            "recall_at_k_vs_exact": float(np.mean([len(set(hl[i, :wl.k].tolist()) & set(el[i].tolist())) / wl.k
                                                                   for i in range(nq)]))}
        h.close()
        return out
    except Exception as e:           # never lose the line to an optional leg
        return {"error": repr(e)}


def hnswlib_leg(wl: Workload, rows_cap: int = 100_000, nq: int = 256):
    """When hnswlib is importable (it is not in the build image; the GPU box may differ): the reference's real
    index -- hnswlib.Index(M=32, ef_construction=128, ef=max(50, 2k)), src/datanode/handler.py:86,360-364 -- timed on
    a bounded subset with its recall against hnswlib.BFIndex, and BFIndex against the oracle bit for bit."""
    try:
        import hnswlib
    except Exception:
        return None
    try:
        c_ref, cores = cpu_setup()
        n = min(wl.rows, rows_cap)
        raw = c_ref.synth_rows(SEED_DB, 0, n, wl.dim)
        q = c_ref.synth_rows(SEED_QUERY, 0, nq, wl.dim)
        bf = hnswlib.BFIndex(space=wl.metric, dim=wl.dim)
        bf.init_index(max_elements=n)
        bf.add_items(raw, np.arange(n))
        bl, bd = bf.knn_query(q, k=wl.k)
        stored = c_ref.normalize(raw) if wl.metric == "cosine" else raw
        ol, od, _ = c_ref.knn(q, stored, None, wl.k, wl.metric, nthreads=cores)
        ix = hnswlib.Index(space=wl.metric, dim=wl.dim)
        ix.init_index(max_elements=n, ef_construction=128, M=32)
        t0 = time.perf_counter()
        ix.add_items(raw, np.arange(n))
        build_s = time.perf_counter() - t0
        ix.set_ef(max(50, 2 * wl.k))
        t0 = time.perf_counter()
        hl, _ = ix.knn_query(q, k=wl.k)
        dt = time.perf_counter() - t0
        rec = np.mean([len(set(hl[i].tolist()) & set(bl[i].tolist())) / wl.k for i in range(nq)])
        return {"rows": n, "queries": nq, "hnsw_qps": nq / dt, "hnsw_recall_vs_bf": float(rec), "hnsw_build_s": build_s,
                "bf_equals_oracle_ids": bool(np.array_equal(bl.astype(np.int64), ol)),
                "bf_equals_oracle_dist_bits": bool(np.array_equal(bd.view(np.uint32), od.view(np.uint32))),
                "version": getattr(hnswlib, "__version__", "?")}
    except Exception as e:           # never lose the line to an optional leg
        return {"error": repr(e)}


# --------------------------------------------------------------------------------------------
# GPU arm
# --------------------------------------------------------------------------------------------
class Arm:
    def __init__(self, a):
        import torch
        import torch.distributed as dist
        import dvdb_b200 as vdb
        self.a, self.torch, self.dist, self.vdb = a, torch, dist, vdb
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if a.gpus != self.world and self.world == 1 and a.gpus > 1:
            raise SystemExit("launch N>1 with torchrun (python -m torch.distributed.run --nproc-per-node N ...)")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.lib = vdb._ffi.lib()
        self.stream = torch.cuda.current_stream().cuda_stream
        self.peaks = measured_peaks()
        self.traffic = {}
        tpath = os.path.join(ROOT, "profiles", "traffic.json")   # dram bytes per launch from the committed ncu captures
        if os.path.exists(tpath):
            with open(tpath) as f:
                self.traffic = json.load(f)

    def barrier(self):
        self.torch.cuda.synchronize()
        if self.world > 1:
            self.dist.barrier()
            self.torch.cuda.synchronize()

    def make_queries(self, wl, nq, start=0):
        q = self.torch.empty((nq, wl.dim), dtype=self.torch.float32, device=self.dev)
        self.vdb._ffi.check(self.lib.vdb_synth_dev(SEED_QUERY, start, nq, wl.dim, q.data_ptr(), self.stream), "synth")
        return q

    # ---- one workload ---------------------------------------------------------------------
    def measure(self, wl: Workload, steps: int, warmup: int, *, headline: bool):
        torch, dist, vdb, a = self.torch, self.dist, self.vdb, self.a
        world, rank, dev, stream = self.world, self.rank, self.dev, self.stream
        lo, hi = wl.rows * rank // world, wl.rows * (rank + 1) // world
        ix = vdb.Index(wl.metric, wl.dim, store_dtype=wl.store, device=self.local)
        ix.init_index(hi - lo)
        ix.add_synthetic(SEED_DB, lo, hi - lo, label_start=lo)
        if a.no_shadow_scan:
            ix.set_option("shadow_scan_nq", 0)
        elem = 2 if wl.store == "f16" else 4
        ld = ix.get_stat("ld")
        shard_bytes = (hi - lo) * ld * elem
        sx = vdb.ShardedIndex(ix, max_batch=wl.batch, max_k=wl.k, exchange=a.exchange) if world > 1 else None
        B, k = wl.batch, wl.k

        def device_leg(nq, nsteps, nwarm, seconds=0.0):
            """queries + results resident in HBM.  seconds > 0: repeat blocks of nsteps until that much time passed."""
            q = self.make_queries(wl, nq)
            ids = torch.empty((nq, k), dtype=torch.int64, device=dev)
            dd = torch.empty((nq, k), dtype=torch.float32, device=dev)
            cnt = torch.empty((nq,), dtype=torch.int32, device=dev)
            res = [None]

            def step():
                if sx is None:
                    ix.search_device(q.data_ptr(), nq, k, ids.data_ptr(), dd.data_ptr(), cnt.data_ptr(), stream)
                else:
                    res[0] = sx.search_device(q, k)           # (dist, ids) of this rank's slice

            for _ in range(nwarm):
                step()
            self.barrier()
            ix.set_option("profile", 1 if seconds == 0.0 else 0)
            l0 = vdb.launch_count()
            total_ms, total_steps = 0.0, 0
            while True:
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(nsteps):
                    step()
                e1.record()
                torch.cuda.synchronize()
                total_ms += e0.elapsed_time(e1)
                total_steps += nsteps
                if total_ms >= seconds * 1e3:
                    break
            self.barrier()
            ix.set_option("profile", 0)
            launches = vdb.launch_count() - l0
            nprof = ix.get_stat("profile_count")
            kern_ns = ix.get_stat("profile_ns")
            ms = torch.tensor([total_ms], device=dev)
            if world > 1:
                dist.all_reduce(ms, op=dist.ReduceOp.MAX)
            if sx is None:
                got = (ids.cpu().numpy(), dd.cpu().numpy())
            else:
                got = (res[0][1].cpu().numpy(), res[0][0].cpu().numpy())      # this rank's slice
            return {"sec": float(ms.item()) * 1e-3, "steps": total_steps, "kern_ns": kern_ns, "nprof": nprof,
                    "launches": launches, "ids": got[0], "dist": got[1]}

        def e2e_leg(nq, nsteps, nwarm, in_flight=1):
            """through the public host-buffer API: H2D of the queries and D2H of the results inside the timed region"""
            qh = vdb.pinned_empty((nq, wl.dim), np.float32)   # page-locked host buffers, filled outside the timed region
            qh[:] = self.make_queries(wl, nq).cpu().numpy()
            if sx is None and in_flight <= 1:
                outs = (vdb.pinned_empty((nq, k), np.int64), vdb.pinned_empty((nq, k), np.float32),
                        vdb.pinned_empty((nq,), np.int32))

                def step():
                    return ix.knn_query_padded(qh, k, out=outs)

                def drain():
                    pass
            elif sx is None:
                # the serving steady state of one GPU: two batches in flight through vdb_search_submit / _collect --
                # the upload of batch i+1 is behind the kernels of batch i; every step still uploads its queries
                # from page-locked host memory and has its results in host memory when it is collected
                outs2 = [(vdb.pinned_empty((nq, k), np.int64), vdb.pinned_empty((nq, k), np.float32),
                          vdb.pinned_empty((nq,), np.int32)) for _ in range(in_flight)]
                pending, n_sub, out = [], [0], [None]

                def step():
                    pending.append(ix.submit_query(qh, k, out=outs2[n_sub[0] % in_flight]))
                    n_sub[0] += 1
                    if len(pending) == in_flight:
                        out[0] = ix.collect_query(pending.pop(0))
                    return out[0]

                def drain():
                    while pending:
                        out[0] = ix.collect_query(pending.pop(0))
            else:
                # the batch arrives split over the ranks' hosts: each rank passes ITS slice and gets that slice's
                # results; a batch that does not divide by N (the single query) is passed whole by every rank
                even = nq % world == 0
                sl = nq // world
                pin_q = torch.from_numpy(qh[rank * sl:(rank + 1) * sl] if even else qh).pin_memory()
                out = [None]
                if even:
                    # the serving steady state: two batches in flight (ShardedIndex.submit_host / collect) -- the upload
                    # + all-gather of batch i+1 overlap the search of batch i; every step still uploads its queries
                    # from page-locked host memory and reads its results back
                    pending = []

                    def step():
                        pending.append(sx.submit_host(pin_q, k))
                        if len(pending) == 2:
                            out[0] = sx.collect(pending.pop(0))
                        return out[0]

                    def drain():
                        while pending:
                            out[0] = sx.collect(pending.pop(0))
                else:
                    def step():
                        out[0] = sx.search_host(pin_q, k, whole_batch=True, out=out[0])
                        return out[0]

                    def drain():
                        pass

            for _ in range(nwarm):
                step()
            drain()
            self.barrier()
            t0 = time.perf_counter()
            for _ in range(nsteps):
                step()
            drain()                                         # the last results are on the host before the clock stops
            torch.cuda.synchronize()
            dt = torch.tensor([time.perf_counter() - t0], device=dev)
            if world > 1:
                dist.all_reduce(dt, op=dist.ReduceOp.MAX)
            return float(dt.item())

        shadow = wl.store == "f32" and ix.get_stat("shadow") == 1     # fp32 rows contracted through their fp16 plane
        tf32_path = wl.store == "f32" and not shadow
        peaks = self.peaks
        tkey = f"{hi - lo}x{wl.dim} {wl.store} {wl.metric} top-{k}"

        with ClockSampler(self.local) as clk:
            main = device_leg(B, steps, warmup)
            e2e_sec = e2e_leg(B, steps, warmup, in_flight=2)
            e2e_serial_sec = e2e_leg(B, steps, warmup) if world == 1 else None
            single = None
            if not a.no_single:
                ss = a.single_steps or max(5 * steps, 50)
                sh0 = ix.get_stat("shadow_scans")
                s = device_leg(1, ss, warmup)
                s_e2e = e2e_leg(1, ss, warmup)
                t_k = s["kern_ns"] * 1e-9 / max(s["nprof"], 1)
                # an fp32 shard with an fp16 shadow plane: the single query streams the SHADOW (half the bytes) and the
                # few candidates that can still matter are recomputed from the fp32 rows -- the algorithmic bytes of
                # the scan launch are the plane it reads
                shadow_scan = ix.get_stat("shadow_scans") > sh0
                fp32_bytes = shard_bytes
                if shadow_scan:
                    shard_bytes = (hi - lo) * ix.get_stat("ld16") * 2
                ach = shard_bytes / t_k / 1e9
                single = {
                    "value": ss / s["sec"], "unit": "queries/s", "ms_per_query": 1e3 * s["sec"] / ss, "steps": ss,
                    "e2e": {"value": ss / s_e2e, "unit": "queries/s", "h2d_bytes_per_step": wl.dim * 4,
                            "d2h_bytes_per_step": k * 12 + 4},
                    "roofline": {"bound": "hbm", "kernel": "scan_topk_kernel", "achieved": ach, "peak": peaks["hbm_gbs"],
                                 "unit": "GB/s", "frac": ach / peaks["hbm_gbs"], "frac_of_nominal_8tbs": ach / 8000.0,
                                 "peak_source": peaks["source"], "algorithmic_bytes_per_launch": shard_bytes,
                                 "kernel_us": t_k * 1e6,
                                 "api_level_gbs": shard_bytes / (s["sec"] / ss) / 1e9,
                                 "plane": "fp16 shadow of the fp32 rows (+ exact fp32 re-rank of the candidates)" if shadow_scan
                                          else "the stored rows",
                                 "fp32_rows_equivalent_gbs": fp32_bytes / (s["sec"] / ss) / 1e9,
                                 "traffic": self.traffic.get(f"scan_topk_kernel|{tkey}" + (" shadow" if shadow_scan else ""))},
                    "gpu_launches": s["launches"],
                }
                shard_bytes = fp32_bytes

        # roofline of the batched leg's dominant kernel
        tensor_batches = ix.get_stat("tensor_batches")
        flops = 2.0 * B * (hi - lo) * wl.dim          # this rank's share: the whole batch against its rows
        peak = peaks["bf16_tflops"] * (0.5 if tf32_path else 1.0)
        peak_s = peaks["bf16_tflops_sustained"] * (0.5 if tf32_path else 1.0)
        if tensor_batches > 0:
            # one search = several launches of gemm_filter_kernel (probe + one per threshold level) that together
            # contract every query with every row once: algorithmic flops per search / summed launch time
            t_step = main["kern_ns"] * 1e-9 / max(main["steps"], 1)
            ach = flops / t_step / 1e12
            roof = {"bound": "tensor", "kernel": "gemm_filter_kernel", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                    "frac": ach / peak,
                    "peak_source": peaks["source"] + (" bf16 burst x 0.5 (kind::tf32 runs at half the bf16 rate)" if tf32_path
                                                      else " bf16 burst (kind::f16, same tensor rate)"),
                    "algorithmic_flops_per_step": flops, "kernel_us_per_step": t_step * 1e6,
                    "launches_per_step": main["nprof"] / max(main["steps"], 1),
                    "step_level_frac": flops / (main["sec"] / main["steps"]) / 1e12 / peak,
                    "traffic": self.traffic.get(f"gemm_filter_kernel|{tkey}, batch {B}")}
        else:
            t_kernel = main["kern_ns"] * 1e-9 / max(main["nprof"], 1)
            if ix.get_stat("shadow_scans") > 0 and B == 1:       # the scan streamed the fp16 shadow plane
                shard_bytes = (hi - lo) * ix.get_stat("ld16") * 2
            ach = shard_bytes / t_kernel / 1e9
            roof = {"bound": "hbm", "kernel": "scan_topk_kernel", "achieved": ach, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                    "frac": ach / peaks["hbm_gbs"], "peak_source": peaks["source"],
                    "algorithmic_bytes_per_launch": shard_bytes, "kernel_us": t_kernel * 1e6,
                    "launches_per_step": main["nprof"] / max(main["steps"], 1),
                    "traffic": self.traffic.get(f"scan_topk_kernel|{tkey}")}

        # sustained: the same step back to back until the board sits at its power-capped steady state
        sustained = None
        if a.sustain_s > 0 and headline:
            with ClockSampler(self.local) as sclk:
                sl = device_leg(B, max(steps, 50), warmup, seconds=a.sustain_s)
            ach_s = flops / (sl["sec"] / sl["steps"]) / 1e12
            sustained = {"value": B * sl["steps"] / sl["sec"], "unit": "queries/s", "steps": sl["steps"],
                         "seconds": sl["sec"], "ms_per_step": 1e3 * sl["sec"] / sl["steps"],
                         "step_level_tflops": ach_s, "frac_of_burst_peak": ach_s / peak,
                         "frac_of_sustained_peak": ach_s / peak_s, "sustained_peak": peak_s, "clocks": sclk.summary()}

        # two searches in flight on two streams (what two server threads produce): the short launches of one search
        # -- probe, selects, re-rank -- fill the ramps and tails of the other's tensor launches.  Reported beside the
        # headline, which stays one search at a time on one stream.
        two_streams = None
        if headline and world == 1 and a.sustain_s > 0:
            qs = [self.make_queries(wl, B, start=i * B) for i in range(2)]
            outs = [(torch.empty((B, k), dtype=torch.int64, device=dev), torch.empty((B, k), dtype=torch.float32, device=dev),
                     torch.empty((B,), dtype=torch.int32, device=dev)) for _ in range(2)]
            streams = [torch.cuda.Stream(device=dev) for _ in range(2)]

            def run(nsteps):
                for i in range(nsteps):
                    j = i & 1
                    ix.search_device(qs[j].data_ptr(), B, k, outs[j][0].data_ptr(), outs[j][1].data_ptr(), outs[j][2].data_ptr(),
                                     streams[j].cuda_stream)
            run(2 * warmup)
            torch.cuda.synchronize()
            n2 = 2 * max(steps, 20)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for st_ in streams:
                st_.wait_event(e0)
            run(n2)
            for st_ in streams:
                torch.cuda.current_stream().wait_stream(st_)
            e1.record()
            torch.cuda.synchronize()
            ms2 = e0.elapsed_time(e1)
            two_streams = {"value": B * n2 / (ms2 * 1e-3), "unit": "queries/s", "steps": n2, "ms_per_step": ms2 / n2,
                           "step_level_frac": flops / (ms2 * 1e-3 / n2) / 1e12 / peak,
                           "note": "two independent batches in flight (one stream + workspace each)"}

        # ---- checks (outside every timed region) -----------------------------------------------
        check = None
        if not a.no_check:
            check = self.check(wl, ix, sx, main, lo, hi)

        cpu = None
        if headline and rank == 0 and world == 1 and not a.no_cpu:
            # bounded sample of the same workload, sized from a short (cold) probe to 10-20 s of CPU work
            nq_cpu = a.cpu_queries
            if not nq_cpu:
                probe_qps = cpu_knn_qps(wl, 64)[0]
                nq_cpu = int(min(B * 8, max(64, 64 * round(20.0 * probe_qps / 64))))
            qps, cores, desc, want = cpu_knn_qps(wl, nq_cpu, return_results=True)
            cpu = {"value": qps, "unit": "queries/s", "cores": cores, "kind": "port",
                   "sample": desc + " (oracle/knn_ref.c exact scan, OpenMP)"}
            if want is not None:      # the whole batch against the oracle (the CPU sample starts with the batch's queries)
                m = min(len(want[0]), B)
                full = compare_topk(main["ids"][:m], main["dist"][:m], want[0][:m], want[1][:m])
                check = dict(check or {}, whole_batch=full)

        out = {
            "workload": wl.name(), "value": B * main["steps"] / main["sec"], "unit": "queries/s",
            "ms_per_step": 1e3 * main["sec"] / main["steps"], "global_batch": B, "rows_total": wl.rows,
            "rows_per_gpu": hi - lo, "k": k, "path": "tensor" if tensor_batches > 0 else "scan",
            "dtype": ("f16 operands (shadow plane of f32 rows) / f32 accumulate, exact f32 re-rank" if shadow else
                      "tf32 operands / f32 accumulate, exact f32 re-rank" if tf32_path else "f16-stored / f32 accumulate"),
            "e2e": {"value": B * steps / e2e_sec, "unit": "queries/s", "h2d_bytes_per_step": B * wl.dim * 4,
                    "d2h_bytes_per_step": B * k * 12 + (B * 4 if world == 1 else 0),
                    "in_flight": 2,
                    "one_at_a_time": B * steps / e2e_serial_sec if e2e_serial_sec else None,
                    "note": "two batches in flight (" + ("vdb_search_submit / vdb_search_collect" if world == 1 else
                                                         "ShardedIndex.submit_host / collect") +
                            "): every step uploads its queries from page-locked host memory and has its results in host "
                            "memory when collected; one_at_a_time = the blocking call, one batch after the other"},
            "gpu_launches": main["launches"], "roofline": roof, "single_query": single, "sustained": sustained,
            "two_streams": two_streams,
            "check": check, "cpu_baseline": cpu, "clocks": clk.summary(),
            "fallback_queries": ix.get_stat("fallback_queries"), "steps": main["steps"],
        }
        if sx is not None:
            sx.close()
        ix.close()
        torch.cuda.empty_cache()
        return out

    # ---- parity of what the timed legs returned ------------------------------------------------
    def check(self, wl, ix, sx, main, lo, hi):
        torch, dist, vdb, a = self.torch, self.dist, self.vdb, self.a
        world, rank, dev, stream = self.world, self.rank, self.dev, self.stream
        B, k = wl.batch, wl.k
        out = {}
        # (1) fused NVLink exchange vs NCCL exchange, one step, bitwise
        if sx is not None:
            q = self.make_queries(wl, B)
            d1, i1 = sx.search_device(q, k)
            d1, i1 = d1.clone(), i1.clone()
            other = vdb.ShardedIndex(ix, max_batch=B, max_k=k, exchange="nccl" if a.exchange == "p2p" else "p2p")
            d2, i2 = other.search_device(q, k)
            same = torch.tensor([int(torch.equal(i1, i2) and torch.equal(d1.view(torch.int32), d2.view(torch.int32)))],
                                device=dev)
            dist.all_reduce(same, op=dist.ReduceOp.MIN)
            out["p2p_equals_nccl"] = bool(same.item())
            other.close()
            # the fused NVLink exchange + merge kernel (K5x) alone: this rank's [B, k] lists out by query slice, its
            # slice's G lists merged; (G-1)/G of the packed keys cross NVLink
            px = sx.px if a.exchange == "p2p" else None
            if px is not None:
              try:
                ids_l = torch.arange(B * k, dtype=torch.int64, device=dev).view(B, k)
                dd_l = torch.rand((B, k), device=dev).sort(dim=1).values
                sl = (B + world - 1) // world
                o_i = torch.empty((sl, k), dtype=torch.int64, device=dev)
                o_d = torch.empty((sl, k), dtype=torch.float32, device=dev)
                reps = 20
                for _ in range(3):
                    px.merge(dd_l.data_ptr(), ids_l.data_ptr(), B, k, o_d.data_ptr(), o_i.data_ptr(), stream)
                self.barrier()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(reps):
                    px.merge(dd_l.data_ptr(), ids_l.data_ptr(), B, k, o_d.data_ptr(), o_i.data_ptr(), stream)
                e1.record()
                torch.cuda.synchronize()
                px.status()
                us = torch.tensor([e0.elapsed_time(e1) * 1e3 / reps], device=dev)
                dist.all_reduce(us, op=dist.ReduceOp.MAX)
                nv_bytes = B * k * 8 * (world - 1) // world
                out["exchange_merge_kernel"] = {"us_per_step": float(us.item()), "nvlink_bytes_out_per_rank": nv_bytes,
                                                "nvlink_gbs_per_rank": nv_bytes / (float(us.item()) * 1e-6) / 1e9,
                                                "note": "K5x alone, back to back, max over ranks (CUDA events)"}
              except Exception as e:          # an optional figure must not cost the line
                out["exchange_merge_kernel"] = {"error": repr(e)}
        # (2) sampled queries of rank 0's slice against the CPU oracle over the whole database
        n_mine = len(main["ids"])
        nsamp = min(a.check_queries, n_mine)
        q_idx = np.unique(np.linspace(0, n_mine - 1, nsamp).astype(np.int64)) if nsamp else np.zeros(0, np.int64)
        # (3) >= 10M rows: GPU cross-check against a chunked fp32 torch.matmul + topk over every shard (test-only torch)
        xc = None
        if wl.rows >= 10_000_000 and len(q_idx):
            xc = self.torch_xcheck(wl, q_idx, lo, hi)
        if rank == 0 and len(q_idx):
            got_i, got_d = main["ids"][q_idx], main["dist"][q_idx]
            want_i, want_d, info = oracle_topk(wl, q_idx, a.oracle_budget_s)
            if want_i is not None:
                out["oracle"] = dict(compare_topk(got_i, got_d, want_i, want_d), **info,
                                     scope="sampled queries of rank 0's slice vs exact CPU scan of all rows")
            else:
                want_d = oracle_distances_of(wl, q_idx, got_i)
                rel = np.abs(got_d.astype(np.float64) - want_d) / np.maximum(1.0, np.abs(want_d))
                out["oracle"] = dict(info, queries=int(len(q_idx)), max_rel_dist_err=float(rel.max()),
                                     sorted=bool((np.diff(got_d, axis=1) >= 0).all()),
                                     scope="full CPU scan over budget: exact CPU distances of the returned rows only; "
                                           "completeness from torch_xcheck")
            if xc is not None:
                out["torch_xcheck"] = dict(compare_topk(got_i, got_d, xc[0], xc[1]),
                                           scope="fp32 torch.matmul + topk over every shard (rtol 1e-5; summation order differs)")
        self.barrier()
        return out

    def torch_xcheck(self, wl, q_idx, lo, hi, chunk=262_144):
        """Exact fp32 top-k of the sampled queries over this rank's rows with plain torch (rows regenerated on the
        device chunk by chunk), gathered and merged over the ranks.  Returns (ids, dist) on every rank."""
        torch, dist = self.torch, self.dist
        old = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        try:
            q = torch.cat([self.make_queries(wl, 1, int(i)) for i in q_idx])
            if wl.metric == "cosine":
                q = q / q.norm(dim=1, keepdim=True)
            k = wl.k
            best_d = torch.full((len(q_idx), 0), float("inf"), device=self.dev)
            best_i = torch.full((len(q_idx), 0), -1, dtype=torch.int64, device=self.dev)
            buf = torch.empty((chunk, wl.dim), dtype=torch.float32, device=self.dev)
            for r0 in range(lo, hi, chunk):
                n = min(chunk, hi - r0)
                rows = buf[:n]
                self.vdb._ffi.check(self.lib.vdb_synth_dev(SEED_DB, r0, n, wl.dim, rows.data_ptr(), self.stream), "synth")
                if wl.metric == "cosine":
                    rows = rows / rows.norm(dim=1, keepdim=True)
                if wl.store == "f16":
                    rows = rows.half().float()
                dot = q @ rows.T
                d = (q * q).sum(1, keepdim=True) + (rows * rows).sum(1)[None, :] - 2 * dot if wl.metric == "l2" else 1.0 - dot
                cd, ci = torch.topk(d, min(k, n), dim=1, largest=False)
                best_d = torch.cat([best_d, cd], dim=1)
                best_i = torch.cat([best_i, ci + r0], dim=1)
                if best_d.shape[1] > 8 * k:
                    cd, sel = torch.topk(best_d, k, dim=1, largest=False)
                    best_d, best_i = cd, torch.gather(best_i, 1, sel)
            cd, sel = torch.topk(best_d, min(k, best_d.shape[1]), dim=1, largest=False)
            best_d, best_i = cd.contiguous(), torch.gather(best_i, 1, sel).contiguous()
            if self.world > 1:
                g_d = torch.empty((self.world,) + tuple(best_d.shape), device=self.dev)
                g_i = torch.empty((self.world,) + tuple(best_i.shape), dtype=torch.int64, device=self.dev)
                dist.all_gather_into_tensor(g_d, best_d)
                dist.all_gather_into_tensor(g_i, best_i)
                g_d = g_d.permute(1, 0, 2).reshape(len(q_idx), -1)
                g_i = g_i.permute(1, 0, 2).reshape(len(q_idx), -1)
                best_d, sel = torch.topk(g_d, k, dim=1, largest=False)
                best_i = torch.gather(g_i, 1, sel)
            # ascending (distance, id) like the product
            order = np.lexsort((best_i.cpu().numpy(), best_d.cpu().numpy()), axis=1)
            return (np.take_along_axis(best_i.cpu().numpy(), order, axis=1),
                    np.take_along_axis(best_d.cpu().numpy(), order, axis=1))
        finally:
            torch.backends.cuda.matmul.allow_tf32 = old


def run_ours(a):
    arm = Arm(a)
    world, rank = arm.world, arm.rank
    wl = headline_workload(a, world)
    head = arm.measure(wl, a.steps, a.warmup, headline=True)
    extras = {}
    for w in extra_workloads(a, world):
        extras[w.key] = arm.measure(w, max(3, min(a.steps, 10)), max(3, min(a.warmup, 3)), headline=False)
    if rank == 0:
        chk = head["check"] or {}
        recall = (chk.get("whole_batch") or chk.get("oracle") or {}).get("recall_at_k")
        line = {
            "metric": "queries/sec exact top-k", "value": head["value"], "unit": "queries/s",
            "n_gpus": world, "steps": a.steps, "warmup": a.warmup, "ms_per_step": head["ms_per_step"],
            "higher_is_better": True, "scaling": a.scaling if world > 1 else "weak", "vs_baseline": None,
            "dtype": head["dtype"], "data": "synthetic",
            "config": {"workload": wl.name(), "global_batch": wl.batch, "rows_total": wl.rows,
                       "rows_per_gpu": head["rows_per_gpu"], "sharding": f"contiguous rows x{world}",
                       "exchange": "none" if world == 1 else (
                           "fused NVLink peer-store exchange + merge kernel, by query slice" if a.exchange == "p2p"
                           else "all-to-all by query slice (NCCL) + GPU merge"),
                       "l2": "inputs larger than L2 (no flush needed)", "path": head["path"]},
            "e2e": head["e2e"], "gpu_launches": head["gpu_launches"], "roofline": head["roofline"],
            "cpu_baseline": head["cpu_baseline"], "single_query": head["single_query"], "sustained": head["sustained"],
            "two_streams": head["two_streams"],
            "clocks": head["clocks"], "recall_at_k": recall, "check": head["check"],
            "fallback_queries": head["fallback_queries"],
        }
        if extras:
            line["configs"] = extras
        print(json.dumps(line), flush=True)
    if world > 1:
        arm.dist.barrier()
        arm.dist.destroy_process_group()


if __name__ == "__main__":
    args = parse()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)
