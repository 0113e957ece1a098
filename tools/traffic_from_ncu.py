#!/usr/bin/env python
"""profiles/traffic.json from the ncu --set full raw pages of this round (tools/profile_round.sh):
dram__bytes_read.sum + dram__bytes_write.sum per launch; for the tensor kernel the SUM over the launches of one search
(probe + levels), like roofline.achieved.   tools/traffic_from_ncu.py r02 > profiles/traffic.json"""
import csv, json, os, sys

R = sys.argv[1] if len(sys.argv) > 1 else "r02"
P = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "profiles")
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}


def launches(path):
    rows = list(csv.reader(open(path)))
    h, u = rows[0], rows[1]
    ir, iw, it, ik = (h.index(c) for c in ("dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "Kernel Name"))
    pipe = [i for i, c in enumerate(h) if c == "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"]
    clk = [i for i, c in enumerate(h) if c == "sm__cycles_elapsed.avg.per_second"]
    out = []
    for r in rows[2:]:
        f = lambda i: float(r[i].replace(",", "")) if r[i] not in ("", "n/a") else 0.0
        out.append({"kernel": r[ik], "dram_bytes": f(ir) * UNIT[u[ir]] + f(iw) * UNIT[u[iw]],
                    "us": f(it) * {"ns": 1e-3, "us": 1.0, "ms": 1e3, "usecond": 1.0, "nsecond": 1e-3, "msecond": 1e3}.get(u[it], 1.0),
                    "tensor_pipe_pct": round(f(pipe[0]), 1) if pipe else None, "sm_ghz": round(f(clk[0]), 3) if clk else None})
    return out


CAPTURES = {   # raw page -> (traffic key, "sum" of all launches or "each")
    f"{R}_gemm_f16shadow_cos_k10_b1024_ncu_raw.csv": ("gemm_filter_kernel|1000000x512 f32 cosine top-10, batch 1024", "sum"),
    f"{R}_gemm_tf32_cos_k10_b1024_ncu_raw.csv": ("gemm_filter_kernel[tf32, VDB_SHADOW=0]|1000000x512 f32 cosine top-10, batch 1024", "sum"),
    f"{R}_gemm_f16shadow_l2_k100_b4096_ncu_raw.csv": ("gemm_filter_kernel|1250000x512 f32 l2 top-100, batch 4096", "sum"),
    f"{R}_scan_f32_k10_ncu_raw.csv": ("scan_topk_kernel|1000000x512 f32 cosine top-10", "each"),
    f"{R}_scan_f16shadow_k10_ncu_raw.csv": ("scan_topk_kernel|1000000x512 f32 cosine top-10 shadow", "each"),
    f"{R}_scan_f32_l2_k100_ncu_raw.csv": ("scan_topk_kernel|1250000x512 f32 l2 top-100", "each"),
    f"{R}_select_l2_k100_b4096_ncu_raw.csv": ("select_kernel|1250000x512 f32 l2 top-100, batch 4096", "each"),
    f"{R}_rerank_l2_k100_b4096_ncu_raw.csv": ("rerank_window_kernel|1250000x512 f32 l2 top-100, batch 4096", "each"),
}
out = {"_comment": "dram__bytes_read.sum + dram__bytes_write.sum from the ncu --set full raw pages committed beside this file "
                   "(tools/profile_round.sh, tools/traffic_from_ncu.py); bench.py copies the matching entry into roofline.traffic. "
                   "gemm_filter_kernel: SUM over the launches of one search (probe + levels), like roofline.achieved",
       "_per_launch": {}}
for name, (key, mode) in CAPTURES.items():
    path = os.path.join(P, name)
    if not os.path.exists(path):
        continue
    ls = launches(path)
    if not ls:
        continue
    if mode == "sum":
        # one search = probe (the shortest launch) + its levels (4 launches at these shapes); a capture window may
        # start in the middle of a search: the levels missing after the probe are taken from the search before it
        cycle = 4
        i0 = min(range(len(ls)), key=lambda i: ls[i]["us"])
        have = ls[i0:i0 + cycle]
        missing = cycle - len(have)
        ls = have + (ls[i0 - missing:i0] if missing > 0 else [])
    out[key] = int(sum(l["dram_bytes"] for l in ls)) if mode == "sum" else int(ls[0]["dram_bytes"])
    out["_per_launch"][key] = [{"us": round(l["us"], 1), "dram_MB": round(l["dram_bytes"] / 1e6, 1), "tensor_pipe_pct": l["tensor_pipe_pct"], "sm_ghz": l["sm_ghz"]} for l in ls]
print(json.dumps(out, indent=1))
