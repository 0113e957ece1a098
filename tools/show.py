import json, sys
for path in sys.argv[1:]:
    try:
        d = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as e:
        print(path, "unreadable", e); continue
    sq = d.get("single_query") or {}
    r = d["roofline"]
    print(path, "| N", d["n_gpus"], d["scaling"], "| value %.0f q/s  %.3f ms/step | e2e %.0f | kernel %.1f us %s %.0f (frac %.3f) | single %.0f q/s e2e %.0f frac %.3f | launches %s fallback %s recall %s" % (
        d["value"], d["ms_per_step"], d["e2e"]["value"], r.get("kernel_us_per_step", r.get("kernel_us", 0)), r["unit"], r["achieved"], r["frac"],
        sq.get("value", 0), (sq.get("e2e") or {}).get("value", 0), (sq.get("roofline") or {}).get("frac", 0), d.get("gpu_launches"), d.get("fallback_queries"), d.get("recall_at_k")))
