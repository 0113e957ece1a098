#!/usr/bin/env bash
# Three quick bench legs on one GPU: headline, config-3 shard, the N=8 per-GPU shape (batch 8192 x 125k rows).
run() {
  label=$1; shift
  out=$(python bench.py --no-cpu --sustain-s 0 --steps 20 --warmup 3 "$@" 2>/dev/null)
  python - "$label" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); r = d["roofline"]; c = (d.get("check") or {}).get("oracle") or {}; s = d.get("single_query") or {}
sr = (s.get("roofline") or {})
print(f"{sys.argv[1]:28s} ms/step {d['ms_per_step']:.4f} kernel_us {r.get('kernel_us_per_step', 0):8.1f} launches {d['gpu_launches']//d['steps']}"
      f" frac {r['frac']:.3f} step_frac {r.get('step_level_frac', 0):.3f} e2e {d['e2e']['value']:.0f} parity {c.get('parity_ok')} fb {d['fallback_queries']}"
      f" | b1 {s.get('ms_per_query', 0):.4f} ms kern {sr.get('kernel_us', 0):.1f} us api {sr.get('api_level_gbs', 0):.0f} GB/s")
PY
}
run "headline" "$@"
run "config3-shard" --rows 1250000 --metric l2 --k 100 --batch 4096 "$@"
run "b8192x125k" --rows 125000 --batch 8192 --no-single "$@"
