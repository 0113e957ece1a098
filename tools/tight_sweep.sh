#!/usr/bin/env bash
# A/B of the level-threshold rule and the level growth on one GPU (headline shape and the config-3 shard).
#   tools/tight_sweep.sh > gpurun_out/tight_sweep.txt
run() {  # label, env..., -- bench args
  label=$1; shift
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  out=$(env "${envs[@]}" python bench.py --no-single --no-cpu --sustain-s 0 --steps 20 --warmup 3 "$@" 2>/dev/null)
  python - "$label" "$out" <<'PY'
import json, sys
d = json.loads(sys.argv[2]); r = d["roofline"]; c = (d.get("check") or {}).get("oracle") or {}
print(f"{sys.argv[1]:44s} ms/step {d['ms_per_step']:.4f}  kernel_us {r.get('kernel_us_per_step', 0):8.1f}  launches/step {r.get('launches_per_step')}"
      f"  frac {r['frac']:.3f} step_frac {r.get('step_level_frac', 0):.3f}  parity {c.get('parity_ok')} fallback {d['fallback_queries']}")
PY
}
H=()
C3=(--rows 1250000 --metric l2 --k 100 --batch 4096)
for m in 0 2.5; do
  for g in 8 16; do
    run "headline margin=$m growth=$g" VDB_TIGHT_MARGIN=$m VDB_GROWTH=$g -- "${H[@]}"
  done
done
run "headline margin=2.5 growth=32" VDB_TIGHT_MARGIN=2.5 VDB_GROWTH=32 -- "${H[@]}"
for m in 0 2.5; do
  for g in 4 8 16; do
    run "config3-shard margin=$m growth=$g" VDB_TIGHT_MARGIN=$m VDB_GROWTH=$g -- "${C3[@]}"
  done
done
run "b8192x125k margin=0 growth=8" VDB_TIGHT_MARGIN=0 VDB_GROWTH=8 -- --rows 125000 --batch 8192
run "b8192x125k margin=2.5 growth=8" VDB_TIGHT_MARGIN=2.5 VDB_GROWTH=8 -- --rows 125000 --batch 8192
run "b8192x125k margin=2.5 growth=16" VDB_TIGHT_MARGIN=2.5 VDB_GROWTH=16 -- --rows 125000 --batch 8192
