#!/usr/bin/env bash
# Level-plan sweep of the tensor path: dense tiles of level 0 x growth per level, at the given bench shape.
# usage: tools/level_sweep.sh [bench.py args]     (one GPU)
for d in 4 1 2; do for g in 8 6 12; do
  VDB_DENSE_TILES=$d VDB_GROWTH=$g python bench.py --no-cpu --no-single --steps 10 "$@" > /tmp/ls.json 2>/tmp/ls.err
  python - "$d" "$g" <<'PY'
import json, sys
d = json.loads(open('/tmp/ls.json').read().strip().splitlines()[-1])
r = d['roofline']
print("dense=%s growth=%s | %.3f ms/step | gemm %.1f us x%.0f launches | fallback %s" % (sys.argv[1], sys.argv[2], d['ms_per_step'], r['kernel_us_per_step'], r['launches_per_step'], d['fallback_queries']))
PY
done; done
