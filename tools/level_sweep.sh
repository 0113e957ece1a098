#!/usr/bin/env bash
# Level-plan sweep of the tensor path: probe tiles x growth per level, at the given bench shape (one GPU).
# usage: tools/level_sweep.sh [bench.py args]
for cfg in "16 8" "16 16" "32 8" "32 16" "8 8" "64 8" "16 12"; do set -- $cfg "${@:3}"
  VDB_PROBE_TILES=$1 VDB_GROWTH=$2 python bench.py --no-cpu --no-single --steps 10 $SWEEP_ARGS > /tmp/ls.json 2>/tmp/ls.err
  python - "$1" "$2" <<'PY'
import json, sys
d = json.loads(open('/tmp/ls.json').read().strip().splitlines()[-1])
r = d['roofline']
print("probe=%s growth=%s | %.3f ms/step | gemm %.1f us x%.0f launches | fallback %s" % (sys.argv[1], sys.argv[2], d['ms_per_step'], r['kernel_us_per_step'], r['launches_per_step'], d['fallback_queries']))
PY
done
