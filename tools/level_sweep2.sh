#!/usr/bin/env bash
# probe tiles x growth x buffer multiplier at the given bench shape: tools/level_sweep2.sh [bench args]
source tools/brun.sh
for cfg in "16 8 16" "64 8 16" "128 8 32" "128 12 32" "128 16 32" "256 8 64" "64 16 16" "128 32 32"; do set -- $cfg "${@:4}"
  echo -n "probe=$1 growth=$2 capmult=$3: "; VDB_PROBE_TILES=$1 VDB_GROWTH=$2 VDB_CAP_MULT=$3 run $SWEEP_ARGS
done
