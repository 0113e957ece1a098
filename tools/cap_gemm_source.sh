#!/usr/bin/env bash
# ncu --set full of ONE launch of the tensor kernel (default: the last level of the third batch-1024 search) with the
# source page (per-line warp-state samples): which wait paces the MMA issuer / the producer / the epilogue.
#   tools/cap_gemm_source.sh <name> [launch-skip] [bench args...]
set -u
name=${1:-gemm_l3}; skip=${2:-11}; shift 2 || true
O=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_filter --launch-skip $skip --launch-count 1 \
    -o $O/$name python bench.py --no-cpu --no-check --sustain-s 0 --no-single --steps 3 --warmup 1 "$@" > $O/ncu_$name.log 2>&1
ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>> $O/ncu_$name.log
ncu -i $O/$name.ncu-rep --page source --csv > $O/${name}_source.csv 2>> $O/ncu_$name.log
rm -f $O/$name.ncu-rep
ls -la $O/${name}_*.csv
