#!/usr/bin/env bash
# Builds libvdb_b200.so in-tree for sm_100a.  Usage: csrc/build.sh [extra nvcc flags]
set -euo pipefail
cd "$(dirname "$0")"
NVCC=${NVCC:-/usr/local/cuda/bin/nvcc}
OUT=../libvdb_b200.so
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -Wall
       -ccbin /usr/bin/g++ --expt-relaxed-constexpr "$@")
mkdir -p _obj
pids=()
for f in scan_topk merge_topk insert gemm_topk vdb_api; do
  if [ ! -f _obj/$f.o ] || [ $f.cu -nt _obj/$f.o ] || [ -n "$(find . -maxdepth 1 \( -name '*.cuh' -o -name '*.h' \) -newer _obj/$f.o)" ] || [ ../../include/vdb.h -nt _obj/$f.o ]; then
    "$NVCC" "${FLAGS[@]}" -c $f.cu -o _obj/$f.o &
    pids+=($!)
  fi
done
for p in "${pids[@]:-}"; do [ -n "$p" ] && wait "$p"; done
"$NVCC" -shared -o "$OUT" _obj/*.o
echo "built $(realpath $OUT)"
