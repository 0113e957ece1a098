#!/usr/bin/env bash
# One GPU box, one call: the round's evidence under gpurun_out/ (copy what should be judged into profiles/).
#   tools/profile_round.sh r02
# 1. the bench line and the reference arm (no profiler)   2. ncu launch list of the SAME bench command
# 3. ncu --set full raw pages: the tensor kernel (fp16 shadow plane; tf32 on the fp32 rows; L2 k=100 batch 4096),
#    K2s select, K4w window re-rank, the scan kernel (single query over the fp16 shadow plane; fp32 rows k=10 and k=100)
set -u
R=${1:-r02}; O=gpurun_out
NCU="ncu --set full --clock-control none --import-source on"
Q="--no-cpu --no-check --sustain-s 0"
C3="--rows 1250000 --metric l2 --k 100 --batch 4096"
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err || exit 1
python bench.py --impl reference > $O/${R}_bench_reference.json 2>> $O/${R}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --no-cpu --sustain-s 0 > $O/${R}_bench_under_ncu.log 2>&1
python tools/launches.py $O/${R}_launches_bench.csv > $O/${R}_launches_bench_summary.txt
cap() {  # name, kernel regex, skip, count, env..., -- bench args
  name=$1; re=$2; skip=$3; count=$4; shift 4
  envs=(); while [ "$1" != "--" ]; do envs+=("$1"); shift; done; shift
  env "${envs[@]}" $NCU --kernel-name regex:$re --launch-skip $skip --launch-count $count -o $O/${R}_$name \
      python bench.py $Q "$@" > $O/ncu_$name.log 2>&1
  ncu -i $O/${R}_$name.ncu-rep --page raw --csv > $O/${R}_${name}_ncu_raw.csv 2>> $O/ncu_$name.log
  rm -f $O/${R}_$name.ncu-rep        # the raw page is what is kept: gpurun brings back at most 64 MiB
}
# steady state: skip the first 2 searches (probe + levels launches each)
cap gemm_f16shadow_cos_k10_b1024 gemm_filter 8 4 X=1 -- --no-single --steps 3 --warmup 1
cap gemm_tf32_cos_k10_b1024 gemm_filter 8 4 VDB_SHADOW=0 -- --no-single --steps 3 --warmup 1
cap gemm_f16shadow_l2_k100_b4096 gemm_filter 10 5 X=1 -- --no-single --steps 3 --warmup 1 $C3
cap select_l2_k100_b4096 select_kernel 9 1 X=1 -- --no-single --steps 3 --warmup 1 $C3
cap rerank_l2_k100_b4096 rerank_window 2 1 X=1 -- --no-single --steps 3 --warmup 1 $C3
cap scan_f16shadow_k10 scan_topk 3 1 X=1 -- --batch 1 --steps 3 --warmup 3 --no-single
cap scan_f32_k10 scan_topk 3 1 X=1 -- --batch 1 --steps 3 --warmup 3 --no-single --no-shadow-scan
cap scan_f32_l2_k100 scan_topk 3 1 X=1 -- --batch 1 --steps 3 --warmup 3 --no-single --rows 1250000 --metric l2 --k 100
python tools/show.py $O/${R}_bench_n1.json
tail -14 $O/${R}_launches_bench_summary.txt
ls -la $O/${R}_*ncu_raw.csv
