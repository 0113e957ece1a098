#!/usr/bin/env bash
# One GPU box, one call: the round's evidence under gpurun_out/ (copy what should be judged into profiles/).
#   tools/profile_round.sh r01
# 1. the bench line and the reference arm (no profiler)   2. ncu launch list of the SAME bench command
# 3. ncu --set full of one search's gemm_filter launches and of one scan_topk launch (raw pages as csv)
set -u
R=${1:-r01}; O=gpurun_out
python bench.py > $O/${R}_bench_n1.json 2> $O/${R}_bench_n1.err || exit 1
python bench.py --impl reference > $O/${R}_bench_reference.json 2>> $O/${R}_bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/${R}_launches_bench.csv \
    python bench.py --no-cpu > $O/${R}_bench_under_ncu.log 2>&1
python tools/launches.py $O/${R}_launches_bench.csv > $O/${R}_launches_bench_summary.txt
# steady state: skip the first 2 searches (3 gemm launches each... the level count is printed by the summary)
ncu --set full --clock-control none --import-source on --kernel-name regex:gemm_filter --launch-skip 8 --launch-count 4 \
    -o $O/${R}_gemm_filter python bench.py --no-cpu --no-single --steps 3 --warmup 1 > $O/ncu_gemm.log 2>&1
ncu --set full --clock-control none --import-source on --kernel-name regex:scan_topk --launch-skip 3 --launch-count 1 \
    -o $O/${R}_scan_topk python bench.py --no-cpu --batch 1 --steps 3 --warmup 3 --no-single > $O/ncu_scan.log 2>&1
ncu -i $O/${R}_gemm_filter.ncu-rep --page raw --csv > $O/${R}_gemm_filter_ncu_raw.csv
ncu -i $O/${R}_scan_topk.ncu-rep --page raw --csv > $O/${R}_scan_topk_ncu_raw.csv
python tools/show.py $O/${R}_bench_n1.json
tail -12 $O/${R}_launches_bench_summary.txt
