#!/usr/bin/env bash
# ncu --set full of ONE launch of a kernel with the SASS source page: tools/cap_kernel_source.sh <name> <kernel regex> <skip> -- <command...>
set -u
name=$1; re=$2; skip=$3; shift 4
O=gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:$re --launch-skip $skip --launch-count 1 -o $O/$name "$@" > $O/ncu_$name.log 2>&1
ncu -i $O/$name.ncu-rep --page raw --csv > $O/${name}_raw.csv 2>> $O/ncu_$name.log
ncu -i $O/$name.ncu-rep --page source --csv > $O/${name}_source.csv 2>> $O/ncu_$name.log
rm -f $O/$name.ncu-rep
ls -la $O/${name}_*.csv
