#!/bin/bash
# Capture one kernel with ncu --set full on the GPU box and keep only the CSV exports (the .ncu-rep files are ~45 MB
# each and gpurun_out/ is capped at 64 MiB).   usage: tools/ncu_capture.sh <tag> <kernel-regex> <skip> <command...>
# The regex is matched against the DEMANGLED name (template arguments included), so one instantiation can be picked:
#   'gp_var_kernel<.int.32, .int.1, .int.12, .*bool.1, .bool.1>'  = the lag-table variance kernel of the bench workload
#   (ncu prints template arguments as (int)32, (bool)1).
TAG=$1; KRE=$2; SKIP=$3; shift 3
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k "regex:$KRE" -s $SKIP -c 1 -f -o /tmp/prof_$TAG "$@" > $O/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $O/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > $O/ncu_${TAG}_source.csv 2>/dev/null
rm -f /tmp/prof_$TAG.ncu-rep
ls -la $O/ncu_${TAG}_raw.csv $O/ncu_${TAG}_source.csv
