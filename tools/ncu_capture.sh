#!/bin/bash
# Capture one kernel with ncu --set full on the GPU box and keep only the CSV exports (the .ncu-rep files are ~45 MB
# each and gpurun_out/ is capped at 64 MiB).   usage: tools/ncu_capture.sh <tag> <kernel-regex> <command...>
TAG=$1; KRE=$2; shift 2
O=gpurun_out; mkdir -p $O
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$KRE -s 1 -c 1 -f -o /tmp/prof_$TAG "$@" > $O/ncu_$TAG.log 2>&1
ncu -i /tmp/prof_$TAG.ncu-rep --page raw --csv > $O/ncu_${TAG}_raw.csv 2>/dev/null
ncu -i /tmp/prof_$TAG.ncu-rep --page source --csv > $O/ncu_${TAG}_source.csv 2>/dev/null
rm -f /tmp/prof_$TAG.ncu-rep
ls -la $O/ncu_${TAG}_raw.csv $O/ncu_${TAG}_source.csv
