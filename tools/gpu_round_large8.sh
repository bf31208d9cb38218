#!/bin/bash
# 8-GPU large window: grouped vs per-column backward sweep on the same box.  usage (under gpurun --gpus 8): bash tools/gpu_round_large8.sh [tag]
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
CNGP_LARGE_PHASES=1 timeout 300 $TR tools/bench_large.py 32768 5 > $O/bench_large_8gpu_$TAG.json 2> $O/bench_large_8gpu_$TAG.err; echo "groups rc=$?"; cut -c1-500 $O/bench_large_8gpu_$TAG.json
CNGP_LARGE_PHASES=1 CNGP_LARGE_NO_GROUPS=1 timeout 300 $TR tools/bench_large.py 32768 5 > $O/bench_large_8gpu_nogroups_$TAG.json 2> $O/bench_large_8gpu_nogroups_$TAG.err; echo "nogroups rc=$?"; cut -c1-500 $O/bench_large_8gpu_nogroups_$TAG.json
tail -c 300 $O/bench_large_8gpu_$TAG.err
