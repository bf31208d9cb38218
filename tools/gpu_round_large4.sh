#!/bin/bash
# configs[4] on 4 and 2 GPUs of one box.  usage (under gpurun --gpus 4): bash tools/gpu_round_large4.sh [tag]
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
for n in 4 2; do
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2954$n \
    tools/bench_large.py 32768 4 > $O/bench_large_${n}gpu_$TAG.json 2> $O/bench_large_${n}gpu_$TAG.err
  echo "N=$n rc=$?"; cut -c1-330 $O/bench_large_${n}gpu_$TAG.json
done
