#!/bin/bash
# Multi-GPU round (run under `gpurun --gpus N`): window-sharded bench, block-cyclic large window, Monte-Carlo shard.
# usage: bash tools/gpu_round_multi.sh <N> [tag]
N=${1:-2}; TAG=${2:-r01}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $O/topo_${N}gpu_$TAG.txt 2>&1
timeout 600 $TR bench.py --gpus $N --steps 10 --warmup 3 > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err; echo "bench rc=$?"
timeout 600 $TR bench.py --impl reference --gpus $N --steps 2 --warmup 1 > $O/bench_ref_${N}gpu_$TAG.json 2> $O/bench_ref_${N}gpu_$TAG.err; echo "ref rc=$?"
timeout 600 $TR tools/bench_large.py 32768 3 > $O/bench_large_${N}gpu_$TAG.json 2> $O/bench_large_${N}gpu_$TAG.err; echo "large rc=$?"
timeout 600 $TR tools/bench_configs.py mc > $O/bench_mc_${N}gpu_$TAG.json 2> $O/bench_mc_${N}gpu_$TAG.err; echo "mc rc=$?"
tail -c 600 $O/bench_${N}gpu_$TAG.err; cut -c1-400 $O/bench_${N}gpu_$TAG.json; cat $O/bench_large_${N}gpu_$TAG.json; tail -c 600 $O/bench_large_${N}gpu_$TAG.err; cat $O/bench_mc_${N}gpu_$TAG.json; tail -c 600 $O/bench_mc_${N}gpu_$TAG.err
