#!/bin/bash
# Multi-GPU round (run under `gpurun --gpus N`): window-sharded bench (its extra.configs carry the Monte-Carlo shard and
# the block-cyclic large window at this N), the reference arm, and the large window alone with per-rank kernel time.
# usage: bash tools/gpu_round_multi.sh <N> [tag]
N=${1:-2}; TAG=${2:-r02}
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
nvidia-smi topo -m > $O/topo_${N}gpu_$TAG.txt 2>&1
timeout 900 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > $O/bench_${N}gpu_$TAG.json 2> $O/bench_${N}gpu_$TAG.err; echo "bench rc=$?"
timeout 600 $TR tools/bench_large.py 32768 3 > $O/bench_large_${N}gpu_$TAG.json 2> $O/bench_large_${N}gpu_$TAG.err; echo "large rc=$?"
tail -c 400 $O/bench_${N}gpu_$TAG.err; python - <<PY
import json
d = json.load(open("$O/bench_${N}gpu_$TAG.json"))
print("value", d["value"], "e2e", d["e2e"]["value"], "ms", d["ms_per_step"])
ex = d.get("extra", {})
print("extra error:", ex.get("error"))
for k, v in ex.get("configs", {}).items():
    print(k, {kk: v[kk] for kk in ("ms_per_callback", "ms_total", "windows_per_s", "best_ms", "ms", "n_gpus", "tflops_n3_over_3", "lml") if kk in v})
PY
cat $O/bench_large_${N}gpu_$TAG.json; tail -c 300 $O/bench_large_${N}gpu_$TAG.err
