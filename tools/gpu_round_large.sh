#!/bin/bash
# The large window (configs[4]) on an N-GPU box: chunk-size variants of the pipelined panel chain at N ranks, then the
# default at the smaller world sizes.  usage (under gpurun --gpus N): bash tools/gpu_round_large.sh <N> <tag> [chunk rows ...]
N=${1:-8}; TAG=${2:-r02}; shift 2
VARIANTS=${@:-"4096 0 2048"}
O=gpurun_out; mkdir -p $O
export CNGP_LARGE_PHASES=1
for cr in $VARIANTS; do
  CNGP_LARGE_CHUNK_ROWS=$cr timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29521 tools/bench_large.py 32768 4 > $O/bench_large_${N}gpu_c${cr}_$TAG.json 2> $O/bench_large_${N}gpu_c${cr}_$TAG.err
  echo "N=$N chunk_rows=$cr rc=$?"; cut -c1-420 $O/bench_large_${N}gpu_c${cr}_$TAG.json; python - <<PY
import json
try:
    d = json.load(open("$O/bench_large_${N}gpu_c${cr}_$TAG.json")); print({k: round(v, 2) for k, v in d["phase_ms_rank"].items()})
except Exception as e:
    print("no json:", e)
PY
  tail -c 300 $O/bench_large_${N}gpu_c${cr}_$TAG.err
done
for n in 4 2; do
  [ $n -lt $N ] || continue
  timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29522 \
    tools/bench_large.py 32768 3 > $O/bench_large_${n}gpu_$TAG.json 2> $O/bench_large_${n}gpu_$TAG.err
  echo "N=$n rc=$?"; cut -c1-300 $O/bench_large_${n}gpu_$TAG.json
done
