#!/bin/bash
# Build in-tree (the .so travels with the snapshot) and run a command on the GPU box.
# usage: tools/gpu.sh <timeout_s> '<command>'
set -e
cd "$(dirname "$0")/.."
python -m corenav_gp_b200.build > /dev/null
python -c "from oracle import stop_oracle, slip_oracle; stop_oracle.build(); slip_oracle.build()" > /dev/null
T=$1; shift
exec /usr/local/graft/bin/gpurun --timeout "$T" -- "$@"
