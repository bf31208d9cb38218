#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list, ncu full captures (CSV exports) of the two predict kernels, large-window bench.
# usage (under gpurun): bash tools/gpu_round.sh [tag]
TAG=${1:-r01}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_$TAG.csv 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log
timeout 300 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?" >> $O/bench_$TAG.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > $O/bench_under_ncu_$TAG.log 2>&1
bash tools/ncu_capture.sh var_$TAG gp_var python tools/prof_predict.py 4096 256 600 1 > /dev/null 2>&1
bash tools/ncu_capture.sh fit_$TAG gp_fit python tools/prof_predict.py 4096 256 600 1 > /dev/null 2>&1
timeout 600 python tools/bench_large.py 32768 3 > $O/bench_large_$TAG.json 2> $O/bench_large_$TAG.err
timeout 600 python tools/bench_configs.py all > $O/bench_configs_$TAG.json 2> $O/bench_configs_$TAG.err
tail -3 $O/pytest_gpu_$TAG.log; tail -2 $O/smoke_$TAG.log; cat $O/bench_$TAG.json | cut -c1-1500; tail -2 $O/bench_$TAG.err
