#!/bin/bash
# One GPU-box round: parity tests, smoke, bench, ncu launch list, ncu full captures (CSV exports) of the predict kernels.
# usage (under gpurun): bash tools/gpu_round.sh [tag] [what...]   what: tests smoke bench launches ncu large configs (default: all but large/configs)
TAG=${1:-r02}; shift
WHAT=${*:-tests smoke bench launches ncu}
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv > $O/smi_$TAG.csv 2>&1
for w in $WHAT; do case $w in
tests) timeout 1500 python -m pytest tests -m gpu -x -q > $O/pytest_gpu_$TAG.log 2>&1; echo "pytest rc=$?" >> $O/pytest_gpu_$TAG.log; tail -3 $O/pytest_gpu_$TAG.log;;
smoke) timeout 300 python __graft_entry__.py --smoke > $O/smoke_$TAG.log 2>&1; echo "smoke rc=$?" >> $O/smoke_$TAG.log; tail -2 $O/smoke_$TAG.log;;
bench) timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_$TAG.json 2> $O/bench_$TAG.err; echo "bench rc=$?" >> $O/bench_$TAG.err; cut -c1-2500 $O/bench_$TAG.json; tail -2 $O/bench_$TAG.err;;
launches) timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file $O/launches_$TAG.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extra > $O/bench_under_ncu_$TAG.log 2>&1;;
ncu) bash tools/ncu_capture.sh var_$TAG 'gp_var_kernel<.int.32, .int.1, .int.12, .*bool.1, .bool.1>' 1 python tools/prof_predict.py 4096 256 600 1 > /dev/null 2>&1
     bash tools/ncu_capture.sh vartail_$TAG 'gp_var_kernel<.int.32, .int.4, .int.3, .*bool.1, .bool.1>' 1 python tools/prof_predict.py 4096 256 600 1 > /dev/null 2>&1
     bash tools/ncu_capture.sh fit_$TAG 'gp_fit_kernel<.int.2, .int.11, .int.3>' 1 python tools/prof_predict.py 4096 256 600 1 > /dev/null 2>&1
     bash tools/ncu_capture.sh var32_$TAG 'gp_var32_kernel' 1 python tools/prof_predict.py 4096 256 600 1 rbf+stdperiodic f32 > /dev/null 2>&1
     bash tools/ncu_capture.sh lookahead_$TAG 'zupt_lookahead_tc_kernel' 1 python tools/bench_configs.py mc 0.125 > /dev/null 2>&1
     bash tools/ncu_capture.sh grad_$TAG 'gp_grad_kernel' 2 python tools/bench_configs.py sweep 0.125 > /dev/null 2>&1;;
large) timeout 600 python tools/bench_large.py 32768 3 > $O/bench_large_$TAG.json 2> $O/bench_large_$TAG.err;;
configs) timeout 900 python tools/bench_configs.py all > $O/bench_configs_$TAG.json 2> $O/bench_configs_$TAG.err;;
esac; done
