O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531"
CNGP_LARGE_PHASES=1 timeout 300 $TR tools/bench_large.py 32768 6 > $O/bench_large_8gpu_r02p.json 2> $O/bench_large_8gpu_r02p.err; echo "groups rc=$?"; cut -c1-500 $O/bench_large_8gpu_r02p.json
CNGP_LARGE_PHASES=1 CNGP_LARGE_NO_GROUPS=1 timeout 300 $TR tools/bench_large.py 32768 6 > $O/bench_large_8gpu_nogroups_r02p.json 2> $O/bench_large_8gpu_nogroups_r02p.err; echo "nogroups rc=$?"; cut -c1-500 $O/bench_large_8gpu_nogroups_r02p.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29532 tools/bench_large.py 32768 4 > $O/bench_large_4gpu_r02p.json 2> $O/bench_large_4gpu_r02p.err; echo "4 rc=$?"; cut -c1-400 $O/bench_large_4gpu_r02p.json
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 tools/bench_large.py 32768 4 > $O/bench_large_2gpu_r02p.json 2> $O/bench_large_2gpu_r02p.err; echo "2 rc=$?"; cut -c1-400 $O/bench_large_2gpu_r02p.json
tail -c 300 $O/bench_large_8gpu_r02p.err
