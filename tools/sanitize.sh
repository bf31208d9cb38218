#!/bin/bash
# compute-sanitizer passes over a small slice of the GPU parity tests (run under gpurun): memcheck over every kernel
# family, racecheck over the shared-memory protocols of the fit / variance / large-window kernels.
O=gpurun_out; mkdir -p $O
SEL='test_predict_matches_oracle_shapes or test_lml_grad_matches_oracle or test_lookahead_per_window or test_matches_oracle or test_pipelined'
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/sanitizer_memcheck.log \
  python -m pytest tests/test_gpu_predict.py tests/test_gpu_lookahead.py tests/test_gpu_slip_record.py tests/test_gpu_ekf_context.py \
  -q -x -k "$SEL" > $O/sanitizer_memcheck.out 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck.out
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/sanitizer_racecheck.log \
  python -m pytest tests/test_gpu_predict.py -q -x -k "test_predict_matches_oracle_shapes" > $O/sanitizer_racecheck.out 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck.out
tail -3 $O/sanitizer_memcheck.out; tail -5 $O/sanitizer_memcheck.log; tail -3 $O/sanitizer_racecheck.out; tail -8 $O/sanitizer_racecheck.log
