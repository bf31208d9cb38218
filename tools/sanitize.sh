#!/bin/bash
# compute-sanitizer passes over a slice of the GPU parity tests (run under gpurun): memcheck over every kernel family
# (round 2: lag-table fit / variance, FP32-mode variance, tensor-core look-ahead, shared-memory gradient, device
# optimiser, EKF covariance), racecheck over the shared-memory protocols of the fit / variance / look-ahead kernels.
TAG=${1:-r02}
O=gpurun_out; mkdir -p $O
SEL='test_predict_matches_oracle_shapes or test_lml_grad_matches_oracle or test_lookahead_per_window or test_pipelined or test_table_path_matches_oracle_and_lazy_path or test_mixed_batch or test_fp32_mode_within_1e4 or test_fp32_mode_interpreter_path or test_tensor_core_kernel_many or test_final_state or test_ekf_covariance or test_optimize_on_device or test_jitter_ladder'
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 --log-file $O/sanitizer_memcheck_$TAG.log \
  python -m pytest tests/test_gpu_predict.py tests/test_gpu_lookahead.py tests/test_gpu_slip_record.py tests/test_gpu_ekf_context.py \
  tests/test_gpu_lag_tables.py tests/test_gpu_fp32_mode.py tests/test_gpu_fit_callback.py tests/test_gpu_parity_report.py \
  -q -x -m gpu -k "$SEL" > $O/sanitizer_memcheck_$TAG.out 2>&1; echo "memcheck rc=$?" >> $O/sanitizer_memcheck_$TAG.out
timeout 1200 compute-sanitizer --tool racecheck --error-exitcode 9 --log-file $O/sanitizer_racecheck_$TAG.log \
  python -m pytest tests/test_gpu_lookahead.py tests/test_gpu_lag_tables.py -q -x -m gpu \
  -k "test_lookahead_per_window or test_final_state or test_reference_grid_overlapping" > $O/sanitizer_racecheck_$TAG.out 2>&1; echo "racecheck rc=$?" >> $O/sanitizer_racecheck_$TAG.out
tail -3 $O/sanitizer_memcheck_$TAG.out; tail -5 $O/sanitizer_memcheck_$TAG.log; tail -3 $O/sanitizer_racecheck_$TAG.out; tail -12 $O/sanitizer_racecheck_$TAG.log
