set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_full.log
tail -5 gpurun_out/r2_gputests_full.log
# memcheck over the small GPU tests of the new kernels (window search incl. wide records, run checkpoints, DENSE4)
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_window.py tests/test_gpu_runs.py -x -q -k "not workload and not full" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
tail -8 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_runs.py -x -q -k "window_kernel_on_wide_records" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
tail -8 gpurun_out/r2_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_runs.py -x -q -k "window_kernel_on_wide_records" > gpurun_out/r2_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/r2_sanitizer_synccheck.log
tail -8 gpurun_out/r2_sanitizer_synccheck.log
