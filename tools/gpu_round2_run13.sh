mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_full.log
tail -5 gpurun_out/r2_gputests_full.log
