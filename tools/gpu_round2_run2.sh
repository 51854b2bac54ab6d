set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2b_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_tests.log
timeout 900 python tools/exp_round2.py --extract 0 --find "win512,win512pf:WINDOW_PREFETCH=1,win256:WINDOW_THREADS=256,win1024:WINDOW_THREADS=1024" > gpurun_out/r2b_exp.log 2>&1
tail -3 gpurun_out/r2b_tests.log; cat gpurun_out/r2b_exp.log
