set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2d_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_tests.log
timeout 900 python tools/exp_round2.py --extract 0 --find "v3,v3pf:WINDOW_PREFETCH=1,v3m64:WINDOW_MARGIN=64" > gpurun_out/r2d_exp.log 2>&1
tail -15 gpurun_out/r2d_tests.log; cat gpurun_out/r2d_exp.log
