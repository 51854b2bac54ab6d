set -x
mkdir -p gpurun_out
GBWT_B200_WINDOW_FINE=4 timeout 900 python -m pytest tests/test_gpu_window.py -x -q > gpurun_out/r2c_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_tests.log
timeout 900 python tools/exp_round2.py --extract 0 --find "fine0,fine2:WINDOW_FINE=2,fine4:WINDOW_FINE=4,fine5:WINDOW_FINE=5,fine7:WINDOW_FINE=7" > gpurun_out/r2c_exp.log 2>&1
tail -3 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_exp.log
