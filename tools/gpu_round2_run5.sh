set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py tests/test_gpu_fullsize.py -x -q > gpurun_out/r2e_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2e_tests.log
timeout 900 python tools/exp_round2.py --extract 0 --find "v3b" > gpurun_out/r2e_exp.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_find_window -c 2 -f -o gpurun_out/r2_find_window_cur python tools/prof_round2.py --what find64,find32 --queries 67108864 > gpurun_out/r2_ncu_find.log 2>&1
tail -15 gpurun_out/r2e_tests.log; cat gpurun_out/r2e_exp.log
