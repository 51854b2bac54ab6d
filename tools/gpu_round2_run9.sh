mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_window.py tests/test_gpu_runs.py tests/test_gpu_fullsize.py -x -q 2>&1 | tail -3
GBWT_B200_WINDOW_SORT=0 timeout 1200 python -m pytest tests/test_gpu_window.py -x -q 2>&1 | tail -2
timeout 900 python tools/exp_round2.py --extract 0 --find "newsort,oldsort:WINDOW_SORT=0" 2>&1 | grep variant | cut -c1-60,330-700
