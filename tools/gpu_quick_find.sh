mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_window.py -x -q 2>&1 | tail -2
timeout 900 python tools/exp_round2.py --extract 0 --find "${1:-cur}" 2>&1 | grep variant | cut -c1-60,330-700
