mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_window.py tests/test_gpu_runs.py -x -q 2>&1 | tail -12
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
timeout 900 python tools/exp_round2.py --extract 0 --find "cur" 2>&1 | grep variant | cut -c1-60,330-700
