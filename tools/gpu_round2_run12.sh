mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_window.py -x -q -k "extraction or checkpoints" 2>&1 | tail -15
timeout 900 python tools/exp_round2.py --find "" --extract 1 --extract-plain 0 2>&1 | grep -v image_s | cut -c1-400
GBWT_B200_EXTRACT_WINDOW=0 timeout 900 python tools/exp_round2.py --find "" --extract 1 --extract-plain 0 2>&1 | grep '"extract": "checkpointed"' | cut -c1-300
