mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_window.py -x -q -k "extraction or checkpoints" 2>&1 | tail -2
for p in 256 128; do
    echo "paths $p default"
    timeout 900 python tools/exp_round2.py --find "" --extract 1 --extract-plain 0 --paths $p 2>&1 | grep '"extract": "checkpointed"' | cut -c1-200
done
echo "paths 256 threads 512"; GBWT_B200_EXTRACT_WINDOW_THREADS=512 timeout 900 python tools/exp_round2.py --find "" --extract 1 --extract-plain 0 --paths 256 2>&1 | grep '"extract": "checkpointed"' | cut -c1-200
echo "paths 128 threads 256"; GBWT_B200_EXTRACT_WINDOW_THREADS=256 timeout 900 python tools/exp_round2.py --find "" --extract 1 --extract-plain 0 --paths 128 2>&1 | grep '"extract": "checkpointed"' | cut -c1-200
