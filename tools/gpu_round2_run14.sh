mkdir -p gpurun_out
GBWT_B200_EXTRACT_WINDOW_THREADS=512 timeout 900 python tools/exp_round2.py --find "" --extract 1 --extract-plain 0 2>&1 | grep '"extract"' | cut -c1-300
