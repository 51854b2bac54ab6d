mkdir -p gpurun_out
timeout 900 python tools/exp_round2.py --find "" --extract 1 --extract-plain 0 2>&1 | grep '"extract"' | cut -c1-300
