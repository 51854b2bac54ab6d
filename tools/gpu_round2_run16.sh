GBWT_B200_WINDOW_STATS=0 timeout 900 python tools/exp_round2.py --extract 0 --find "hint,hintpf:WINDOW_PREFETCH=1" 2>&1 | grep variant | cut -c1-60,330-900
