set -x
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_find_window -c 2 -f -o gpurun_out/r2_find_window_cur python tools/prof_round2.py --what find64,find32 --queries 67108864 > gpurun_out/r2_ncu_find.log 2>&1
tail -3 gpurun_out/r2_ncu_find.log
