mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extract_window -c 1 -f -o gpurun_out/r2_extract_window python tools/prof_round2.py --what extract --paths 1024 > gpurun_out/r2_ncu_extract.log 2>&1
tail -3 gpurun_out/r2_ncu_extract.log
