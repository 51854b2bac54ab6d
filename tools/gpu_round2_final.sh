set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_full.log
tail -3 gpurun_out/r2_gputests_full.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_find_window -c 2 -f -o gpurun_out/r2_find_window_cur python tools/prof_round2.py --what find64,find32 --queries 67108864 > gpurun_out/r2_ncu_find.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extract_window -c 1 -f -o gpurun_out/r2_extract_window python tools/prof_round2.py --what extract --paths 1024 > gpurun_out/r2_ncu_extract.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_b.log 2>&1
bash tools/gpu_round2_bench.sh 1
