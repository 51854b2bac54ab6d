set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_full.log
tail -3 gpurun_out/r2_gputests_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_find_window -c 2 -f -o gpurun_out/r2_find_window_cur python tools/prof_round2.py --what find64,find32 --queries 67108864 > gpurun_out/r2_ncu_find.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_extract_window -c 1 -f -o gpurun_out/r2_extract_window python tools/prof_round2.py --what extract --paths 1024 > gpurun_out/r2_ncu_extract.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"k_bd_window|k_bd_place_direct" -c 2 -f -o gpurun_out/r2_bd_window_final python tools/prof_round2.py --what bd --queries 67108864 > gpurun_out/r2_ncu_bd.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline --no-e2e > gpurun_out/r2_b.log 2>&1
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_window.py tests/test_gpu_runs.py -x -q -k "not workload and not full and not many_sequences and not imported and not bubble_chains" > gpurun_out/r2_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_memcheck.log
tail -3 gpurun_out/r2_sanitizer_memcheck.log
timeout 1500 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_runs.py tests/test_gpu_window.py -x -q -k "window_kernel_on_wide_records or checkpointed_extraction_matches_chain_walks or pattern_lengths or bd_window_kernel_on_fixtures" > gpurun_out/r2_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_racecheck.log
tail -3 gpurun_out/r2_sanitizer_racecheck.log
timeout 900 compute-sanitizer --tool synccheck --error-exitcode 7 python -m pytest tests/test_gpu_runs.py tests/test_gpu_window.py -x -q -k "window_kernel_on_wide_records or checkpointed_extraction_matches_chain_walks or bd_window_kernel_on_fixtures" > gpurun_out/r2_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?" >> gpurun_out/r2_sanitizer_synccheck.log
tail -3 gpurun_out/r2_sanitizer_synccheck.log
bash tools/gpu_round2_bench.sh 1
