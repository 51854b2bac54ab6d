set -x
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2_gputests_full.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_gputests_full.log
tail -3 gpurun_out/r2_gputests_full.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1 | cut -c1-200
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python -m pytest tests/test_gpu_dna.py -x -q -k "checkpointed" > gpurun_out/r2_sanitizer_dna_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2_sanitizer_dna_memcheck.log
tail -3 gpurun_out/r2_sanitizer_dna_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python -m pytest tests/test_gpu_dna.py -x -q -k "checkpointed_dna_on_random" > gpurun_out/r2_sanitizer_dna_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2_sanitizer_dna_racecheck.log
tail -3 gpurun_out/r2_sanitizer_dna_racecheck.log
timeout 400 python tools/exp_dna.py --skip-nodes > gpurun_out/r2_dna.log 2>&1; cut -c1-260 gpurun_out/r2_dna.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_find_window -c 2 -f -o gpurun_out/r2_find_window_cur python tools/prof_round2.py --what find64,find32 --queries 67108864 > gpurun_out/r2_ncu_find.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_extract_dna_checkpointed --launch-skip 1 -c 1 -f -o gpurun_out/r2_extract_dna_ckpt python tools/exp_dna.py --skip-nodes --haplotypes 256 > gpurun_out/r2_ncu_dna.log 2>&1; tail -2 gpurun_out/r2_ncu_dna.log
bash tools/gpu_round2_bench.sh 1
