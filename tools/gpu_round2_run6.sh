set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_runs.py -x -q > gpurun_out/r2f_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_tests.log
tail -15 gpurun_out/r2f_tests.log
timeout 600 python bench.py --workload find-runs --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2f_runs_ckpt.json 2> gpurun_out/r2f_runs_ckpt.err
GBWT_B200_CKPT=0 timeout 600 python bench.py --workload find-runs --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2f_runs_nockpt.json 2> gpurun_out/r2f_runs_nockpt.err
tail -3 gpurun_out/r2f_runs_ckpt.err gpurun_out/r2f_runs_nockpt.err
python - <<'PY'
import json
for f in ("r2f_runs_ckpt","r2f_runs_nockpt"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"]/1e9, d["ms_per_step"], d["extra"]["find_u32"]["value"]/1e9, d["extra"]["index_device_bytes"], d["roofline"]["deferred_queries_per_step"])
    except Exception as e: print(f, "failed", e)
PY
