set -x
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_runs.py tests/test_gpu_window.py -x -q > gpurun_out/r2h_tests.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2h_tests.log
tail -25 gpurun_out/r2h_tests.log
timeout 600 python bench.py --workload find-runs --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2h_runs.json 2> gpurun_out/r2h_runs.err
timeout 600 python bench.py --steps 5 --warmup 2 --no-e2e --no-cpu-baseline --no-extract > gpurun_out/r2h_find.json 2> gpurun_out/r2h_find.err
python - <<'PY'
import json
for f in ("r2h_runs","r2h_find"):
    try:
        d=json.load(open(f"gpurun_out/{f}.json")); print(f, d["value"]/1e9, d["ms_per_step"], d["extra"]["find_u32"]["value"]/1e9, d["extra"]["index_device_bytes"], d["roofline"]["deferred_queries_per_step"], d["roofline"]["launch_ms"])
    except Exception as e: print(f, "failed", e); print(open(f"gpurun_out/{f}.err").read()[-1500:])
PY
