mkdir -p gpurun_out
for v in "w0_ckpt:GBWT_B200_FIND_WINDOW=0" "w0_nockpt:GBWT_B200_FIND_WINDOW=0:GBWT_B200_CKPT=0" "w0_general:GBWT_B200_FIND_WINDOW=0:GBWT_B200_FIND_LEAN=0"; do
  name=${v%%:*}; envs=$(echo ${v#*:} | tr ':' ' ')
  env $envs timeout 600 python bench.py --workload find-runs --steps 5 --warmup 2 --no-e2e --no-cpu-baseline > gpurun_out/r2g_$name.json 2> gpurun_out/r2g_$name.err
  python - <<PY
import json
try:
    d=json.load(open("gpurun_out/r2g_$name.json")); print("$name", d["value"]/1e9, d["ms_per_step"], d["extra"]["find_u32"]["value"]/1e9)
except Exception as e: print("$name", "failed", e)
PY
done
