mkdir -p gpurun_out
N=${1:-1}
if [ "$N" = "1" ]; then
  timeout 1200 python bench.py > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
else
  timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?"
fi
tail -5 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r2_bench_n$N.json").read().strip().splitlines()[-1])
print("value", d["value"]/1e9, "ms", d["ms_per_step"], "e2e", d["e2e"]["value"]/1e9, "u64 e2e", d["e2e"]["u64"]["value"]/1e9, "wire", d["e2e"]["wire_bound_queries_per_s"]/1e9)
print("u32", d["extra"]["find_u32"]["value"]/1e9, "build_s", d["extra"]["index_build_s"], "bd", d["extra"]["bd_search"] and (d["extra"]["bd_search"]["value"]/1e9, d["extra"]["bd_search"]["lf_steps_per_s"]/1e9))
x=d["extra"]["extract"]; print("extract warm", x["warm_lf_steps_per_s"]/1e9, "cold", x["cold_lf_steps_per_s"]/1e9, x["frac_bytes"], x["frac_latency_hbm"])
r=d["extra"]["find_runs"]; print("runs", r["value"]/1e9, r["ms_per_step"], r["deferred_queries_per_step"], r["window_kernel_ms"])
print("roofline", {k:v for k,v in (d["roofline"] or {}).items() if k in ("traffic","frac","compulsory_frac","compulsory_frac_kernel_only","launch_ms","step_ms")})
print("cpu", d["cpu_baseline"] and d["cpu_baseline"]["value"], d["clocks"])
PY
