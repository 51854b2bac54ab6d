"""Development experiment: bidirectional search and extraction throughput on one GPU (device-resident I/O)."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gbwt_rs_b200 as gb
from synth import synth

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1 << 24)
ap.add_argument("--sites", type=int, default=3_333_333)
ap.add_argument("--haplotypes", type=int, default=1024)
ap.add_argument("--layout", default="auto")
ap.add_argument("--paths", type=int, default=0, help="number of forward paths to extract (0 = all)")
args = ap.parse_args()
S, H, Q = args.sites, args.haplotypes, args.queries
img = synth.bubble_chain(S, H, 42)
index = gb.GBWT.from_bytes(img.array, layout=args.layout)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream

def timed(fn, reps=3):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps

# bidirectional: bd_find(p[16]) + 15 forward + 16 backward extensions per pattern
d_pat = torch.empty((Q, 32), dtype=torch.int64, device=dev)
synth.patterns_device(S, H, 42, Q, d_pat.data_ptr(), stream=stream)
offs = torch.arange(Q + 1, dtype=torch.int64, device=dev) * 32
first = torch.full((Q,), 16, dtype=torch.int64, device=dev)
start = torch.zeros(Q, dtype=torch.int64, device=dev)
end = torch.full((Q,), 32, dtype=torch.int64, device=dev)
out = torch.zeros((Q, 6), dtype=torch.int64, device=dev)
ms = timed(lambda: index.bd_search_device(d_pat.data_ptr(), offs.data_ptr(), first.data_ptr(), start.data_ptr(), end.data_ptr(), Q, out.data_ptr(), stream))
ok = bool(torch.all(out[:, 2] > out[:, 1]).item()) and bool(torch.all((out[:, 2] - out[:, 1]) == (out[:, 5] - out[:, 4])).item())
print(json.dumps({"op": "bd_search k=32 first=16", "ms": ms, "mqps": Q / ms / 1e3, "steps_per_s": Q * 32 / ms * 1e3, "ok": ok}), flush=True)
del d_pat, offs, first, start, end, out

m = args.paths or H
L = 2 * S + 1
ids = torch.arange(m, dtype=torch.int64, device=dev) * 2
offs = torch.arange(m + 1, dtype=torch.int64, device=dev) * L
nodes = torch.empty(m * L, dtype=torch.int64, device=dev)
lens = torch.empty(m, dtype=torch.int64, device=dev)
ms = timed(lambda: index.extract_device(ids.data_ptr(), m, offs.data_ptr(), nodes.data_ptr(), lens.data_ptr(), stream), reps=2)
print(json.dumps({"op": f"extract {m} paths", "ms": ms, "steps_per_s": m * L / ms * 1e3, "ok": bool(torch.all(lens == L).item())}), flush=True)
for ahead in ("0", "12", "48", "192", "768"):
    os.environ["GBWT_B200_EXTRACT_AHEAD"] = ahead
    ms = timed(lambda: index.extract_device(ids.data_ptr(), m, offs.data_ptr(), nodes.data_ptr(), lens.data_ptr(), stream), reps=1)
    print(json.dumps({"op": f"extract {m} paths, ahead {ahead}", "ms": ms, "steps_per_s": m * L / ms * 1e3}), flush=True)
