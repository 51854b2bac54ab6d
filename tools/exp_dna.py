"""Development experiment: DNA-level extraction vs node extraction on one GPU (device-resident I/O)."""
import argparse, json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gbwt_rs_b200 as gb
from synth import synth

ap = argparse.ArgumentParser()
ap.add_argument("--sites", type=int, default=3_333_333)
ap.add_argument("--haplotypes", type=int, default=1024)
ap.add_argument("--layout", default="auto")
ap.add_argument("--max-anchor", type=int, default=32)
ap.add_argument("--reps", type=int, default=2)
ap.add_argument("--skip-nodes", action="store_true")
ap.add_argument("--nodes-only", action="store_true")
ap.add_argument("--reverse", action="store_true", help="walk the reverse-strand sequences")
args = ap.parse_args()
S, H = args.sites, args.haplotypes
img = synth.bubble_chain(S, H, 42)
index = gb.GBWT.from_bytes(img.array, layout=args.layout)
starts, labels = synth.node_labels(3 * S + 1, seed=42, max_anchor=args.max_anchor)
import time
t_attach = time.time()
index.attach_graph(starts, labels)
print(json.dumps({"op": "attach_graph (with the DNA positions of the checkpoints when the index has checkpoints)", "s": time.time() - t_attach,
                  "checkpoints": index.checkpoint_info()}), flush=True)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, reps=args.reps):
    fn(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


m, L = H, 2 * S + 1
ids = torch.arange(m, dtype=torch.int64, device=dev) * 2 + (1 if args.reverse else 0)
lens = torch.empty(m, dtype=torch.int64, device=dev)
if not args.skip_nodes:
    offs = torch.arange(m + 1, dtype=torch.int64, device=dev) * L
    nodes = torch.empty(m * L, dtype=torch.int64, device=dev)
    ms = timed(lambda: index.extract_device(ids.data_ptr(), m, offs.data_ptr(), nodes.data_ptr(), lens.data_ptr(), stream))
    print(json.dumps({"op": "extract nodes", "ms": ms, "steps_per_s": m * L / ms * 1e3}), flush=True)
    del nodes
ms = timed(lambda: index.sequence_lengths_device(ids.data_ptr(), m, lens.data_ptr(), stream))
print(json.dumps({"op": "sequence lengths (no output)", "ms": ms, "steps_per_s": m * L / ms * 1e3}), flush=True)
if args.nodes_only:
    sys.exit(0)
ms = timed(lambda: index.dna_lengths_device(ids.data_ptr(), m, lens.data_ptr(), stream))
print(json.dumps({"op": "dna lengths", "ms": ms, "steps_per_s": m * L / ms * 1e3}), flush=True)
offs = torch.zeros(m + 1, dtype=torch.int64, device=dev)
offs[1:] = torch.cumsum(lens, 0)
total = int(offs[-1].item())
out = torch.empty(total, dtype=torch.uint8, device=dev)
for ctas in ("1", "3", "4"):   # (read at every launch; the last one is the default)
    os.environ["GBWT_B200_DNA_CTAS"] = ctas
    ms = timed(lambda: index.extract_dna_device(ids.data_ptr(), m, 0, offs.data_ptr(), out.data_ptr(), lens.data_ptr(), stream))
    print(json.dumps({"op": "extract dna", "GBWT_B200_DNA_CTAS": ctas, "ms": ms, "bases_per_s": total / ms * 1e3}), flush=True)
print(json.dumps({"op": "extract dna", "ms": ms, "steps_per_s": m * L / ms * 1e3, "bases_per_s": total / ms * 1e3,
                  "bases_per_node": total / (m * L)}), flush=True)
def checksum_of(t, piece=1 << 28):   # (summed in pieces: a 60 GB byte tensor does not widen in one go)
    return sum(int(t[i:i + piece].sum(dtype=torch.int64).item()) for i in range(0, t.numel(), piece))


checksum = checksum_of(out)
probe = out[:: max(1, total // (1 << 20))].clone()
if index.checkpoint_info()["present"]:
    os.environ["GBWT_B200_DNA_CHECKPOINTS"] = "0"   # the same index by whole-sequence walks
    out.zero_()
    ms = timed(lambda: index.extract_dna_device(ids.data_ptr(), m, 0, offs.data_ptr(), out.data_ptr(), lens.data_ptr(), stream), 1)
    print(json.dumps({"op": "extract dna, whole-sequence walks", "ms": ms, "bases_per_s": total / ms * 1e3,
                      "same_bytes": checksum_of(out) == checksum and bool(torch.equal(out[:: max(1, total // (1 << 20))], probe))}), flush=True)
