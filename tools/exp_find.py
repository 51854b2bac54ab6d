"""Development experiment: times the find/extend kernel variants and L2 fetch granularities on one GPU.
Usage (GPU box): python tools/exp_find.py [--queries N]"""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gbwt_rs_b200 as gb
from synth import synth

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1 << 25)
ap.add_argument("--sites", type=int, default=3_333_333)
ap.add_argument("--haplotypes", type=int, default=1024)
ap.add_argument("--variants", default="3")
ap.add_argument("--locality", default="0,1")
ap.add_argument("--minblocks", default="1")
ap.add_argument("--l2", default="32")
ap.add_argument("--layout", default="auto")
args = ap.parse_args()
S, H, Q = args.sites, args.haplotypes, args.queries
img = synth.bubble_chain(S, H, 42)
dev = torch.device("cuda", 0)
d_pat = torch.empty((Q, 32), dtype=torch.int64, device=dev)
d_out = torch.empty((Q, 3), dtype=torch.int64, device=dev)
stream = torch.cuda.current_stream().cuda_stream
synth.patterns_device(S, H, 42, Q, d_pat.data_ptr(), stream=stream)
torch.cuda.synchronize()
ref = None
for l2 in args.l2.split(","):
    os.environ["GBWT_B200_L2_FETCH"] = l2
    index = gb.GBWT.from_bytes(img.array, layout=args.layout)
    for v, loc, mb in [(v, loc, mb) for loc in args.locality.split(",") for v in args.variants.split(",") for mb in args.minblocks.split(",")]:
        os.environ["GBWT_B200_MINBLOCKS"] = mb
        os.environ["GBWT_B200_FIND_VARIANT"] = v
        os.environ["GBWT_B200_LOCALITY"] = loc
        for _ in range(2):
            index.find_extend_device(d_pat.data_ptr(), Q, 32, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            index.find_extend_device(d_pat.data_ptr(), Q, 32, d_out.data_ptr(), stream)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        chk = int((d_out[:, 2] - d_out[:, 1]).sum().item())
        if ref is None: ref = chk
        print(json.dumps({"locality": loc, "variant": v, "minblocks": mb, "ms": ms, "mqps": Q / ms / 1e3, "checksum_ok": chk == ref}), flush=True)
    del index
