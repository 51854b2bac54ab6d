"""One launch each of the round-2 kernels on the config-4 index, for ncu (see profiles/README.md for the command).
Usage (GPU box, under ncu): python tools/prof_round2.py [--queries N] [--paths M]"""
import argparse, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import gbwt_rs_b200 as gb
from synth import synth

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1 << 25)
ap.add_argument("--paths", type=int, default=256)
ap.add_argument("--sites", type=int, default=3_333_333)
ap.add_argument("--haplotypes", type=int, default=1024)
ap.add_argument("--what", default="find64,find32,extract")
args = ap.parse_args()
S, H, Q = args.sites, args.haplotypes, args.queries
img = synth.bubble_chain(S, H, 42)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream
what = args.what.split(",")
index = gb.GBWT.from_bytes(img.array, checkpoints="extract" in what)
if "find64" in what or "find32" in what or "bd" in what:
    d_pat = torch.empty((Q, 32), dtype=torch.int64, device=dev)
    d_out = torch.empty((Q, 3), dtype=torch.int64, device=dev)
    synth.patterns_device(S, H, 42, Q, d_pat.data_ptr(), stream=stream)
    torch.cuda.synchronize()
    if "find64" in what:
        index.find_extend_device(d_pat.data_ptr(), Q, 32, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
    if "find32" in what:
        d_pat32 = d_pat.to(torch.int32)
        index.find_extend_u32_device(d_pat32.data_ptr(), Q, 32, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
    if "bd" in what:  # the bench's bidirectional searches: a random anchor inside every pattern, random ends either side
        gen = torch.Generator(device=dev); gen.manual_seed(11)
        first = torch.randint(0, 32, (Q,), device=dev, generator=gen, dtype=torch.int64)
        start = (first.double() * torch.rand(Q, device=dev, generator=gen, dtype=torch.float64)).long()
        end = first + 1 + ((31 - first).double() * torch.rand(Q, device=dev, generator=gen, dtype=torch.float64)).long()
        offs = torch.arange(Q + 1, dtype=torch.int64, device=dev) * 32
        d_bd = torch.empty((Q, 6), dtype=torch.int64, device=dev)
        index.bd_search_device(d_pat.data_ptr(), offs.data_ptr(), first.data_ptr(), start.data_ptr(), end.data_ptr(), Q, d_bd.data_ptr(), stream)
        torch.cuda.synchronize()
        assert bool(torch.all(d_bd[:, 2] > d_bd[:, 1]).item())
        del first, start, end, offs, d_bd
    del d_pat, d_out
if "extract" in what:
    m, length = args.paths, 2 * S + 1
    ids = torch.arange(0, m, dtype=torch.int64, device=dev) * 2
    offs = torch.arange(m + 1, dtype=torch.int64, device=dev) * length
    nodes = torch.empty(m * length, dtype=torch.int64, device=dev)
    lens = torch.empty(m, dtype=torch.int64, device=dev)
    index.extract_device(ids.data_ptr(), m, offs.data_ptr(), nodes.data_ptr(), lens.data_ptr(), stream)
    torch.cuda.synchronize()
    assert bool(torch.all(lens == length).item())
print("done")
