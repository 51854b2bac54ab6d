"""Development experiment (round 2): times the record-window search kernel against the scattered-load kernels, for
both pattern widths and a few window shapes, and the checkpointed extraction against the chain walks, on one GPU.
Usage (GPU box): python tools/exp_round2.py [--queries N] [--find VARIANTS] [--extract 0|1]
A find variant is name[:KNOB=value[:KNOB=value...]] with GBWT_B200_ knobs read when the index is created."""
import argparse, json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import gbwt_rs_b200 as gb
from synth import synth

ap = argparse.ArgumentParser()
ap.add_argument("--queries", type=int, default=1 << 26)
ap.add_argument("--sites", type=int, default=3_333_333)
ap.add_argument("--haplotypes", type=int, default=1024)
ap.add_argument("--find", default="old:FIND_WINDOW=0,win512,win256:WINDOW_THREADS=256,win1024:WINDOW_THREADS=1024")
ap.add_argument("--extract", type=int, default=1)
ap.add_argument("--bd", type=int, default=0, help="also time the bench's bidirectional searches with every find variant")
ap.add_argument("--extract-plain", type=int, default=1)
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--ckpt-shifts", default="")
ap.add_argument("--extract-threads", default="256")
ap.add_argument("--paths", type=int, default=0, help="forward paths to extract (default: all)")
args = ap.parse_args()
S, H, Q = args.sites, args.haplotypes, args.queries
t = time.time()
img = synth.bubble_chain(S, H, 42)
print(json.dumps({"image_s": time.time() - t}), flush=True)
dev = torch.device("cuda", 0)
stream = torch.cuda.current_stream().cuda_stream


def timed(fn, reps):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


if args.find:
    d_pat = torch.empty((Q, 32), dtype=torch.int64, device=dev)
    d_out = torch.empty((Q, 3), dtype=torch.int64, device=dev)
    synth.patterns_device(S, H, 42, Q, d_pat.data_ptr(), stream=stream)
    d_pat32 = d_pat.to(torch.int32)
    torch.cuda.synchronize()
    ref = None
    for variant in args.find.split(","):
        name, *knobs = variant.split(":")
        set_knobs = []
        for kv in knobs:
            k, v = kv.split("=")
            os.environ["GBWT_B200_" + k] = v
            set_knobs.append("GBWT_B200_" + k)
        os.environ["GBWT_B200_WINDOW_STATS"] = "0"
        t = time.time()
        index = gb.GBWT.from_bytes(img.array, checkpoints=False)
        build_s = time.time() - t
        row = {"variant": variant, "build_s": build_s, "window": index.window_info()}
        for width, pat in (("u64", d_pat), ("u32", d_pat32)):
            fn = (lambda: index.find_extend_device(pat.data_ptr(), Q, 32, d_out.data_ptr(), stream)) if width == "u64" else \
                 (lambda: index.find_extend_u32_device(pat.data_ptr(), Q, 32, d_out.data_ptr(), stream))
            ms = timed(fn, args.reps)
            chk = int((d_out[:, 2] - d_out[:, 1]).sum().item())
            ok = bool(torch.equal(d_out[:, 0], d_pat[:, 31]))
            if ref is None:
                ref = chk
            row[width] = {"ms": ms, "gqps": Q / ms / 1e6, "checksum_ok": chk == ref and ok}
        os.environ["GBWT_B200_WINDOW_STATS"] = "1"
        index.find_extend_device(d_pat.data_ptr(), Q, 32, d_out.data_ptr(), stream)
        torch.cuda.synchronize()
        info = index.window_info()
        row["deferred"] = info["deferred"]
        if args.bd:
            gen = torch.Generator(device=dev); gen.manual_seed(11)
            first = torch.randint(0, 32, (Q,), device=dev, generator=gen, dtype=torch.int64)
            start = (first.double() * torch.rand(Q, device=dev, generator=gen, dtype=torch.float64)).long()
            end = first + 1 + ((31 - first).double() * torch.rand(Q, device=dev, generator=gen, dtype=torch.float64)).long()
            offs = torch.arange(Q + 1, dtype=torch.int64, device=dev) * 32
            d_bd = torch.empty((Q, 6), dtype=torch.int64, device=dev)
            ms = timed(lambda: index.bd_search_device(d_pat.data_ptr(), offs.data_ptr(), first.data_ptr(), start.data_ptr(), end.data_ptr(), Q,
                                                      d_bd.data_ptr(), stream), args.reps)
            ok = bool(torch.all(d_bd[:, 2] > d_bd[:, 1]).item()) and bool(torch.equal(d_bd[:, 2] - d_bd[:, 1], d_bd[:, 5] - d_bd[:, 4]))
            row["bd"] = {"ms": ms, "g_searches_per_s": Q / ms / 1e6, "sizes_ok": ok, "checksum": int(d_bd.sum().item())}
            del first, start, end, offs, d_bd
        print(json.dumps(row), flush=True)
        del index
        for k in set_knobs:
            del os.environ[k]
    del d_pat, d_out, d_pat32

if args.extract:
    m, length = (args.paths or H), 2 * S + 1
    ids = torch.arange(0, m, dtype=torch.int64, device=dev) * 2
    offs = torch.arange(m + 1, dtype=torch.int64, device=dev) * length
    nodes = torch.empty(m * length, dtype=torch.int64, device=dev)
    lens = torch.empty(m, dtype=torch.int64, device=dev)
    want_first, want_last = synth.sequence(S, H, 42, 0), synth.sequence(S, H, 42, 2 * (m - 1))
    checksum = None
    for shift in (args.ckpt_shifts.split(",") if args.ckpt_shifts else [""]):
        if shift:
            os.environ["GBWT_B200_CHECKPOINT_SHIFT"] = shift
        t = time.time()
        index = gb.GBWT.from_bytes(img.array, checkpoints=True)
        print(json.dumps({"build_with_checkpoints_s": time.time() - t, "checkpoints": index.checkpoint_info()}), flush=True)
        fn = lambda: index.extract_device(ids.data_ptr(), m, offs.data_ptr(), nodes.data_ptr(), lens.data_ptr(), stream)
        for threads in args.extract_threads.split(","):
            os.environ["GBWT_B200_EXTRACT_THREADS"] = threads
            nodes.zero_()
            ms = timed(fn, args.reps)
            ok = bool(torch.all(lens == length).item())
            ok = ok and np.array_equal(nodes[:length].cpu().numpy().view(np.uint64), want_first)
            ok = ok and np.array_equal(nodes[(m - 1) * length:].cpu().numpy().view(np.uint64), want_last)
            chk = int(nodes.sum().item())
            checksum = chk if checksum is None else checksum
            print(json.dumps({"extract": "checkpointed", "shift": shift, "threads": threads, "ms": ms, "g_lf_steps_per_s": m * length / ms / 1e6,
                              "ok": ok, "checksum_ok": chk == checksum}), flush=True)
        del os.environ["GBWT_B200_EXTRACT_THREADS"]
        os.environ["GBWT_B200_EXTRACT_DISCARD"] = "1"
        ms = timed(fn, args.reps)
        del os.environ["GBWT_B200_EXTRACT_DISCARD"]
        print(json.dumps({"extract": "checkpointed, walks only (nodes not stored)", "ms": ms, "g_lf_steps_per_s": m * length / ms / 1e6}), flush=True)
        t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0.record(); nodes.fill_(1); t1.record(); torch.cuda.synchronize()
        print(json.dumps({"plain fill of the output buffer, ms": t0.elapsed_time(t1), "GB/s": nodes.numel() * 8 / t0.elapsed_time(t1) / 1e6}), flush=True)
        if shift != (args.ckpt_shifts.split(",")[-1] if args.ckpt_shifts else ""):
            del index
    if args.extract_plain:
        os.environ["GBWT_B200_EXTRACT_CHECKPOINTS"] = "0"
        nodes.zero_()
        ms = timed(fn, 1)
        print(json.dumps({"extract": "two-ended chains (lengths known)", "ms": ms, "g_lf_steps_per_s": m * length / ms / 1e6,
                          "checksum_ok": int(nodes.sum().item()) == checksum}), flush=True)
        os.environ["GBWT_B200_EXTRACT_SPLIT"] = "0"
        ms = timed(fn, 1)
        print(json.dumps({"extract": "one-ended chains", "ms": ms, "g_lf_steps_per_s": m * length / ms / 1e6,
                          "checksum_ok": int(nodes.sum().item()) == checksum}), flush=True)
