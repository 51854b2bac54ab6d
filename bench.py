#!/usr/bin/env python
"""bench.py -- headline benchmark: batched GBWT find/extend queries per second (k = 32) on B200.

    python bench.py [--gpus N --steps K --warmup W]                      (N = 1)
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
    python bench.py --impl reference ...       the CPU restatement of the reference path on the host cores

Workload (BASELINE.json configs[3]): synthetic chromosome-scale bubble-chain GBWT, 10 M nodes x 1024
haplotypes (3 333 333 sites), length-32 patterns sampled from the haplotypes; the index is
replicated on every GPU and the queries are sharded (weak scaling: 2^26 queries per GPU per step, ~1.07 G
queries per step at 8 GPUs). A step is one pass of the hot path over the rank's batch. One JSON line on stdout.

Other workloads for development: --workload find-c3 (configs[2]), --workload extract (configs[4]).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (sites, haplotypes, queries per GPU per step)
    "find": (3_333_333, 1024, 1 << 26),
    "find-c3": (33_333, 64, 10_000_000),
    "extract": (3_333_333, 1024, 0),
    "extract-dna": (3_333_333, 1024, 0),
}
SEED, SEED_Q, K_LEN = 42, 7, 32


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def parse_args():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="find")
    ap.add_argument("--sites", type=int, default=0, help="override the number of sites (development only)")
    ap.add_argument("--haplotypes", type=int, default=0)
    ap.add_argument("--queries", type=int, default=0, help="queries per GPU per step")
    ap.add_argument("--e2e-queries", type=int, default=1 << 24, help="queries per GPU per end-to-end step")
    ap.add_argument("--layout", choices=["auto", "runs"], default="auto")
    ap.add_argument("--cpu-sample", type=int, default=500_000, help="queries in the CPU baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


def dist_env():
    return int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))


# ---- index image shared between the ranks of one box --------------------------------------------------

def get_image(sites, haplotypes, rank, world, barrier):
    """Rank 0 generates the Simple-SDS GBWT image (into /dev/shm when there are several ranks)."""
    from synth import synth
    if world == 1:
        t = time.time()
        img = synth.bubble_chain(sites, haplotypes, SEED)
        log(f"[bench] generated index image: {img.nbytes / 1e9:.2f} GB in {time.time() - t:.1f} s")
        return img.array, img
    path = f"/dev/shm/gbwt_b200_bench_{sites}_{haplotypes}_{SEED}_{os.environ.get('MASTER_PORT', '0')}.gbwt"
    if rank == 0:
        t = time.time()
        img = synth.bubble_chain(sites, haplotypes, SEED, threads=os.cpu_count() or 0)  # the other ranks wait
        img.array.tofile(path + ".tmp")
        os.replace(path + ".tmp", path)
        log(f"[bench] generated index image: {img.nbytes / 1e9:.2f} GB in {time.time() - t:.1f} s")
        img.free()
    barrier()
    arr = np.fromfile(path, dtype=np.uint8)
    barrier()
    if rank == 0:
        os.unlink(path)
    return arr, None


# ---- clocks ---------------------------------------------------------------------------------------------

class ClockSampler:
    """Samples SM clock, power and throttle reasons through NVML every few milliseconds while the timed region
    runs (the same fields as the nvidia-smi line of B200_PROFILING.md, without the process start-up latency,
    so that even a 100 ms timed region gets samples)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, cuda_index):
        self.samples, self.handle, self.nv = [], None, None
        self._stop = threading.Event()
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            uuid = str(torch.cuda.get_device_properties(cuda_index).uuid)
            try:
                self.handle = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid).encode())
            except Exception:
                self.handle = pynvml.nvmlDeviceGetHandleByIndex(cuda_index)
            self.nv = pynvml
            self.sm_max = pynvml.nvmlDeviceGetMaxClockInfo(self.handle, pynvml.NVML_CLOCK_SM)
            self.thread = threading.Thread(target=self._run, daemon=True)
            self.thread.start()
        except Exception as exc:
            self.error = str(exc)
            self.handle = None

    def _run(self):
        nv, h = self.nv, self.handle
        while not self._stop.is_set():
            try:
                try:
                    mask = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    mask = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM),
                                     nv.nvmlDeviceGetPowerUsage(h) / 1000.0, mask))
            except Exception:
                pass
            time.sleep(0.004)

    def stop(self, t0, t1):
        if self.handle is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["NVML unavailable: " + getattr(self, "error", "?")]}
        self._stop.set()
        self.thread.join(timeout=1.0)
        inside = [s for s in self.samples if t0 <= s[0] <= t1]
        if not inside:
            return {"sm_mhz": None, "sm_max_mhz": float(self.sm_max), "reasons": ["no samples inside the timed region"]}
        reasons = set()
        for s in inside:
            for bit, name in self.REASONS.items():
                if s[3] & bit:
                    reasons.add(name)
        return {"sm_mhz": float(statistics.median(s[1] for s in inside)), "sm_max_mhz": float(self.sm_max),
                "power_w_max": max(s[2] for s in inside), "samples": len(inside), "reasons": sorted(reasons),
                "how": "NVML, 4 ms period, samples with timestamps inside the timed region only"}


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs"
    except Exception:
        return 6650.0, "fallback 6.65 TB/s (B200_PROFILING.md)"


# ---- CPU side (oracle = C restatement of the reference path) ------------------------------------------

def cpu_find_baseline(image, sites, haplotypes, sample, steps=1, warmup=0):
    """Times the oracle (kind "port") on all host threads over `sample` queries per step."""
    from oracle import oracle as orc
    from synth import synth
    g = orc.GBWT.load(image, native=True)
    threads = orc.max_threads()
    pats = synth.patterns(sites, haplotypes, SEED, n=sample, k=K_LEN, seed_q=SEED_Q, q0=1 << 40)
    for _ in range(warmup):
        g.find_extend_batch(pats, threads=threads)
    t = time.perf_counter()
    for _ in range(steps):
        out = g.find_extend_batch(pats, threads=threads)
    dt = time.perf_counter() - t
    assert np.all(out["end"] > out["start"])
    t1 = time.perf_counter()
    g.find_extend_batch(pats[: max(1, sample // 16)], threads=1)
    single = (time.perf_counter() - t1) / max(1, sample // 16)
    return {"value": sample * steps / dt, "unit": "queries/s", "cores": threads, "kind": "port",
            "sample": f"{sample} length-{K_LEN} patterns of the same workload per step, oracle/gbwt_oracle.c (-O3 -march=native, "
                      f"OpenMP dynamic over queries = rayon analogue), {threads} threads of {os.cpu_count()} host CPUs",
            "single_thread_us_per_query": single * 1e6, "single_thread_ns_per_node": single * 1e9 / K_LEN,
            "seconds": dt, "ms_per_step": dt / steps * 1e3}, g


def run_reference(args, rank, world):
    """--impl reference: the reference's CPU path (C restatement; the Rust crate cannot be built here)."""
    if rank != 0:
        return
    sites, haplotypes, _ = resolve_workload(args)
    from synth import synth
    img = synth.bubble_chain(sites, haplotypes, SEED)
    base, _ = cpu_find_baseline(img.array, sites, haplotypes, args.cpu_sample, steps=args.steps, warmup=args.warmup)
    line = {
        "impl": "reference", "metric": "gbwt_find_queries_per_s_k32", "value": base["value"], "unit": "queries/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": base["ms_per_step"],
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u64", "data": "synthetic",
        "config": workload_config(args, sites, haplotypes, args.cpu_sample, "cpu"),
        "cpu_baseline": {k: base[k] for k in ("value", "unit", "cores", "kind", "sample", "single_thread_us_per_query",
                                              "single_thread_ns_per_node")},
        "e2e": {"value": base["value"], "unit": "queries/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def resolve_workload(args):
    sites, haplotypes, queries = WORKLOADS[args.workload]
    return args.sites or sites, args.haplotypes or haplotypes, args.queries or queries


def workload_config(args, sites, haplotypes, queries, where):
    return {"workload": f"synthetic bubble-chain GBWT, {3 * sites + 1} nodes x {haplotypes} haplotypes ({sites} sites, iid alleles, "
                        f"seed {SEED}), length-{K_LEN} find/extend patterns sampled from the haplotypes (seed {SEED_Q})",
            "baseline_config": "BASELINE.json configs[3]" if args.workload == "find" else args.workload,
            "queries_per_gpu_per_step": queries, "pattern_len": K_LEN, "layout": args.layout, "where": where,
            "l2": "inputs larger than L2 (index + pattern batch >> 126 MB), no flush needed",
            "parallelism": f"index replicated, queries sharded over {args.gpus} GPU(s), no collective on the timed path"}


# ---- the GPU benchmark -------------------------------------------------------------------------------------

def main():
    args = parse_args()
    rank, world, local_rank = dist_env()
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    import gbwt_rs_b200 as gb
    from synth import synth
    ncpu = os.cpu_count() or 8
    os.environ["GBWT_B200_BUILD_THREADS"] = str(max(1, ncpu // max(1, world)))  # K0 runs on every rank at once

    sites, haplotypes, Q = resolve_workload(args)
    image, keep = get_image(sites, haplotypes, rank, world, barrier)
    t = time.time()
    index = gb.GBWT.from_bytes(image, device=local_rank, layout=args.layout)
    stats = index.device_bytes()
    if rank == 0:
        log(f"[bench] device index built in {time.time() - t:.1f} s: {stats}")
    stream = torch.cuda.current_stream().cuda_stream

    if args.workload == "extract-dna":
        line = bench_extract_dna(args, index, image, sites, haplotypes, rank, world, local_rank, barrier, max_over_ranks, stats)
        if rank == 0:
            print(json.dumps(line), flush=True)
        return
    if args.workload == "extract":
        line = bench_extract(args, index, image, sites, haplotypes, rank, world, local_rank, barrier, max_over_ranks, stats)
        if rank == 0:
            print(json.dumps(line), flush=True)
        return

    # inputs resident in HBM before the timed region
    q0 = rank * Q
    d_pat = torch.empty((Q, K_LEN), dtype=torch.int64, device=dev)
    d_out = torch.empty((Q, 3), dtype=torch.int64, device=dev)
    synth.patterns_device(sites, haplotypes, SEED, Q, d_pat.data_ptr(), k=K_LEN, seed_q=SEED_Q, q0=q0, stream=stream)
    torch.cuda.synchronize()

    def step():
        index.find_extend_device(d_pat.data_ptr(), Q, K_LEN, d_out.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    time.sleep(0.25)
    launches0 = gb.kernel_launches()
    events = [torch.cuda.Event(enable_timing=True) for _ in range(args.steps + 1)]
    barrier(); torch.cuda.synchronize()
    t0 = time.time()
    events[0].record()
    for i in range(args.steps):
        step()
        events[i + 1].record()
    torch.cuda.synchronize(); barrier()
    t1 = time.time()
    launches = gb.kernel_launches() - launches0
    clocks = sampler.stop(t0, t1)
    step_ms = [events[i].elapsed_time(events[i + 1]) for i in range(args.steps)]
    total_ms = max_over_ranks(events[0].elapsed_time(events[-1]))
    value = Q * world * args.steps / (total_ms / 1e3)

    # size-independent parity properties on the full batch (every sampled pattern occurs; state.node = last node)
    ok = bool(torch.all(d_out[:, 2] > d_out[:, 1]).item()) and bool(torch.equal(d_out[:, 0], d_pat[:, K_LEN - 1]))
    checksum = int((d_out[:, 2] - d_out[:, 1]).sum().item())
    if not ok:
        raise SystemExit("parity property violated: a sampled pattern was not found")

    # end to end through the host C ABI with pinned buffers (H2D + kernel + D2H inside the timed region)
    e2e = None
    if not args.no_e2e:
        Qe = min(args.e2e_queries, Q)
        h_pat = torch.empty((Qe, K_LEN), dtype=torch.int64).pin_memory()
        h_out = torch.empty((Qe, 3), dtype=torch.int64).pin_memory()
        h_pat.copy_(d_pat[:Qe])
        lib = gb.library()
        def e2e_step():
            rc = lib.gbwt_b200_find_extend(index._h, h_pat.data_ptr(), Qe, K_LEN, h_out.data_ptr())
            if rc != 0:
                raise SystemExit(f"gbwt_b200_find_extend failed: {lib.gbwt_b200_last_error().decode()}")
        e2e_step()
        e2e_steps = max(3, min(args.steps, 10))
        barrier(); torch.cuda.synchronize()
        te = time.perf_counter()
        for _ in range(e2e_steps):
            e2e_step()
        torch.cuda.synchronize()
        dt = max_over_ranks(time.perf_counter() - te)
        barrier()
        if not torch.equal(h_out, d_out[:Qe].cpu()):
            raise SystemExit("host-path results differ from device-path results")
        e2e = {"value": Qe * world * e2e_steps / dt, "unit": "queries/s", "h2d_bytes_per_step": Qe * K_LEN * 8,
               "d2h_bytes_per_step": Qe * 24, "queries_per_gpu_per_step": Qe, "steps": e2e_steps,
               "api": "gbwt_b200_find_extend (host pointers, pinned), chunked double-buffered H2D/kernel/D2H"}

    roofline, cpu = None, None
    if rank == 0:
        # algorithmic bytes per query of SURVEY.md 8(d), on the reference's compressed records, from a sample
        from oracle import oracle as orc
        t = time.time()
        sample_n = 20_000
        if args.no_cpu_baseline or world > 1:
            g = orc.GBWT.load(image, native=True)
        else:
            cpu, g = cpu_find_baseline(image, sites, haplotypes, args.cpu_sample)
        sample = synth.patterns(sites, haplotypes, SEED, n=sample_n, k=K_LEN, seed_q=SEED_Q, q0=q0)
        bytes_per_query = g.find_extend_bytes(sample) / sample_n
        want = g.find_extend_batch(sample)
        got = d_out[:sample_n].cpu().numpy().view(np.uint64)
        if not np.array_equal(got, want.view(np.uint64).reshape(-1, 3)):
            raise SystemExit("parity failure against the oracle on the sampled queries")
        peak, peak_src = measured_peak_gbs()
        traffic = None
        try:  # per-launch DRAM bytes of the dominant kernel from the committed ncu capture of this very step
            with open(os.path.join(ROOT, "profiles", "find_extend_traffic.json")) as f:
                prof = json.load(f)
            if prof["queries_per_launch"] == Q and prof["layout"] == args.layout and args.workload == "find":
                traffic = prof["dram_bytes_read"] + prof["dram_bytes_write"]
        except Exception:
            traffic = None
        launch_s = statistics.mean(step_ms) / 1e3
        achieved = bytes_per_query * Q / launch_s / 1e9
        roofline = {"bound": "hbm", "kernel": "k_find_extend", "achieved": achieved, "peak": peak, "unit": "GB/s",
                    "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src + " (of measured)",
                    "algorithmic_bytes_per_query": bytes_per_query, "algorithmic_bytes_per_lf_step": bytes_per_query / K_LEN,
                    "launch_ms": launch_s * 1e3, "launches_per_step": launches // max(1, args.steps),
                    "note": "achieved = algorithmic bytes of SURVEY.md 8(d) (counted on the reference's compressed records: "
                            "~330 B of a 518 B record per anchor step) / device time of the whole step (bucket pre-pass + "
                            "k_find_extend). It exceeds the HBM peak because the work is not done by moving those bytes: the dense "
                            "device layout answers a rank from one 32-byte block, and the locality schedule makes the batch share "
                            "records through L1/L2, so measured DRAM traffic (`traffic`, ncu) is ~0.45 KB/query against 5.7 KB/query "
                            "algorithmic. The kernel is bound by L1 wavefronts / latency (profiles/README.md), not by HBM"}
        log(f"[bench] oracle sample + cpu baseline took {time.time() - t:.1f} s")

    if rank == 0:
        line = {
            "metric": "gbwt_find_queries_per_s_k32", "value": value, "unit": "queries/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": workload_config(args, sites, haplotypes, Q, "inputs and outputs resident in HBM"),
            "e2e": e2e, "gpu_launches": launches, "roofline": roofline, "cpu_baseline": cpu, "clocks": clocks,
            "extra": {"lf_steps_per_s": value * K_LEN, "occurrences_checksum": checksum, "index_device_bytes": stats,
                      "step_ms": step_ms, "parity": "all patterns found, state.node == last node on the full batch; "
                                                    "first 20000 queries bit-exact against the CPU oracle"},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def bench_extract(args, index, image, sites, haplotypes, rank, world, local_rank, barrier, max_over_ranks, stats):
    """configs[4]: extraction of all forward haplotype paths via LF, partitioned by path."""
    import torch
    import gbwt_rs_b200 as gb
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    lo, hi = gb.shard_range(haplotypes, rank, world)
    m = hi - lo
    length = 2 * sites + 1
    ids = (torch.arange(lo, hi, dtype=torch.int64, device=dev) * 2)
    offs = torch.arange(m + 1, dtype=torch.int64, device=dev) * length
    nodes = torch.empty(m * length, dtype=torch.int64, device=dev)
    lens = torch.empty(m, dtype=torch.int64, device=dev)

    def step():
        index.extract_device(ids.data_ptr(), m, offs.data_ptr(), nodes.data_ptr(), lens.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    launches0 = gb.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(); barrier()
    t1 = time.time()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t0, t1)
    assert bool(torch.all(lens == length).item())
    from synth import synth
    first = nodes[:length].cpu().numpy().view(np.uint64)
    assert np.array_equal(first, synth.sequence(sites, haplotypes, SEED, 2 * lo))
    steps_total = haplotypes * length * args.steps
    value = steps_total / (total_ms / 1e3)
    roofline = None
    if rank == 0:
        from oracle import oracle as orc
        g = orc.GBWT.load(image, native=True)
        sample_ids = np.arange(0, 2 * haplotypes, max(2, 2 * (haplotypes // 8)), dtype=np.uint64)
        per_step = g.extract_bytes(sample_ids) / (len(sample_ids) * length)
        peak, src = measured_peak_gbs()
        achieved = per_step * haplotypes * length / world / (total_ms / args.steps / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "k_extract_split", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                    "traffic": None, "peak_source": src, "algorithmic_bytes_per_lf_step": per_step,
                    "note": "a path walk is a dependent chain: the bound that matters is chains in flight / access latency"}
    return {"metric": "gbwt_extract_lf_steps_per_s", "value": value, "unit": "LF steps/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u32", "data": "synthetic",
            "config": {"workload": f"extraction of all {haplotypes} forward haplotype paths ({length} nodes each) of the "
                                   f"{3 * sites + 1}-node bubble-chain GBWT", "baseline_config": "BASELINE.json configs[4]",
                       "paths_per_gpu": m, "layout": args.layout,
                       "note": "path lengths are known to the index after the first (warm-up) extraction, so every path is "
                               "walked from both ends (k_extract_split)"},
            "gpu_launches": gb.kernel_launches() - launches0, "roofline": roofline, "clocks": clocks,
            "extra": {"index_device_bytes": stats}}


def bench_extract_dna(args, index, image, sites, haplotypes, rank, world, local_rank, barrier, max_over_ranks, stats):
    """DNA-level extraction (extract_sequence of src/bin/gbz-extract.rs:173-189) of all forward
    haplotype paths of the configs[4] graph with synthetic node labels, partitioned by path."""
    import torch
    import gbwt_rs_b200 as gb
    from synth import synth
    dev = torch.device("cuda", local_rank)
    stream = torch.cuda.current_stream().cuda_stream
    starts, labels = synth.node_labels(3 * sites + 1, seed=SEED, max_anchor=32)
    index.attach_graph(starts, labels)
    lo, hi = gb.shard_range(haplotypes, rank, world)
    m = hi - lo
    length = 2 * sites + 1
    ids = (torch.arange(lo, hi, dtype=torch.int64, device=dev) * 2)
    lens = torch.empty(m, dtype=torch.int64, device=dev)
    index.dna_lengths_device(ids.data_ptr(), m, lens.data_ptr(), stream)
    torch.cuda.synchronize()
    offs = torch.zeros(m + 1, dtype=torch.int64, device=dev)
    offs[1:] = torch.cumsum(lens, 0)
    total_bytes = int(offs[-1].item())
    out = torch.empty(total_bytes, dtype=torch.uint8, device=dev)
    got = torch.empty(m, dtype=torch.int64, device=dev)

    def step():
        index.extract_dna_device(ids.data_ptr(), m, 0, offs.data_ptr(), out.data_ptr(), got.data_ptr(), stream)

    for _ in range(args.warmup):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(local_rank)
    launches0 = gb.kernel_launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier(); torch.cuda.synchronize()
    t0 = time.time()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    torch.cuda.synchronize(); barrier()
    t1 = time.time()
    total_ms = max_over_ranks(e0.elapsed_time(e1))
    clocks = sampler.stop(t0, t1)
    assert bool(torch.all(got == lens).item())
    launches = gb.kernel_launches() - launches0
    # the same walks without the byte copy (label ranges and lengths only, one-ended), for the cost split
    l0, l1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0.record()
    index.dna_lengths_device(ids.data_ptr(), m, got.data_ptr(), stream)
    l1.record()
    torch.cuda.synchronize()
    lengths_only_ms = l0.elapsed_time(l1)
    all_bytes = max_over_ranks(float(total_bytes)) * world if world > 1 else float(total_bytes)
    value = all_bytes * args.steps / (total_ms / 1e3)
    cpu_baseline = None
    if rank == 0:
        # parity of the first and last path of this rank against the oracle, and the CPU baseline on a sample
        from oracle import oracle as orc
        gbz = synth.gbz_image(image, starts, labels, 3, as_image=True)
        g = orc.GBWT.load(gbz.array, native=True)
        sample = np.array([2 * lo, 2 * (hi - 1)], dtype=np.uint64)
        t = time.time()
        w_offsets, w_bytes, _ = g.extract_dna_batch(sample, 0, threads=2)
        dt = time.time() - t
        for j, i in enumerate((0, m - 1)):
            a, b = int(offs[i].item()), int(offs[i + 1].item())
            assert np.array_equal(out[a:b].cpu().numpy(), w_bytes[int(w_offsets[j]):int(w_offsets[j + 1])])
        cpu_baseline = {"value": float(w_offsets[-1]) / dt, "unit": "bases/s", "cores": 2, "kind": "port",
                        "sample": "2 of the paths, one thread each"}
    return {"metric": "gbz_extract_dna_bases_per_s", "value": value, "unit": "bases/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"DNA sequences of all {haplotypes} forward haplotype paths ({length} nodes, "
                                   f"{total_bytes // max(1, m)} bases each) of the {3 * sites + 1}-node bubble-chain GBZ",
                       "paths_per_gpu": m, "layout": args.layout, "label_bytes": int(starts[-1]),
                       "note": "sequence and DNA lengths are known to the index from the dna_lengths call that sized the output, "
                               "so every path is walked from both ends by two warps that relay their nodes to a spelling warp (k_extract_dna_relay)"},
            "gpu_launches": launches, "clocks": clocks, "cpu_baseline": cpu_baseline,
            "extra": {"lf_steps_per_s": haplotypes * length * args.steps / (total_ms / 1e3), "lengths_only_ms": lengths_only_ms,
                      "index_device_bytes": stats,
                      "output_bytes_per_gpu": total_bytes}}


if __name__ == "__main__":
    main()
