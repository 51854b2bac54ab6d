"""world_size-2 gloo test of the N > 1 host logic: each rank takes its shard_range block of a query batch
and of a set of sequence ids, runs it, and the results are gathered on rank 0 and compared with the unsharded
answer. On this CPU box the per-rank engine is tests/hostsim (the product's one-lane code on the CPU); the
sharding, barrier and gather code is exactly what bench.py and a multi-GPU caller use."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, result_path):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    import gbwt_rs_b200 as gb
    from hostsim_build import HostSim
    from synth import synth
    S, H, seed, n = 300, 16, 42, 5001
    img = synth.bubble_chain(S, H, seed)
    engine = HostSim(img.array)                      # index replicated on every rank
    lo, hi = gb.shard_range(n, rank, world)           # queries sharded
    pats = synth.patterns(S, H, seed, n=hi - lo, k=32, q0=lo)
    local = engine.find_extend(pats)
    dist.barrier()
    gathered = gb.gather_states(local, n)
    plo, phi = gb.shard_range(2 * H, rank, world)     # paths partitioned by sequence id
    ids = np.arange(plo, phi, dtype=np.uint64)
    lengths = engine.sequence_lengths(ids)
    all_lengths = gb.gather_states(lengths.reshape(-1, 1), 2 * H)
    if rank == 0:
        np.save(result_path + ".states.npy", gathered)
        np.save(result_path + ".lengths.npy", all_lengths)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_sharded_search_equals_unsharded(tmp_path, world):
    import torch.multiprocessing as mp
    from hostsim_build import HostSim
    from oracle import oracle as orc
    from synth import synth
    result = str(tmp_path / "result")
    mp.spawn(_worker, args=(world, _free_port(), result), nprocs=world, join=True)
    S, H, seed, n = 300, 16, 42, 5001
    img = synth.bubble_chain(S, H, seed)
    g = orc.GBWT.load(img.array)
    pats = synth.patterns(S, H, seed, n=n, k=32)
    want = g.find_extend_batch(pats).view(np.uint64).reshape(-1, 3)
    got = np.load(result + ".states.npy")
    assert np.array_equal(got, want)
    lengths = np.load(result + ".lengths.npy")
    assert np.all(lengths == 2 * S + 1) and len(lengths) == 2 * H
